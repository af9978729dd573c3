"""GPU drop-in for the reference's ``Metrics/ConfusionMatrix_based_metrics.py``.

Same function names and ``f(y_true, y_pred)`` signatures; each call runs one K=2 confusion-matrix
kernel on the B200 (``octm_confusion_u8``) and evaluates the reference's expression from the exact
counts in float64.  Masks must be binary (the reference is only meaningful for 0/1 data).
"""
import numpy as np

from retinal_oct_image_segmentation_via_deep_learning_b200 import _dropin, derive


def _scalars(y_true, y_pred):
    return derive.count_metrics(*_dropin.binary_counts(y_true, y_pred))


def accuracy(y_true, y_pred):
    """(TP + TN) / size -- reference ConfusionMatrix_based_metrics.py:4-18."""
    return np.float64(_scalars(y_true, y_pred)["accuracy"])


def sensitivity(y_true, y_pred):
    """TP / (TP + FN + 1e-7) -- reference :20-33."""
    return np.float64(_scalars(y_true, y_pred)["sensitivity"])


def precision(y_true, y_pred):
    """TP / (TP + FP + 1e-7) -- reference :35-48."""
    return np.float64(_scalars(y_true, y_pred)["cm_precision"])


def specificity(y_true, y_pred):
    """TN / (TN + FP + 1e-7) -- reference :50-63."""
    return np.float64(_scalars(y_true, y_pred)["specificity"])


def auc_score(y_true, y_pred):
    """Area under the ROC curve of a score / probability map -- reference :65-84
    (``roc_auc_score(y_true.flatten(), y_pred.flatten())``, ``except ValueError: return 0.0``).

    One CTA-wide radix sort + tie-aware rank sum on the GPU (``octm_auc_u8``).  Like the reference it
    returns a Python float; the cases in which scikit-learn raises ValueError (more than two label
    values, NaN / inf scores, mismatching lengths) return 0.0.  A single-class ``y_true`` returns NaN, as
    the reference does with the scikit-learn of this image (>= 1.6 warns instead of raising)."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    dev = _dropin._device()
    yt = y_true if isinstance(y_true, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(y_true)))
    sc = y_pred if isinstance(y_pred, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(y_pred)))
    if yt.numel() != sc.numel():
        return 0.0                                      # check_consistent_length -> ValueError
    if yt.numel() == 0:
        return 0.0                                      # "Found array with 0 sample(s)" -> ValueError
    if yt.dtype != torch.bool and yt.dtype != torch.uint8:
        # any two distinct values are a binary problem for scikit-learn; rank them into uint8
        vals = torch.unique(yt)
        if vals.numel() > 2:
            return 0.0                                  # multiclass without multi_class= -> ValueError
        if vals.is_floating_point() and not bool(torch.isfinite(vals).all()):
            return 0.0
        yt = (yt == vals[-1]).to(torch.uint8) if vals.numel() == 2 else torch.zeros(yt.shape, dtype=torch.uint8)
    if not sc.is_floating_point():
        sc = sc.to(torch.float64)
    yt = yt.reshape(1, -1).to(dev)
    sc = sc.reshape(1, -1).to(dev)
    return float(suite.auc_scores(yt, sc)[0].item())
