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
    """Area under the ROC curve of a probability map -- reference :65-84 (sklearn.roc_auc_score).

    Not on the label-map hot path; the GPU rank-statistic kernel is not built
    yet, and this package has no CPU fallback."""
    raise NotImplementedError("auc_score: GPU implementation pending (SURVEY.md 8f rank 2)")
