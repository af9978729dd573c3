"""GPU drop-in for the reference's ``Metrics/Biomarker_based_metrics.py``.

``thickness_difference`` runs the column-scan kernel (``octm_column_scan_u8``) on the pair viewed as
``(shape[0], prod(shape[1:]))`` -- the reference sums over axis 0 -- and returns the signed-integer
result (the reference's uint8 path wraps; its bool/int64 path is what is reproduced).
"""
import numpy as np

from retinal_oct_image_segmentation_via_deep_learning_b200 import _dropin, derive, suite


def thickness_difference(y_true, y_pred):
    """mean over columns of |sum_rows(y_true) - sum_rows(y_pred)| -- reference :3-21."""
    t, p = _dropin.as_mask_u8(y_true, "y_true"), _dropin.as_mask_u8(y_pred, "y_pred")
    if t.shape != p.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(t.shape)} {tuple(p.shape)}")
    if t.dim() == 0:
        raise ValueError("axis 0 is out of bounds for array of dimension 0")
    h = t.shape[0]
    w = t.numel() // h if h else 0
    if h == 0 or w == 0:
        return np.float64(np.nan) if w == 0 else np.float64(0.0)
    lp = suite.label_pass(t.reshape(1, h, w), p.reshape(1, h, w), 2, counts=False, columns=True)
    return np.float64(derive.thickness_difference(lp.thick_absdiff.cpu().numpy()[0, 1], w))


def vascularity_index(y_true, y_pred):
    """|sum(y_true)/size - sum(y_pred)/size| -- reference :23-38."""
    return np.float64(derive.count_metrics(*_dropin.binary_counts(y_true, y_pred))["vascularity_index"])
