"""GPU drop-in for the reference's ``Metrics/Contour_based_metrics.py``.

``hausdorff_distance`` / ``hausdorff_distance_95`` / ``assd`` keep the reference's literal semantics:
only contour ``[0]`` of ``skimage.measure.find_contours(mask, 0.5)`` of each mask is used (traced on
the GPU, ``octm_contour2d_trace_u8``), distances are exact integers on the doubled lattice
(``octm_contour2d_distance``), and the final max / percentile / mean follow reference lines 22, 39, 56.
"""
import numpy as np

from retinal_oct_image_segmentation_via_deep_learning_b200 import _dropin, derive


def hausdorff_distance(y_true, y_pred):
    """max(d1, d2) of the directed max-min contour distances -- reference :5-22."""
    return float(_dropin.contour_scalars(y_true, y_pred)["hausdorff_distance"])


def hausdorff_distance_95(y_true, y_pred):
    """max of the two directed 95th percentiles (numpy linear method) -- reference :24-39."""
    return float(_dropin.contour_scalars(y_true, y_pred)["hausdorff_distance_95"])


def assd(y_true, y_pred):
    """mean of the two directed mean distances -- reference :41-56."""
    return np.float64(_dropin.contour_scalars(y_true, y_pred)["assd"])


def mad(y_true, y_pred):
    """mean(|y_true - y_pred|) -- reference :58-73.  Binary masks use the confusion kernel,
    integer arrays (boundary positions) the boundary-error kernel."""
    if _dropin.is_binary_like(y_true) and _dropin.is_binary_like(y_pred):
        return np.float64(derive.count_metrics(*_dropin.binary_counts(y_true, y_pred))["mad"])
    _, ab, n = _dropin.error_sums(y_true, y_pred)
    return np.float64(ab) / n if n else np.float64(np.nan)
