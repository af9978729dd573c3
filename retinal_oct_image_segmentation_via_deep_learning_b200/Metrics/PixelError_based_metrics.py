"""GPU drop-in for the reference's ``Metrics/PixelError_based_metrics.py`` (MSE, RMSE).

Binary masks go through the K=2 confusion kernel ((FP + FN) / size is exactly the reference's
float mean); integer arrays such as ``(K-1, W)`` boundary positions go through
``octm_boundary_error_i32`` (exact int64 sums, one float64 division); floating arrays (soft boundary
positions of a layer model) through ``octm_boundary_error_float`` (float64 sums, fixed order).
"""
import numpy as np

from retinal_oct_image_segmentation_via_deep_learning_b200 import _dropin, derive


def mean_squared_error(y_true, y_pred):
    """mean((y_true - y_pred)**2) -- reference PixelError_based_metrics.py:3-19."""
    if _dropin.is_binary_like(y_true) and _dropin.is_binary_like(y_pred):
        return np.float64(derive.count_metrics(*_dropin.binary_counts(y_true, y_pred))["mean_squared_error"])
    sq, _, n = _dropin.error_sums(y_true, y_pred)
    return np.float64(sq) / n if n else np.float64(np.nan)


def root_mean_squared_error(y_true, y_pred):
    """sqrt(mean_squared_error) -- reference :21-37."""
    return np.sqrt(mean_squared_error(y_true, y_pred))
