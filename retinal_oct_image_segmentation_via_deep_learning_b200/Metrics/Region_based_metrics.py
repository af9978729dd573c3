"""GPU drop-in for the reference's ``Metrics/Region_based_metrics.py`` (Dice, IoU, precision, recall).

One K=2 confusion kernel per call; ratios formed in float64 with the reference's operation order.
"""
import numpy as np

from retinal_oct_image_segmentation_via_deep_learning_b200 import _dropin, derive


def _scalars(y_true, y_pred):
    return derive.count_metrics(*_dropin.binary_counts(y_true, y_pred))


def dice_coefficient(y_true, y_pred):
    """2 |X n Y| / (|X| + |Y| + 1e-7) -- reference Region_based_metrics.py:3-16."""
    return np.float64(_scalars(y_true, y_pred)["dice_coefficient"])


def iou_score(y_true, y_pred):
    """|X n Y| / (|X| + |Y| - |X n Y| + 1e-7) -- reference :18-31."""
    return np.float64(_scalars(y_true, y_pred)["iou_score"])


def precision(y_true, y_pred):
    """|X n Y| / (|Y| + 1e-7) -- reference :33-46."""
    return np.float64(_scalars(y_true, y_pred)["region_precision"])


def recall(y_true, y_pred):
    """|X n Y| / (|X| + 1e-7) -- reference :48-61."""
    return np.float64(_scalars(y_true, y_pred)["recall"])
