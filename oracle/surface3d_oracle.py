"""CPU oracle of the 3-D surface-distance metrics (TEST ORACLE -- never imported by the product).

The reference has no 3-D metric at all (its contour functions raise on ``ndim != 2``), so this is the
build's own definition, the customary one (e.g. medpy's ``__surface_distances``):

    surface(mask) = mask & ~scipy.ndimage.binary_erosion(mask)         # 6-connectivity, border eroded
    d(A -> B)     = scipy.ndimage.distance_transform_edt(~surface(B))[surface(A)]

and then exactly the reference's 2-D recipes on the two distance lists: ``max(max d1, max d2)``
(Contour_based_metrics.py:22), ``max(percentile(d1, 95), percentile(d2, 95))`` (:39),
``(mean(d1) + mean(d2)) / 2`` (:53-56).  Parity unpinned by the reference; pinned to scipy's EDT.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def surface(mask):
    mask = np.asarray(mask, dtype=bool)
    return mask & ~ndimage.binary_erosion(mask)


def squared_distances(mask_from, mask_to):
    """int64 squared distances from every surface voxel of ``mask_from`` to the surface of ``mask_to``
    (None when either surface is empty)."""
    sa, sb = surface(mask_from), surface(mask_to)
    if not sa.any() or not sb.any():
        return None
    dt = ndimage.distance_transform_edt(~sb)
    return np.rint(dt[sa] ** 2).astype(np.int64)


def class_metrics(vol_true, vol_pred, cls):
    """dict with the integer intermediates and hd / hd95 / assd of one class (NaN when undefined)."""
    d1 = squared_distances(vol_pred == cls, vol_true == cls)     # pred surface -> true surface (direction 0)
    d2 = squared_distances(vol_true == cls, vol_pred == cls)     # direction 1
    if d1 is None or d2 is None:
        return {"sq_pred_to_true": None, "sq_true_to_pred": None, "hausdorff_distance": np.nan,
                "hausdorff_distance_95": np.nan, "assd": np.nan}
    e1, e2 = np.sqrt(d1.astype(np.float64)), np.sqrt(d2.astype(np.float64))
    return {"sq_pred_to_true": d1, "sq_true_to_pred": d2,
            "hausdorff_distance": max(e1.max(), e2.max()),
            "hausdorff_distance_95": max(np.percentile(e1, 95), np.percentile(e2, 95)),
            "assd": (np.mean(e1) + np.mean(e2)) / 2}
