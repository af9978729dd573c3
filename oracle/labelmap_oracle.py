"""Label-map level oracle (TEST ORACLE): how a user of the reference scores a K-class B-scan.

The reference only ships binary-mask functions ``f(y_true, y_pred)``; a K-class label map is
scored by looping ``for c in range(K): f(gt == c, pr == c)``.  This
module does exactly that with the restated functions of ``metrics_oracle`` and also provides
the integer intermediates the CUDA path must reproduce bit-exactly:

* ``confusion_matrix``      -- K x K joint histogram ``cm[t][p]``
* ``column_thickness``      -- per-class per-column pixel counts (``np.sum(mask, axis=0)``,
                               Biomarker_based_metrics.py:14-15)
* ``boundaries``            -- BUILD-DEFINED (no reference counterpart):
                               ``b_k(x) = #{y : L[y, x] < k}`` for ``k = 1..K-1``
* ``contour_intermediates`` -- doubled-lattice vertices of contour ``[0]`` and the exact squared
                               distances behind hausdorff / hd95 / assd
"""
from __future__ import annotations

import numpy as np

from . import metrics_oracle as mo
from .contours_oracle import directed_sq_distances, first_contour_lattice

COUNT_METRICS = ("accuracy", "sensitivity", "cm_precision", "specificity", "dice_coefficient",
                 "iou_score", "region_precision", "recall", "mean_squared_error",
                 "root_mean_squared_error", "mad", "vascularity_index")
CONTOUR_METRICS = ("hausdorff_distance", "hausdorff_distance_95", "assd")


def confusion_matrix(y_true, y_pred, num_classes):
    t = np.asarray(y_true).astype(np.int64).ravel()
    p = np.asarray(y_pred).astype(np.int64).ravel()
    if t.size and (t.max() >= num_classes or p.max() >= num_classes):
        raise ValueError("label >= num_classes")
    return np.bincount(t * num_classes + p, minlength=num_classes * num_classes) \
        .reshape(num_classes, num_classes).astype(np.uint64)


def column_thickness(label_map, num_classes):
    """int64 [K, W]: pixels of each class in every column (A-scan) of one (H, W) B-scan."""
    lm = np.asarray(label_map)
    return np.stack([np.sum((lm == c).astype(np.int64), axis=0) for c in range(num_classes)])


def boundaries(label_map, num_classes):
    """int32 [K-1, W]: ``b_k(x) = #{y : L[y, x] < k}``, k = 1..K-1 (build-defined)."""
    lm = np.asarray(label_map)
    return np.stack([np.sum(lm < k, axis=0) for k in range(1, num_classes)]).astype(np.int32)


def contour_intermediates(mask_true, mask_pred):
    """Exact integers behind the three contour metrics for one binary mask pair.

    Returns None when either mask has no contour (the reference raises IndexError)."""
    try:
        ct = first_contour_lattice(mask_true)
        cp = first_contour_lattice(mask_pred)
    except IndexError:
        return None
    d_p2t = directed_sq_distances(ct, cp)     # for each pred vertex: nearest true vertex (d1)
    d_t2p = directed_sq_distances(cp, ct)     # for each true vertex: nearest pred vertex (d2)
    return {"verts_true": ct, "verts_pred": cp, "sq_pred_to_true": d_p2t, "sq_true_to_pred": d_t2p}


def contour_metrics_from_sq(d_p2t, d_t2p):
    """hausdorff / hd95 / assd from exact squared doubled-lattice distances.

    ``sqrt(D2 / 4.0)`` equals the reference's float64 ``min(sqrt(sum((A - p)**2)))`` bit-exactly
    (SURVEY.md 8a-C); the tail follows Contour_based_metrics.py:22, 39, 56."""
    d1 = np.sqrt(np.asarray(d_p2t, dtype=np.float64) / 4.0)
    d2 = np.sqrt(np.asarray(d_t2p, dtype=np.float64) / 4.0)
    return {
        "hausdorff_distance": max(np.max(d1), np.max(d2)),
        "hausdorff_distance_95": max(np.percentile(d1, 95), np.percentile(d2, 95)),
        "assd": (np.mean(d1) + np.mean(d2)) / 2,
    }


def score_bscan(y_true, y_pred, num_classes, contours=True, functions=None):
    """Score one (H, W) label-map pair the reference way: per class, per function.

    ``functions`` may be a module-like namespace providing the reference's function names (the
    real reference modules when available); default is the restated ``metrics_oracle``."""
    f = functions or mo
    yt, yp = np.asarray(y_true), np.asarray(y_pred)
    res = {name: np.full(num_classes, np.nan) for name in COUNT_METRICS + ("thickness_difference",)}
    if contours:
        res.update({name: np.full(num_classes, np.nan) for name in CONTOUR_METRICS})
    for c in range(num_classes):
        mt = (yt == c).astype(np.int64)
        mp = (yp == c).astype(np.int64)
        for name in COUNT_METRICS + ("thickness_difference",):
            res[name][c] = getattr(f, name)(mt, mp)
        if contours:
            for name in CONTOUR_METRICS:
                try:
                    res[name][c] = getattr(f, name)(mt, mp)
                except IndexError:          # class absent (or filling) a map: no contour
                    pass
    bt, bp = boundaries(yt, num_classes), boundaries(yp, num_classes)
    res["boundary_true"], res["boundary_pred"] = bt, bp
    res["boundary_mse"] = np.array([f.mean_squared_error(bt[k], bp[k]) for k in range(num_classes - 1)])
    res["boundary_rmse"] = np.array([f.root_mean_squared_error(bt[k], bp[k]) for k in range(num_classes - 1)])
    res["boundary_mad"] = np.array([f.mad(bt[k], bp[k]) for k in range(num_classes - 1)])
    res["confusion"] = confusion_matrix(yt, yp, num_classes)
    return res


def score_bscan_fast(y_true, y_pred, num_classes):
    """Integer intermediates only (vectorised; for large parity cases)."""
    yt, yp = np.asarray(y_true), np.asarray(y_pred)
    tt, tp_ = column_thickness(yt, num_classes), column_thickness(yp, num_classes)
    bt, bp = boundaries(yt, num_classes), boundaries(yp, num_classes)
    d = bt.astype(np.int64) - bp.astype(np.int64)
    return {
        "confusion": confusion_matrix(yt, yp, num_classes),
        "thickness_true": tt, "thickness_pred": tp_,
        "thickness_absdiff": np.abs(tt - tp_).sum(axis=1),
        "boundary_true": bt, "boundary_pred": bp,
        "boundary_sq": (d * d).sum(axis=1), "boundary_abs": np.abs(d).sum(axis=1),
    }


def labels_from_boundaries(boundaries, height):
    """label[i, y, x] = #{k : b_k(i, x) <= y}; NaN boundaries are counted for no row (build-defined, the inverse of
    boundary extraction on layered maps; include/octm.h octm_labels_from_boundaries)."""
    b = np.asarray(boundaries)
    n, kb, w = b.shape
    y = np.arange(int(height), dtype=np.float64)[None, :, None]
    lab = np.zeros((n, int(height), w), dtype=np.uint8)
    for k in range(kb):
        bk = b[:, k, None, :].astype(np.float64)
        with np.errstate(invalid="ignore"):
            lab += (bk <= y).astype(np.uint8)
    return lab
