"""Numpy restatement of the reference's scalar metric formulas (TEST ORACLE).

Every function takes two same-shape binary masks and returns the scalar the
reference returns for them.  The arithmetic follows the reference expression
by expression so that results are bit-identical; citations are relative to
``/root/reference/Metrics``.  Checked against the executed reference by
``oracle/make_golden.py`` / ``tests/test_oracle_vs_reference.py``.

The masks are promoted to int64 first: the reference's uint8 path wraps in
``1 - y`` and in ``thickness_difference``, and its
int64/bool/float64 paths all agree to 0 ulp, so int64 is the canonical input.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-7  # the reference's denominator guard, e.g. ConfusionMatrix_based_metrics.py:32


def _i64(a):
    return np.asarray(a).astype(np.int64)


def _overlap_counts(y_true, y_pred):
    """TP, FP, FN, TN as numpy int64 scalars (sum-of-products form of the reference)."""
    t, p = _i64(y_true), _i64(y_pred)
    tp = np.sum(t * p)
    fp = np.sum((1 - t) * p)
    fn = np.sum(t * (1 - p))
    tn = np.sum((1 - t) * (1 - p))
    return tp, fp, fn, tn


# ---------------------------------------------------------------- ConfusionMatrix_based_metrics.py
def accuracy(y_true, y_pred):
    """ConfusionMatrix_based_metrics.py:14-17 -- (TP + TN) / prod(shape), no epsilon."""
    tp, _, _, tn = _overlap_counts(y_true, y_pred)
    return (tp + tn) / np.prod(np.asarray(y_true).shape)


def sensitivity(y_true, y_pred):
    """ConfusionMatrix_based_metrics.py:30-32 -- TP / (TP + FN + 1e-7)."""
    tp, _, fn, _ = _overlap_counts(y_true, y_pred)
    return tp / (tp + fn + EPS)


def cm_precision(y_true, y_pred):
    """ConfusionMatrix_based_metrics.py:45-47 -- TP / (TP + FP + 1e-7)."""
    tp, fp, _, _ = _overlap_counts(y_true, y_pred)
    return tp / (tp + fp + EPS)


def specificity(y_true, y_pred):
    """ConfusionMatrix_based_metrics.py:60-62 -- TN / (TN + FP + 1e-7)."""
    _, fp, _, tn = _overlap_counts(y_true, y_pred)
    return tn / (tn + fp + EPS)


def auc_score(y_true, y_pred, single_class_value=float("nan")):
    """ConfusionMatrix_based_metrics.py:65-84 -- ``roc_auc_score(y_true.flatten(), y_pred.flatten())``,
    ``except ValueError: return 0.0``.

    scikit-learn (unpinned by the reference; restated from ``sklearn/metrics/_ranking.py``:
    ``_binary_clf_curve`` + ``roc_curve`` + ``auc``): stable sort by descending score, one ROC point per
    DISTINCT score (tps = positives with score >= threshold, fps likewise), trapezoidal area of tpr
    over fpr.  ValueError cases (-> 0.0): non-finite scores, more than two label values, length
    mismatch, empty input.  A single-class y_true is the version-dependent case (>= 1.6: NaN + warning;
    before: ValueError -> 0.0), hence ``single_class_value``."""
    t = np.asarray(y_true).reshape(-1)
    s = np.asarray(y_pred).reshape(-1).astype(np.float64)
    if t.size != s.size or t.size == 0 or not np.all(np.isfinite(s)):
        return 0.0
    vals = np.unique(t)
    if vals.size > 2:
        return 0.0
    if vals.size < 2:
        return single_class_value
    pos = (t == vals[-1]).astype(np.float64)
    order = np.argsort(s, kind="mergesort")[::-1]
    s_sorted, pos_sorted = s[order], pos[order]
    distinct = np.where(np.diff(s_sorted))[0]
    idx = np.r_[distinct, pos_sorted.size - 1]
    tps = np.cumsum(pos_sorted)[idx]
    fps = 1 + idx - tps
    tps, fps = np.r_[0, tps], np.r_[0, fps]
    fpr, tpr = fps / fps[-1], tps / tps[-1]
    return float(np.trapezoid(tpr, fpr)) if hasattr(np, "trapezoid") else float(np.trapz(tpr, fpr))


def auc_rank_sum(y_true, y_pred):
    """The same area as an exact rational: (2 * Mann-Whitney U with ties counted one half, n_pos, n_neg)
    as Python ints -- what the CUDA kernel accumulates before its single division."""
    t = np.asarray(y_true).reshape(-1)
    s = np.asarray(y_pred).reshape(-1).astype(np.float64)
    vals = np.unique(t)
    pos = t == vals[-1]
    neg_scores = np.sort(s[~pos])
    less = np.searchsorted(neg_scores, s[pos], side="left")
    leq = np.searchsorted(neg_scores, s[pos], side="right")
    return int(less.sum()) + int(leq.sum()), int(pos.sum()), int((~pos).sum())


# ---------------------------------------------------------------- Region_based_metrics.py
def dice_coefficient(y_true, y_pred):
    """Region_based_metrics.py:13-15 -- 2 I / (sum(t) + sum(p) + 1e-7)."""
    t, p = _i64(y_true), _i64(y_pred)
    inter = np.sum(t * p)
    return (2.0 * inter) / (np.sum(t) + np.sum(p) + EPS)


def iou_score(y_true, y_pred):
    """Region_based_metrics.py:28-30 -- I / (sum(t) + sum(p) - I + 1e-7)."""
    t, p = _i64(y_true), _i64(y_pred)
    inter = np.sum(t * p)
    return inter / (np.sum(t) + np.sum(p) - inter + EPS)


def region_precision(y_true, y_pred):
    """Region_based_metrics.py:43-45 -- I / (sum(p) + 1e-7)."""
    t, p = _i64(y_true), _i64(y_pred)
    return np.sum(t * p) / (np.sum(p) + EPS)


def recall(y_true, y_pred):
    """Region_based_metrics.py:58-60 -- I / (sum(t) + 1e-7)."""
    t, p = _i64(y_true), _i64(y_pred)
    return np.sum(t * p) / (np.sum(t) + EPS)


# ---------------------------------------------------------------- PixelError_based_metrics.py
def mean_squared_error(y_true, y_pred):
    """PixelError_based_metrics.py:14-17 -- mean((t - p)**2) in float64."""
    d = np.asarray(y_true).astype(float) - np.asarray(y_pred).astype(float)
    return np.mean(d ** 2)


def root_mean_squared_error(y_true, y_pred):
    """PixelError_based_metrics.py:32-35 -- sqrt of the above."""
    return np.sqrt(mean_squared_error(y_true, y_pred))


# ---------------------------------------------------------------- Contour_based_metrics.py:58-73
def mad(y_true, y_pred):
    """Contour_based_metrics.py:68-71 -- mean(|t - p|) in float64."""
    d = np.asarray(y_true).astype(float) - np.asarray(y_pred).astype(float)
    return np.mean(np.abs(d))


# ---------------------------------------------------------------- Biomarker_based_metrics.py
def thickness_difference(y_true, y_pred):
    """Biomarker_based_metrics.py:14-21 -- mean over columns of |colsum(t) - colsum(p)|, axis 0."""
    t, p = _i64(y_true), _i64(y_pred)
    return np.mean(np.abs(np.sum(t, axis=0) - np.sum(p, axis=0)))


def vascularity_index(y_true, y_pred):
    """Biomarker_based_metrics.py:34-38 -- |sum(t)/size - sum(p)/size| (two divisions, then subtract)."""
    t, p = _i64(y_true), _i64(y_pred)
    return np.abs(np.sum(t) / t.size - np.sum(p) / p.size)


# ---------------------------------------------------------------- Contour_based_metrics.py:5-56
def _directed_min_distances(contour_a, contour_b):
    """For each vertex of ``contour_b`` the Euclidean distance to the nearest vertex of
    ``contour_a`` -- the list comprehension of Contour_based_metrics.py:19-20 (also 36-37, 53-54)."""
    return [np.min(np.sqrt(np.sum((contour_a - q) ** 2, axis=1))) for q in contour_b]


def _first_contours(y_true, y_pred):
    """Contour_based_metrics.py:15-16 -- ``find_contours(mask, 0.5)[0]`` for both masks.
    Raises IndexError when a mask has no iso-contour (all 0 or all 1), like the reference."""
    from .contours_oracle import find_contours
    return find_contours(np.asarray(y_true), 0.5)[0], find_contours(np.asarray(y_pred), 0.5)[0]


def hausdorff_distance(y_true, y_pred):
    """Contour_based_metrics.py:15-22."""
    ct, cp = _first_contours(y_true, y_pred)
    d1 = np.max(_directed_min_distances(ct, cp))
    d2 = np.max(_directed_min_distances(cp, ct))
    return max(d1, d2)


def hausdorff_distance_95(y_true, y_pred):
    """Contour_based_metrics.py:33-39 -- numpy's default (linear) percentile on each direction."""
    ct, cp = _first_contours(y_true, y_pred)
    d1 = _directed_min_distances(ct, cp)
    d2 = _directed_min_distances(cp, ct)
    return max(np.percentile(d1, 95), np.percentile(d2, 95))


def assd(y_true, y_pred):
    """Contour_based_metrics.py:50-56 -- mean of the two directed means (not the pooled mean)."""
    ct, cp = _first_contours(y_true, y_pred)
    d1 = np.mean(_directed_min_distances(ct, cp))
    d2 = np.mean(_directed_min_distances(cp, ct))
    return (d1 + d2) / 2


# ---------------------------------------------------------------- closed forms over confusion counts
def from_counts(tp, fp, fn, tn):
    """The ten count-derived scalars evaluated from integer TP/FP/FN/TN with the reference's
    operation order.  Used to pin the closed forms the GPU host epilogue uses."""
    tp, fp, fn, tn = (np.int64(v) for v in (tp, fp, fn, tn))
    n = tp + fp + fn + tn
    st, sp = tp + fn, tp + fp
    with np.errstate(divide="ignore", invalid="ignore"):
        return {
            "accuracy": (tp + tn) / n,
            "sensitivity": tp / (tp + fn + EPS),
            "cm_precision": tp / (tp + fp + EPS),
            "specificity": tn / (tn + fp + EPS),
            "dice_coefficient": (2.0 * tp) / (st + sp + EPS),
            "iou_score": tp / (st + sp - tp + EPS),
            "region_precision": tp / (sp + EPS),
            "recall": tp / (st + EPS),
            "mean_squared_error": np.float64(fp + fn) / n,
            "root_mean_squared_error": np.sqrt(np.float64(fp + fn) / n),
            "mad": np.float64(fp + fn) / n,
            "vascularity_index": np.abs(st / n - sp / n),
        }
