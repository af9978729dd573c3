"""Multi-GPU sharding: one process per GPU, B-scans partitioned contiguously, ONE small all-reduce.

Every reference function is a pure function of one (y_true, y_pred) pair, so items shard with no
data-path collective.  What crosses NVLink is a packed float64 vector of a few
hundred bytes per rank: summed confusion counts, column-scan sums, per-class sums of the contour
metrics and their valid-item counts.  Integer partials ride in float64 exactly (each must stay
below 2**53, which is checked), so the totals are identical for every world size; the genuine
floating-point sums differ by rounding only (<= 1e-12 relative).  ``want_max=True`` adds a second
(MAX) all-reduce for the dataset-level Hausdorff maximum.
"""
from __future__ import annotations

import numpy as np

from . import derive

_EXACT_LIMIT = float(2 ** 53)


def shard_range(n_items, rank, world):
    """Contiguous [start, stop) of the items rank `rank` scores; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_partials(ints, metrics, num_classes):
    """Pack this rank's sums into one float64 vector (layout mirrored by ``unpack``)."""
    k = num_classes
    parts = [np.asarray([ints["confusion"].shape[0]], np.float64),
             ints["confusion"].astype(np.int64).sum(0).reshape(-1).astype(np.float64)]
    if "thickness_absdiff" in ints:
        parts += [ints["thickness_absdiff"].sum(0).astype(np.float64),
                  ints["boundary_sq"].sum(0).astype(np.float64), ints["boundary_abs"].sum(0).astype(np.float64)]
    else:
        parts += [np.zeros(k), np.zeros(k - 1), np.zeros(k - 1)]
    if "contour_valid" in metrics:
        valid = metrics["contour_valid"]
        parts.append(valid.sum(0).astype(np.float64))
        for name in ("hausdorff_distance", "hausdorff_distance_95", "assd"):
            parts.append(np.where(valid, metrics[name], 0.0).sum(0))
    else:
        parts += [np.zeros(k)] * 4
    vec = np.concatenate(parts)
    n_exact = 1 + k * k + k + 2 * (k - 1) + k
    if np.any(np.abs(vec[:n_exact]) >= _EXACT_LIMIT):
        raise OverflowError("an integer partial exceeds 2**53 and would not be exact in the float64 all-reduce")
    return vec


def unpack(vec, num_classes, width):
    k = num_classes
    o = 0

    def take(m):
        nonlocal o
        out = vec[o:o + m]
        o += m
        return out

    n_items = int(round(take(1)[0]))
    cm = np.rint(take(k * k)).astype(np.int64).reshape(k, k)
    thick = np.rint(take(k)).astype(np.int64)
    bsq, bab = np.rint(take(k - 1)).astype(np.int64), np.rint(take(k - 1)).astype(np.int64)
    nvalid = np.rint(take(k)).astype(np.int64)
    s_hd, s_hd95, s_assd = take(k), take(k), take(k)
    out = {"n_items": n_items, "confusion": cm, "contour_items": nvalid}
    out.update(derive.count_metrics(*derive.class_counts(cm)))          # pooled (micro) ratios per class
    denom = max(n_items, 1) * width
    out["thickness_difference"] = thick.astype(np.float64) / denom       # mean over all columns of all items
    out["boundary_mse"] = bsq.astype(np.float64) / denom
    out["boundary_rmse"] = np.sqrt(out["boundary_mse"])
    out["boundary_mad"] = bab.astype(np.float64) / denom
    with np.errstate(invalid="ignore", divide="ignore"):
        out["hausdorff_distance_mean"] = s_hd / nvalid
        out["hausdorff_distance_95_mean"] = s_hd95 / nvalid
        out["assd_mean"] = s_assd / nvalid
    return out


def all_reduce_sum(vec, world, device=None, group=None):
    """float64 SUM all-reduce of a small numpy vector (NCCL on `device`, gloo when device is None/cpu)."""
    if world == 1:
        return vec
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(vec))
    if device is not None and str(device) != "cpu":
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def base_len(num_classes):
    k = num_classes
    return 1 + k * k + k + 2 * (k - 1) + 4 * k


def dataset_totals(res, world, device=None, group=None, want_max=False):
    """Dataset-level numbers over all ranks' shards from one SuiteResult per rank.

    The per-rank partial sums come from the device-side totals kernel.  With world > 1 they are
    summed across ranks by ONE float64 all-reduce issued directly on the device vector (stream
    ordered after the kernels, no host round trip before the collective); a single small D2H then
    brings the reduced sums, this rank's maxima and its overflow flags to the host."""
    k, w = res.labels.num_classes, res.labels.width
    nb = base_len(k)
    if world > 1 and res.totals is not None and res.totals.is_cuda and not res._final and res._totals_host is None:
        import torch.distributed as dist
        local = res.totals
        reduced = local.clone()
        dist.all_reduce(reduced[:nb], op=dist.ReduceOp.SUM, group=group)
        if want_max:
            dist.all_reduce(reduced[nb:nb + k], op=dist.ReduceOp.MAX, group=group)
        both = __import__("torch").stack([local, reduced]).cpu().numpy()      # one D2H
        res._totals_host = None
        vec_local, vec = both[0], both[1]
        if int(vec_local[-1]) & 12 and res.contours is not None:             # a contour overflowed on this rank:
            res.totals_host()                                                 # redo it, then reduce again (rare)
            return dataset_totals(res, world, device, group, want_max)
        res._totals_host, res._final, res._inputs = vec_local, True, None
        if np.any(np.abs(vec_local[:nb - 3 * k]) >= _EXACT_LIMIT):
            raise OverflowError("an integer partial exceeds 2**53 and would not be exact in the float64 all-reduce")
        out = unpack(vec[:nb], k, w)
        if want_max:
            out["hausdorff_distance_max"] = np.where(vec[nb:nb + k] < 0, np.nan, vec[nb:nb + k])
        return out
    vec = res.totals_host()
    base = vec[:nb].copy()
    if np.any(np.abs(base[:nb - 3 * k]) >= _EXACT_LIMIT):
        raise OverflowError("an integer partial exceeds 2**53 and would not be exact in the float64 all-reduce")
    if device is None and world > 1:
        device = res.labels.counts.device
    out = unpack(all_reduce_sum(base, world, device, group), k, w)
    if want_max:
        local = vec[nb:nb + k].copy()
        if world > 1:
            import torch
            import torch.distributed as dist
            t = torch.from_numpy(local).to(device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            local = t.cpu().numpy()
        out["hausdorff_distance_max"] = np.where(local < 0, np.nan, local)
    return out


class PendingTotals:
    """Dataset totals whose device work (totals kernel, all-reduce) is enqueued but not yet read back."""

    def __init__(self, res, world, group, want_max, local, reduced):
        self.res, self.world, self.group, self.want_max = res, world, group, want_max
        self.local, self.reduced = local, reduced

    def result(self):
        """One small D2H; redoes (and re-reduces) the rare items whose contour overflowed."""
        res, k, w = self.res, self.res.labels.num_classes, self.res.labels.width
        nb = base_len(k)
        if self.reduced is None:
            return dataset_totals(res, self.world, group=self.group, want_max=self.want_max)
        import torch
        both = torch.stack([self.local, self.reduced]).cpu().numpy()
        vec_local, vec = both[0], both[1]
        if int(vec_local[-1]) & 12 and res.contours is not None and not res._final:
            res.totals_host()
            return dataset_totals(res, self.world, group=self.group, want_max=self.want_max)
        if res._totals_host is None:
            res._totals_host, res._final, res._inputs = vec_local, True, None
        if np.any(np.abs(vec_local[:nb - 3 * k]) >= _EXACT_LIMIT):
            raise OverflowError("an integer partial exceeds 2**53 and would not be exact in the float64 all-reduce")
        out = unpack(vec[:nb], k, w)
        if self.want_max:
            out["hausdorff_distance_max"] = np.where(vec[nb:nb + k] < 0, np.nan, vec[nb:nb + k])
        return out


def dataset_totals_async(res, world, group=None, want_max=False):
    """Enqueue the cross-rank reduction of ``res.totals`` on the current stream and return a
    ``PendingTotals``; nothing is copied to the host until ``.result()``.  Lets a caller keep several
    evaluations in flight (the benchmark's device-timed region does)."""
    if res.totals is None or not res.totals.is_cuda:
        return PendingTotals(res, world, group, want_max, None, None)
    k = res.labels.num_classes
    nb = base_len(k)
    local = res.totals
    reduced = local.clone()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(reduced[:nb], op=dist.ReduceOp.SUM, group=group)
        if want_max:
            dist.all_reduce(reduced[nb:nb + k], op=dist.ReduceOp.MAX, group=group)
    return PendingTotals(res, world, group, want_max, local, reduced)


def surface_distance_3d_sharded(vol_true, vol_pred, num_classes, rank, world, group=None):
    """BASELINE config 5 on several GPUs: the volume pair is replicated (it is small), the ``2 * K``
    independent (class, direction) distance transforms are split into contiguous unit ranges per rank,
    and the per-unit integers / sums are combined by ONE small all-reduce (every unit is written by exactly
    one rank, the others contribute zeros, so the sum is exact).  Returns the same dict as
    ``suite.surface_distance_3d`` with identical contents on every rank."""
    import torch
    from . import suite
    ub, ue = shard_range(2 * num_classes, rank, world)
    ints = suite.surface_distance_3d(vol_true, vol_pred, num_classes, units=(ub, ue))
    if world > 1:
        import torch.distributed as dist
        packed = torch.cat([ints["n_pts"].to(torch.float64).reshape(-1), ints["max_sq"].to(torch.float64).reshape(-1),
                            ints["p95_sq"].to(torch.float64).reshape(-1), ints["sum_dist"].reshape(-1)])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        k = num_classes
        o = 0
        for key, shape, dt in (("n_pts", (k, 2), torch.int32), ("max_sq", (k, 2), torch.int32),
                               ("p95_sq", (k, 2, 2), torch.int32), ("sum_dist", (k, 2), torch.float64)):
            m = int(np.prod(shape))
            v = packed[o:o + m].reshape(shape)
            ints[key] = v.to(torch.int64).to(dt) if dt != torch.float64 else v.clone()
            o += m
    return ints
