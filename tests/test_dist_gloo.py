"""CPU tier: the N>1 path (shard -> local partial sums -> one all-reduce) with world_size 2 on gloo."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from oracle import labelmap_oracle as lo
from retinal_oct_image_segmentation_via_deep_learning_b200 import dist as odist
from retinal_oct_image_segmentation_via_deep_learning_b200 import derive, synth

K, H, W, N = 5, 24, 32, 7


def _fake_rank_outputs(yt, yp):
    """What a rank's SuiteResult.integers()/metrics() would hold, built from the oracle (no GPU here)."""
    per = [lo.score_bscan_fast(yt[i], yp[i], K) for i in range(len(yt))]
    ints = {"confusion": np.stack([p["confusion"] for p in per]),
            "thickness_absdiff": np.stack([p["thickness_absdiff"] for p in per]),
            "boundary_sq": np.stack([p["boundary_sq"] for p in per]),
            "boundary_abs": np.stack([p["boundary_abs"] for p in per])}
    rng = np.random.default_rng(len(yt))
    valid = rng.random((len(yt), K)) < 0.8
    metrics = {"contour_valid": valid, "hausdorff_distance": rng.random((len(yt), K)) * 9,
               "hausdorff_distance_95": rng.random((len(yt), K)) * 7, "assd": rng.random((len(yt), K))}
    return ints, metrics


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    yt, yp = synth.layered_pair(N, H, W, K, seed=77, noise=0.05)
    s, e = odist.shard_range(N, rank, world)
    ints, metrics = _fake_rank_outputs(yt[s:e], yp[s:e])
    vec = odist.all_reduce_sum(odist.local_partials(ints, metrics, K), world, device="cpu")
    q.put((rank, vec))
    torch.distributed.destroy_process_group()


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 100_000):
        for world in (1, 2, 4, 8):
            spans = [odist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_equals_single_process_totals():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1])                     # every rank holds the same totals
    yt, yp = synth.layered_pair(N, H, W, K, seed=77, noise=0.05)
    tot = odist.unpack(got[0], K, W)
    cm = sum(lo.confusion_matrix(yt[i], yp[i], K).astype(np.int64) for i in range(N))
    assert tot["n_items"] == N
    assert np.array_equal(tot["confusion"], cm)                # integers identical for every world size
    pooled = derive.count_metrics(*derive.class_counts(cm))
    assert np.array_equal(tot["dice_coefficient"], pooled["dice_coefficient"])
    per = [lo.score_bscan_fast(yt[i], yp[i], K) for i in range(N)]
    assert np.array_equal(tot["boundary_mse"], sum(p["boundary_sq"] for p in per) / (N * W))
    # single-process reduction of the same shards gives the same vector (float sums to rounding)
    vecs = []
    for r in range(2):
        s, e = odist.shard_range(N, r, 2)
        vecs.append(odist.local_partials(*_fake_rank_outputs(yt[s:e], yp[s:e]), K))
    np.testing.assert_allclose(got[0], vecs[0] + vecs[1], rtol=1e-12)


# ---------------------------------------------------------------------------------- collective decisions
class _FakeLabels:
    num_classes, width = K, W


class _FakeResult:
    """What dist._finish needs of a SuiteResult, with CPU totals: rank `over_rank` starts with a contour-overflow
    flag (its sums are stale until settle_overflow() redoes the item), the other rank is clean."""

    def __init__(self, rank, over_rank, bad_rank):
        self.labels, self.contours, self.validate = _FakeLabels(), object(), True
        self._totals_host, self._final, self._inputs = None, False, None
        self.settled = 0
        nb = odist.base_len(K)
        v = np.zeros(nb + K + 1)
        v[0] = 3                                  # items on this rank
        v[1] = 100 + rank                         # a confusion count
        v[nb:nb + K] = -1.0                       # no Hausdorff maxima
        if rank == over_rank:
            v[nb - 2] = 1                         # n_overflow_items
            v[-1] = 4                             # OR of contour flags: CF_TRUE_OVERFLOW
            v[1] = -1000                          # stale partial that the redo replaces
        if rank == bad_rank:
            v[nb - 1] = 2                         # n_bad_label_items
        self.totals = torch.from_numpy(v)
        self.rank = rank

    def settle_overflow(self, vec):
        self.settled += 1
        if not int(vec[-1]) & 12:
            return False
        v = self.totals.numpy().copy()
        nb = odist.base_len(K)
        v[nb - 2], v[-1], v[1] = 0, 0, 100 + self.rank
        self.totals = torch.from_numpy(v)
        return True


def _collective_worker(rank, world, port, q, over_rank, bad_rank, use_async):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    calls = {"n": 0}
    real = torch.distributed.all_reduce

    def counting(*a, **kw):
        calls["n"] += 1
        return real(*a, **kw)
    torch.distributed.all_reduce = counting
    res = _FakeResult(rank, over_rank, bad_rank)
    try:
        if use_async:
            out = odist.dataset_totals_async(res, world).result()
        else:
            out = odist.dataset_totals(res, world)
        q.put((rank, "ok", calls["n"], res.settled, int(out["confusion"][0, 0]), out["n_items"]))
    except ValueError as e:
        q.put((rank, "ValueError", calls["n"], res.settled, str(e), 0))
    torch.distributed.barrier()                   # a rank that left a collective early would hang or mismatch here
    torch.distributed.destroy_process_group()


def _run_collective(over_rank, bad_rank, use_async):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_collective_worker, args=(r, 2, port, q, over_rank, bad_rank, use_async)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def test_overflow_on_one_rank_is_settled_collectively():
    """only rank 1 overflows; BOTH ranks must join the second reduction and end with the totals
    that include rank 1's redone items (100 + 101), not its stale partial."""
    for use_async in (False, True):
        got = _run_collective(over_rank=1, bad_rank=-1, use_async=use_async)
        assert [g[1] for g in got] == ["ok", "ok"]
        assert [g[2] for g in got] == [2, 2]                  # two all-reduces on every rank
        assert [g[4] for g in got] == [201, 201] and [g[5] for g in got] == [6, 6]
    got = _run_collective(over_rank=-1, bad_rank=-1, use_async=True)
    assert [g[2] for g in got] == [1, 1] and [g[4] for g in got] == [201, 201]     # the usual case: one collective


def test_invalid_label_on_one_rank_raises_on_every_rank():
    got = _run_collective(over_rank=-1, bad_rank=0, use_async=False)
    assert [g[1] for g in got] == ["ValueError", "ValueError"]
    assert all("label >= num_classes" in g[4] for g in got)
