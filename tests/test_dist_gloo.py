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
