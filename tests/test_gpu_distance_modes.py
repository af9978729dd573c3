"""GPU parity: the distance-search strategies (column-sorted, tiled cooperative boxes, per-lane boxes,
brute force) return identical squared distances -- pruning and tiling never change a result."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys, numpy as np, torch
sys.path.insert(0, %r)
from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth
dev = torch.device("cuda", 0)
out = {}
for name, (yt, yp), k in [("layered", synth.layered_pair(3, 200, 256, 6, seed=71, noise=0.002), 6),
                          ("lesion", synth.lesion_pair(3, 160, 160, 4, seed=72, single_blob_interior=False), 4)]:
    ct = suite.contour_pass(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev), k, return_sq=True)
    n = ct.n_pts.cpu().numpy().view(np.uint32)
    sq = ct.sq.cpu().numpy().view(np.uint32)
    tot = 0
    for i in range(sq.shape[0]):
        for c in range(k):
            tot += int(sq[i, c, 0, :n[i, c, 1]].astype(np.int64).sum()) * 7 + int(sq[i, c, 1, :n[i, c, 0]].astype(np.int64).sum())
    # without return_sq the search kernel counts the distances instead of storing them: same integers, sums
    # equal up to the order of the float64 additions
    ct2 = suite.contour_pass(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev), k)
    assert torch.equal(ct2.max_sq, ct.max_sq) and torch.equal(ct2.p95_sq, ct.p95_sq), name
    np.testing.assert_allclose(ct2.sum_dist.cpu().numpy(), ct.sum_dist.cpu().numpy(), rtol=1e-12)
    out[name] = [tot, ct.max_sq.cpu().numpy().view(np.uint32).tolist(), ct.p95_sq.cpu().numpy().view(np.uint32).tolist()]
print(json.dumps(out))
""" % ROOT


def _run(mode, tile=None):
    env = dict(os.environ, OCTM_DISTANCE_MODE=mode)
    if tile is not None:
        env["OCTM_DIST_TILE"] = str(tile)
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_all_modes_agree(cuda):
    ref = _run("brute")
    assert _run("column") == ref                # default: column-sorted cooperative search
    assert _run("tiled") == ref
    assert _run("tiled", tile=64) == ref        # contours span many source tiles
    assert _run("lane") == ref
