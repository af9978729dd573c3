"""GPU parity: the layered fast path of the contour stage (contours verified and emitted in parallel from the
label pass's boundary rows) produces exactly the vertices and integers of the walk, and falls back to the walk
wherever a map is not layered (blobs, noise, touching or missing layers)."""
import numpy as np
import pytest
import torch

from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth

pytestmark = pytest.mark.gpu


def _both(yt, yp, k, dev):
    a, b = torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)
    lp = suite.label_pass(a, b, k, seeds=True, boundaries=True)
    walk = suite.contour_pass(a, b, k, lp.first_pos, return_vertices=True, max_pts=8192)
    fast = suite.contour_pass(a, b, k, lp.first_pos, return_vertices=True, max_pts=8192, boundaries=(lp.bnd_true, lp.bnd_pred))
    return walk, fast


def _assert_same(walk, fast):
    for f in ("n_pts", "flags", "max_sq", "p95_sq"):
        assert torch.equal(getattr(walk, f), getattr(fast, f)), f
    np.testing.assert_allclose(fast.sum_dist.cpu().numpy(), walk.sum_dist.cpu().numpy(), rtol=1e-12)
    n = walk.n_pts.cpu().numpy()
    vw, vf = walk.verts.cpu().numpy().view(np.uint32), fast.verts.cpu().numpy().view(np.uint32)
    ordered = total = 0
    for idx in np.ndindex(n.shape):
        m = n[idx]
        a, b = vw[idx][:m], vf[idx][:m]
        np.testing.assert_array_equal(np.sort(a), np.sort(b), err_msg=str(idx))
        if m > 8:
            total += 1
            ordered += bool(np.all(np.diff((b & 0xffff).astype(np.int64)) >= 0))
    return ordered, total


def test_layered_maps_take_the_fast_path(cuda):
    yt, yp = synth.layered_pair(6, 496, 512, 8, seed=901)
    ordered, total = _assert_same(*_both(yt, yp, 8, cuda))
    assert total == 6 * 8 * 2 and ordered >= 0.9 * total        # emitted left to right, not in walk order


@pytest.mark.parametrize("h,w,k,kw", [(200, 256, 6, dict(noise=0.002)), (64, 512, 4, dict(jitter=3.0)),
                                      (96, 48, 3, dict()), (130, 512, 8, dict(min_gap=1, jitter=2.5)),
                                      (40, 32, 2, dict()), (37, 50, 3, dict()), (45, 33, 3, dict()),
                                      (64, 130, 5, dict(jitter=2.0)), (120, 96, 10, dict()), (200, 64, 16, dict(noise=0.001)),
                                      (2, 64, 2, dict()), (3, 32, 2, dict())])
def test_layered_variants_and_noise(cuda, h, w, k, kw):
    yt, yp = synth.layered_pair(5, h, w, k, seed=77 + h, **kw)
    _assert_same(*_both(yt, yp, k, cuda))


def test_blobs_and_random_maps_fall_back(cuda):
    yt, yp = synth.lesion_pair(4, 160, 160, 4, seed=5, single_blob_interior=False)
    _assert_same(*_both(yt, yp, 4, cuda))
    yt, yp = synth.random_pair(3, 64, 64, 5, seed=6)
    _assert_same(*_both(yt, yp, 5, cuda))


def test_hand_made_edge_cases(cuda):
    h, w, k = 48, 64, 4
    base = np.zeros((h, w), np.uint8)
    base[10:20] = 1
    base[20:30] = 2
    base[30:] = 3
    cases = [base.copy() for _ in range(8)]
    cases[1][5, 40] = 2                      # a pixel of class 2 above its layer: the seed is not on the path
    cases[2][10:20, 17] = 0                  # layer 1 missing in one column
    cases[3][0:10, 30] = 1                   # layer 1 reaches the top border in one column
    cases[4][25, 10] = 1                     # a hole inside layer 2 (not on any path)
    cases[5][19, 0:32] = 2                   # a step of one row at a group boundary
    cases[6][10:14, 33:] = 0                 # a step of four rows
    cases[6][20:27, 50:] = 1                 # and a second layer stepping the other way
    cases[7][30:, :] = 2                     # class 3 absent
    yt = np.stack(cases)
    yp = np.stack(cases[::-1])
    ordered, total = _assert_same(*_both(yt, yp, k, cuda))
    assert ordered > 0


def test_suite_uses_it_and_matches_walk_only_run(cuda, monkeypatch):
    yt, yp = synth.layered_pair(8, 128, 256, 5, seed=33, noise=0.001)
    a, b = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    res = suite.evaluate(a, b, 5)
    lp = suite.label_pass(a, b, 5, seeds=True)
    walk = suite.contour_pass(a, b, 5, lp.first_pos)
    for f in ("n_pts", "flags", "max_sq", "p95_sq"):
        assert torch.equal(getattr(walk, f), getattr(res.contours, f)), f
    assert res.labels.bnd_true is None        # internal boundary rows are not kept unless asked for
