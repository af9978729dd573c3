"""GPU parity: 3-D surface distances (exact separable EDT) against the scipy oracle: bit-exact squared-distance
statistics, hd bit-exact, hd95 / assd within 1e-6 relative."""
import numpy as np
import pytest

from oracle import surface3d_oracle as so
from retinal_oct_image_segmentation_via_deep_learning_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _check(vt, vp, k, cuda, units=None):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    ints = suite.surface_distance_3d(torch.from_numpy(vt).to(cuda), torch.from_numpy(vp).to(cuda), k, units=units)
    m = suite.surface_metrics_3d(ints)
    n_pts = ints["n_pts"].cpu().numpy().view(np.uint32)
    max_sq = ints["max_sq"].cpu().numpy().view(np.uint32)
    sums = ints["sum_dist"].cpu().numpy()
    for c in range(k):
        ref = so.class_metrics(vt, vp, c)
        if ref["sq_pred_to_true"] is None:
            assert n_pts[c, 0] == 0 and n_pts[c, 1] == 0 and np.isnan(m["hausdorff_distance"][c])
            continue
        assert n_pts[c, 1] == len(ref["sq_pred_to_true"]) and n_pts[c, 0] == len(ref["sq_true_to_pred"]), c
        assert max_sq[c, 0] == ref["sq_pred_to_true"].max() and max_sq[c, 1] == ref["sq_true_to_pred"].max(), c
        np.testing.assert_allclose(sums[c, 0], np.sqrt(ref["sq_pred_to_true"].astype(np.float64)).sum(), rtol=1e-12)
        assert m["hausdorff_distance"][c] == ref["hausdorff_distance"]                 # bit-exact
        for name in ("hausdorff_distance_95", "assd"):
            np.testing.assert_allclose(m[name][c], ref[name], rtol=RTOL, atol=0, err_msg=f"{name} class {c}")
    return ints


def test_layered_volume(cuda):
    vt, vp = synth.layered_volume_pair(48, 40, 24, 5, seed=51)
    _check(vt, vp, 5, cuda)


def test_random_blobs_and_missing_class(cuda):
    rng = np.random.default_rng(52)
    vt = (rng.random((20, 33, 17)) < 0.3).astype(np.uint8) + (rng.random((20, 33, 17)) < 0.1).astype(np.uint8)
    vp = (rng.random((20, 33, 17)) < 0.3).astype(np.uint8) + (rng.random((20, 33, 17)) < 0.1).astype(np.uint8)
    vp[vp == 2] = 1                                    # class 2 absent from the prediction -> undefined
    _check(vt, vp, 4, cuda)                            # class 3 absent from both


def test_vectorised_first_pass_shapes(cuda):
    """D2 a multiple of 16: the 16-voxel first pass (borders, thin sheets, runs crossing group boundaries)"""
    rng = np.random.default_rng(54)
    for shape in ((12, 9, 32), (7, 11, 48), (5, 6, 16)):
        vt = (rng.random(shape) < 0.5).astype(np.uint8)
        vp = (rng.random(shape) < 0.5).astype(np.uint8)
        vt[:, :, 10:20] = 1                                # long run along the contiguous axis
        _check(vt, vp, 2, cuda)
    vt, vp = synth.layered_volume_pair(40, 24, 32, 4, seed=55)
    _check(vt, vp, 4, cuda)


def test_far_apart_and_border_touching(cuda):
    vt = np.zeros((40, 30, 12), np.uint8)
    vp = np.zeros((40, 30, 12), np.uint8)
    vt[0:6, 0:5, :] = 1                                # touches three faces of the volume
    vp[30:40, 22:30, 3:9] = 1                          # far corner: large distances take the global-histogram route
    _check(vt, vp, 2, cuda)


def test_unit_ranges_add_up(cuda):
    import torch
    vt, vp = synth.layered_volume_pair(32, 24, 16, 4, seed=53)
    full = _check(vt, vp, 4, cuda)
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    a = suite.surface_distance_3d(torch.from_numpy(vt).to(cuda), torch.from_numpy(vp).to(cuda), 4, units=(0, 3))
    b = suite.surface_distance_3d(torch.from_numpy(vt).to(cuda), torch.from_numpy(vp).to(cuda), 4, units=(3, 8))
    for key in full:
        assert torch.equal(a[key] + b[key], full[key]), key


def test_near_field_path_and_its_overflow(cuda):
    """D2 % 16 == 0 takes the capped windowed transform (exact below (R+1)^2 = 121); a unit with a single query voxel at
    or beyond the cap must be redone by the general kernels, while the other units of the same call keep the fast result."""
    rng = np.random.default_rng(56)
    vt = np.zeros((40, 36, 32), np.uint8)
    vp = np.zeros((40, 36, 32), np.uint8)
    vt[5:20, 4:30, 3:29] = 1                           # class 1: two boxes one or two voxels apart (near)
    vp[6:21, 5:30, 2:28] = 1
    vt[25:38, 2:10, 2:12] = 2                          # class 2: the prediction has an extra island 20 voxels away (far)
    vp[25:38, 2:10, 2:12] = 2
    vp[26:30, 30:34, 24:30] = 2
    vt[22, 20:24, 16] = 3                              # class 3: exactly at the cap: distance 11 along one axis
    vp[22, 20:24, 27] = 3
    _check(vt, vp, 4, cuda)
    # distances 10 (below the cap) and 11 (at the cap) along each axis in turn
    for axis in range(3):
        for gap in (10, 11):
            a = np.zeros((48, 48, 48), np.uint8)
            b = np.zeros((48, 48, 48), np.uint8)
            idx = [slice(20, 23)] * 3
            a[tuple(idx)] = 1
            idx[axis] = slice(20 + 2 + gap, 23 + 2 + gap)
            b[tuple(idx)] = 1
            _check(a, b, 2, cuda)
    # random surfaces with D2 = 16: dense enough that everything is near
    vt = (rng.random((24, 20, 16)) < 0.5).astype(np.uint8)
    vp = (rng.random((24, 20, 16)) < 0.5).astype(np.uint8)
    _check(vt, vp, 2, cuda)
