"""GPU: size-independent properties of the suite at BASELINE.json's full B-scan size (cfg4: 496 x 512, 8 classes)
on a batch far larger than the oracle could score: conservation, symmetry, identity and batch-split invariance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, H, W, K = 1536, 496, 512, 8          # 0.78 GB of labels: > 6x the 126 MB L2


@pytest.fixture(scope="module")
def batch(cuda):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import synth
    return synth.layered_pair_device(N, H, W, K, seed=4004, device=cuda, noise=0.002)


def test_conservation_and_totals(cuda, batch):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import dist, suite
    yt, yp = batch
    res = suite.evaluate(yt, yp, K)
    ints = res.integers()
    cm = ints["confusion"].astype(np.int64)
    assert (cm.sum(axis=(1, 2)) == H * W).all()                       # every pixel lands in exactly one bin
    # row / column sums are the class areas of each map (independent torch reductions)
    import torch
    area_t = torch.stack([(yt == c).sum(dim=(1, 2)) for c in range(K)], 1).cpu().numpy()
    area_p = torch.stack([(yp == c).sum(dim=(1, 2)) for c in range(K)], 1).cpu().numpy()
    assert np.array_equal(cm.sum(2), area_t) and np.array_equal(cm.sum(1), area_p)
    # boundaries are cumulative thicknesses: sum_k |d_k| >= |thickness diff| per class is implied; check the exact link
    assert (ints["boundary_abs"] >= 0).all() and (ints["boundary_sq"] >= ints["boundary_abs"]).all()
    tot = dist.dataset_totals(res, 1)
    assert tot["n_items"] == N and np.array_equal(tot["confusion"], cm.sum(0))


def test_identity_pair_scores_perfectly(cuda, batch):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    yt, _ = batch
    sub = yt[:256]
    res = suite.evaluate(sub, sub, K)
    ints, m = res.integers(), res.metrics()
    cm = ints["confusion"].astype(np.int64)
    assert (cm[:, ~np.eye(K, dtype=bool)] == 0).all()
    assert (ints["thickness_absdiff"] == 0).all() and (ints["boundary_sq"] == 0).all()
    assert (ints["contour_max_sq"] == 0).all() and (ints["contour_sum_dist"] == 0).all()
    assert np.array_equal(ints["contour_n_pts"][..., 0], ints["contour_n_pts"][..., 1])
    assert (m["hausdorff_distance"] == 0).all() and (m["accuracy"] == 1.0).all()
    tp = np.diagonal(cm, axis1=1, axis2=2).astype(np.float64)
    assert np.array_equal(m["dice_coefficient"], (2.0 * tp) / (tp + tp + 1e-7))       # the reference's expression


def test_swapping_arguments_mirrors_the_results(cuda, batch):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    yt, yp = batch
    a = suite.evaluate(yt[:256], yp[:256], K).integers()
    b = suite.evaluate(yp[:256], yt[:256], K).integers()
    assert np.array_equal(a["confusion"], b["confusion"].transpose(0, 2, 1))
    assert np.array_equal(a["thickness_absdiff"], b["thickness_absdiff"])
    assert np.array_equal(a["boundary_sq"], b["boundary_sq"])
    assert np.array_equal(a["contour_n_pts"], b["contour_n_pts"][..., ::-1])
    assert np.array_equal(a["contour_max_sq"], b["contour_max_sq"][..., ::-1])          # d1 <-> d2
    assert np.array_equal(a["contour_p95_sq"], b["contour_p95_sq"][:, :, ::-1, :])
    assert np.array_equal(a["first_pos"], b["first_pos"][:, ::-1, :])


def test_batch_split_invariance(cuda, batch):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    yt, yp = batch
    whole = suite.evaluate(yt, yp, K).integers()
    cut = 517                                                          # not a multiple of anything in the kernels
    parts = [suite.evaluate(yt[:cut], yp[:cut], K).integers(), suite.evaluate(yt[cut:], yp[cut:], K).integers()]
    for key in whole:
        joined = np.concatenate([p[key] for p in parts])
        assert np.array_equal(whole[key], joined), key
