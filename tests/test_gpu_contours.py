"""GPU parity: contour [0] tracing, exact squared distances and hausdorff / hd95 / assd vs the oracle."""
import numpy as np
import pytest

from oracle import labelmap_oracle as lo
from oracle import contours_oracle as co
from retinal_oct_image_segmentation_via_deep_learning_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-6      # north_star tolerance for derived floating-point values


def _contours(yt, yp, k, cuda, **kw):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    out = suite.contour_pass(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda), k,
                             return_vertices=True, return_sq=True, **kw)
    torch.cuda.synchronize()
    return out


def _check_against_oracle(yt, yp, k, cuda, max_pts=2048):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import derive
    out = _contours(yt, yp, k, cuda, max_pts=max_pts)
    n_pts = out.n_pts.cpu().numpy().view(np.uint32)
    verts = out.verts.cpu().numpy().view(np.uint32)
    sq = out.sq.cpu().numpy().view(np.uint32)
    m = derive.contour_metrics(n_pts, out.max_sq.cpu().numpy().view(np.uint32),
                               out.p95_sq.cpu().numpy().view(np.uint32), out.sum_dist.cpu().numpy())
    for i in range(yt.shape[0]):
        for c in range(k):
            im = lo.contour_intermediates(yt[i] == c, yp[i] == c)
            if im is None:
                assert n_pts[i, c, 0] == 0 or n_pts[i, c, 1] == 0
                assert np.isnan(m["hausdorff_distance"][i, c])
                continue
            for mm, key in ((0, "verts_true"), (1, "verts_pred")):
                got = verts[i, c, mm, :n_pts[i, c, mm]]
                got = np.stack([got >> 16, got & 0xffff], 1).astype(np.int64)
                ref = im[key]
                assert len(got) == len(ref), (i, c, mm)
                # same multiset of vertices (order inside the array does not affect any metric)
                assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, ref.tolist())), (i, c, mm)
            # exact squared distances, as multisets per direction
            np.testing.assert_array_equal(np.sort(sq[i, c, 0, :n_pts[i, c, 1]]), np.sort(im["sq_pred_to_true"]))
            np.testing.assert_array_equal(np.sort(sq[i, c, 1, :n_pts[i, c, 0]]), np.sort(im["sq_true_to_pred"]))
            ref_m = lo.contour_metrics_from_sq(im["sq_pred_to_true"], im["sq_true_to_pred"])
            assert m["hausdorff_distance"][i, c] == ref_m["hausdorff_distance"]          # bit-exact
            for name in ("hausdorff_distance_95", "assd"):
                np.testing.assert_allclose(m[name][i, c], ref_m[name], rtol=RTOL, atol=0, err_msg=f"{name} {i} {c}")


def test_golden_contours(cuda, golden_dir):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import derive
    g = np.load(f"{golden_dir}/contours_golden.npz")
    for name in g["names"]:
        a, b = g[f"{name}/mask_true"], g[f"{name}/mask_pred"]
        out = _contours(a[None], b[None], 2, cuda)
        n_pts = out.n_pts.cpu().numpy().view(np.uint32)[0, 1]
        sq = out.sq.cpu().numpy().view(np.uint32)[0, 1]
        assert n_pts[0] == len(g[f"{name}/verts_true"]) and n_pts[1] == len(g[f"{name}/verts_pred"]), name
        np.testing.assert_array_equal(np.sort(sq[0, :n_pts[1]]), np.sort(g[f"{name}/sq_pred_to_true"]), err_msg=name)
        np.testing.assert_array_equal(np.sort(sq[1, :n_pts[0]]), np.sort(g[f"{name}/sq_true_to_pred"]), err_msg=name)
        m = derive.contour_metrics(n_pts, out.max_sq.cpu().numpy().view(np.uint32)[0, 1],
                                   out.p95_sq.cpu().numpy().view(np.uint32)[0, 1], out.sum_dist.cpu().numpy()[0, 1])
        ref = g[f"{name}/metrics"]
        assert float(m["hausdorff_distance"]) == ref[0], name
        np.testing.assert_allclose(float(m["hausdorff_distance_95"]), ref[1], rtol=RTOL, err_msg=name)
        np.testing.assert_allclose(float(m["assd"]), ref[2], rtol=RTOL, err_msg=name)


def test_random_binary_masks_all_topologies(cuda):
    """Small random masks: saddles, holes, border-touching (open) and closed contours, tiny islands."""
    rng = np.random.default_rng(41)
    for _ in range(40):
        h, w = rng.integers(2, 20, size=2)
        n = 6
        a = (rng.random((n, h, w)) < rng.uniform(0.2, 0.8)).astype(np.uint8)
        b = (rng.random((n, h, w)) < rng.uniform(0.2, 0.8)).astype(np.uint8)
        a[0] = 0                      # class 1 absent -> no contour for either class
        _check_against_oracle(a, b, 2, cuda)


def test_layered_multiclass(cuda):
    yt, yp = synth.layered_pair(3, 120, 160, 6, seed=42)
    _check_against_oracle(yt, yp, 6, cuda)
    yt, yp = synth.layered_pair(2, 96, 128, 8, seed=43, noise=0.01, min_gap=1)
    _check_against_oracle(yt, yp, 8, cuda)


def test_lesions(cuda):
    yt, yp = synth.lesion_pair(3, 128, 128, 4, seed=44)
    _check_against_oracle(yt, yp, 4, cuda)
    yt, yp = synth.lesion_pair(3, 128, 128, 4, seed=45, single_blob_interior=False)
    _check_against_oracle(yt, yp, 4, cuda)


def test_full_size_bscan_and_overflow_retry(cuda):
    """One 496x512 8-class B-scan through evaluate(); max_pts small enough that the retry path runs."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    yt, yp = synth.layered_pair(2, 496, 512, 8, seed=46)
    res = suite.evaluate(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda), 8, max_pts=256)
    m = res.metrics()
    for c in (0, 3, 7):
        im = lo.contour_intermediates(yt[1] == c, yp[1] == c)
        ref = lo.contour_metrics_from_sq(im["sq_pred_to_true"], im["sq_true_to_pred"])
        assert m["hausdorff_distance"][1, c] == ref["hausdorff_distance"]
        np.testing.assert_allclose(m["hausdorff_distance_95"][1, c], ref["hausdorff_distance_95"], rtol=RTOL)
        np.testing.assert_allclose(m["assd"][1, c], ref["assd"], rtol=RTOL)


def test_sq_distances_match_edt(cuda):
    """Independent check of D2: scipy's exact EDT on the doubled lattice."""
    from scipy.ndimage import distance_transform_edt
    yt, yp = synth.lesion_pair(1, 96, 96, 3, seed=47)
    out = _contours(yt, yp, 3, cuda)
    n_pts = out.n_pts.cpu().numpy().view(np.uint32)[0]
    verts = out.verts.cpu().numpy().view(np.uint32)[0]
    sq = out.sq.cpu().numpy().view(np.uint32)[0]
    for c in range(1, 3):
        vt = verts[c, 0, :n_pts[c, 0]]
        vp = verts[c, 1, :n_pts[c, 1]]
        grid = np.ones((2 * 96 - 1, 2 * 96 - 1), bool)
        grid[vt >> 16, vt & 0xffff] = False
        edt2 = np.rint(distance_transform_edt(grid) ** 2).astype(np.int64)
        np.testing.assert_array_equal(sq[c, 0, :n_pts[c, 1]], edt2[vp >> 16, vp & 0xffff])
