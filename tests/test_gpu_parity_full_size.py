"""GPU parity at the BASELINE shapes, through the production path (``suite.evaluate``) and through the C ABI entry
points SURVEY 8b names: every class of full-size cfg4 B-scans (clean, noisy, thin layers), cfg3 lesion slices at
512x512 (both variants), a cfg5-like 256x256x64 volume against scipy, plus input validation (labels >= K)."""
import ctypes

import numpy as np
import pytest

from oracle import labelmap_oracle as lo
from oracle import surface3d_oracle as so
from retinal_oct_image_segmentation_via_deep_learning_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _oracle_contour_ints(mt, mp):
    """n_pts, max_sq, p95_sq ([lo], [lo+1] as numpy's linear percentile picks them), sum_dist per direction."""
    im = lo.contour_intermediates(mt, mp)
    if im is None:
        return None
    out = {"n_pts": (len(im["verts_true"]), len(im["verts_pred"])), "im": im}
    for d, key in ((0, "sq_pred_to_true"), (1, "sq_true_to_pred")):
        s = np.sort(im[key])
        pos = (len(s) - 1) * 0.95
        lo_i = int(np.floor(pos))
        out[d] = (int(s[-1]), int(s[lo_i]), int(s[min(lo_i + 1, len(s) - 1)]), np.sqrt(s.astype(np.float64) / 4.0).sum())
    return out


def _check_suite_vs_oracle(yt, yp, k, cuda, **kw):
    """evaluate() -- whatever contour path it picks -- against the oracle: integers exact, floats 1e-6."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    res = suite.evaluate(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda), k, boundaries=True, **kw)
    ints, m = res.integers(), res.metrics()
    n_contours = 0
    for i in range(len(yt)):
        fast = lo.score_bscan_fast(yt[i], yp[i], k)
        for key in ("confusion", "thickness_absdiff", "boundary_sq", "boundary_abs", "boundary_true", "boundary_pred"):
            assert np.array_equal(ints[key][i], fast[key]), (i, key)
        for c in range(k):
            ref = _oracle_contour_ints(yt[i] == c, yp[i] == c)
            if ref is None:
                assert ints["contour_n_pts"][i, c, 0] == 0 or ints["contour_n_pts"][i, c, 1] == 0, (i, c)
                assert np.isnan(m["hausdorff_distance"][i, c])
                continue
            n_contours += 1
            assert tuple(ints["contour_n_pts"][i, c]) == ref["n_pts"], (i, c)
            for d in (0, 1):
                mx, vlo, vhi, sm = ref[d]
                assert ints["contour_max_sq"][i, c, d] == mx, (i, c, d)
                assert tuple(ints["contour_p95_sq"][i, c, d]) == (vlo, vhi), (i, c, d)
                np.testing.assert_allclose(ints["contour_sum_dist"][i, c, d], sm, rtol=1e-12)
            rm = lo.contour_metrics_from_sq(ref["im"]["sq_pred_to_true"], ref["im"]["sq_true_to_pred"])
            assert m["hausdorff_distance"][i, c] == rm["hausdorff_distance"]
            np.testing.assert_allclose(m["hausdorff_distance_95"][i, c], rm["hausdorff_distance_95"], rtol=RTOL)
            np.testing.assert_allclose(m["assd"][i, c], rm["assd"], rtol=RTOL)
    return n_contours


def _check_vertices_with_boundaries(yt, yp, k, cuda, max_pts=2048):
    """The layered emit path fed the label pass's boundary rows, vertex multisets and D2 multisets vs the oracle."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    t, p = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    lp = suite.label_pass(t, p, k, seeds=True, boundaries=True)
    out = suite.contour_pass(t, p, k, lp.first_pos, boundaries=(lp.bnd_true, lp.bnd_pred), return_vertices=True,
                             return_sq=True, max_pts=max_pts)
    n_pts = out.n_pts.cpu().numpy().view(np.uint32)
    verts = out.verts.cpu().numpy().view(np.uint32)
    sq = out.sq.cpu().numpy().view(np.uint32)
    for i in range(len(yt)):
        for c in range(k):
            im = lo.contour_intermediates(yt[i] == c, yp[i] == c)
            if im is None:
                assert n_pts[i, c, 0] == 0 or n_pts[i, c, 1] == 0
                continue
            for mm, key in ((0, "verts_true"), (1, "verts_pred")):
                got = verts[i, c, mm, :n_pts[i, c, mm]]
                got = np.stack([got >> 16, got & 0xffff], 1).astype(np.int64)
                assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, im[key].tolist())), (i, c, mm)
            np.testing.assert_array_equal(np.sort(sq[i, c, 0, :n_pts[i, c, 1]]), np.sort(im["sq_pred_to_true"]))
            np.testing.assert_array_equal(np.sort(sq[i, c, 1, :n_pts[i, c, 0]]), np.sort(im["sq_true_to_pred"]))


def test_cfg4_full_size_all_classes_clean(cuda):
    yt, yp = synth.layered_pair(3, 496, 512, 8, seed=4004)
    assert _check_suite_vs_oracle(yt, yp, 8, cuda) == 3 * 8
    _check_vertices_with_boundaries(yt, yp, 8, cuda)


def test_cfg4_full_size_all_classes_noisy(cuda):
    """stray pixels: most contours [0] become tiny blobs around the raster-first stray pixel, the rest are walked"""
    for noise, seed in ((2e-5, 61), (2e-3, 62), (0.02, 63)):
        yt, yp = synth.layered_pair(2, 496, 512, 8, seed=seed, noise=noise)
        assert _check_suite_vs_oracle(yt, yp, 8, cuda) == 2 * 8
        _check_vertices_with_boundaries(yt, yp, 8, cuda)


def test_cfg4_full_size_thin_and_touching_layers(cuda):
    yt, yp = synth.layered_pair(2, 496, 512, 8, seed=64, min_gap=1)
    _check_suite_vs_oracle(yt, yp, 8, cuda)
    _check_vertices_with_boundaries(yt, yp, 8, cuda)
    yt, yp = synth.layered_pair(2, 496, 512, 8, seed=65, min_gap=0, jitter=3.0)       # layers may vanish in places
    _check_suite_vs_oracle(yt, yp, 8, cuda)
    _check_vertices_with_boundaries(yt, yp, 8, cuda, max_pts=8192)


def test_layers_and_blobs_far_apart(cuda):
    """Squared distances beyond what the 16-bit counters hold (>= 2048): layers displaced by tens of pixels (table x
    table), a blob far above its layer (table x short list), two far blobs (short x short): the radix select over
    recomputed distances in the fused kernel's second pass."""
    k, h, w = 6, 496, 512
    yt, _ = synth.layered_pair(4, h, w, k, seed=71)
    yp = yt.copy()
    yp[0] = np.roll(yt[0], 37, axis=0)                       # every boundary 37 rows lower (rows wrap into class k-1 ... 0)
    yp[0, :37] = 0
    yp[1] = np.roll(yt[1], -29, axis=0)
    yp[1, -29:] = k - 1
    yp[2, 3, 400] = 4                                        # one stray pixel of class 4 far above its layer: contour [0] = that pixel
    yp[3, 5, 17] = 3
    yt[3, 480, 500] = 1                                      # ... and a stray pixel below in y_true as well (not a seed: the layer comes first)
    n = _check_suite_vs_oracle(yt, yp, k, cuda)
    assert n == 4 * k
    yt2 = np.zeros((2, 64, 128), np.uint8)                   # two single-pixel blobs 100+ columns apart
    yp2 = np.zeros((2, 64, 128), np.uint8)
    yt2[:, 5, 3] = 1
    yp2[:, 60, 120] = 1
    yt2[1, 30:34, 60:70] = 1
    _check_suite_vs_oracle(yt2, yp2, 2, cuda)


def test_cfg4_uniform_random_labels(cuda):
    """SURVEY 8d adversarial variant: every pixel a random class (worst case for the histogram and the walk)."""
    yt, yp = synth.random_pair(1, 496, 512, 8, seed=66)
    _check_suite_vs_oracle(yt, yp, 8, cuda)


def test_cfg1_and_cfg2_shapes(cuda):
    yt, yp = synth.layered_pair(1, 496, 768, 8, seed=1001, noise=0.01)
    _check_suite_vs_oracle(yt, yp, 8, cuda)
    yt, yp = synth.layered_pair(1, 496, 1024, 10, seed=2002)                           # K = 10: 9 boundaries
    _check_suite_vs_oracle(yt, yp, 10, cuda)


def test_cfg3_lesions_512(cuda):
    for single in (True, False):
        yt, yp = synth.lesion_pair(2, 512, 512, 4, seed=3003, single_blob_interior=single)
        _check_suite_vs_oracle(yt, yp, 4, cuda)


def test_cfg5_volume_256x256x64(cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    k = 5
    vt, vp = synth.layered_volume_pair(256, 256, 64, k, seed=5005)
    ints = suite.surface_distance_3d(torch.from_numpy(vt).to(cuda), torch.from_numpy(vp).to(cuda), k)
    m = suite.surface_metrics_3d(ints)
    n_pts = ints["n_pts"].cpu().numpy().view(np.uint32)
    max_sq = ints["max_sq"].cpu().numpy().view(np.uint32)
    p95 = ints["p95_sq"].cpu().numpy().view(np.uint32)
    sums = ints["sum_dist"].cpu().numpy()
    for c in range(k):
        ref = so.class_metrics(vt, vp, c)
        assert ref["sq_pred_to_true"] is not None
        for d, key in ((0, "sq_pred_to_true"), (1, "sq_true_to_pred")):
            s = np.sort(ref[key])
            assert n_pts[c, 1 - d] == len(s), (c, d)
            assert max_sq[c, d] == s[-1], (c, d)
            lo_i = int(np.floor((len(s) - 1) * 0.95))
            assert tuple(p95[c, d]) == (s[lo_i], s[min(lo_i + 1, len(s) - 1)]), (c, d)
            np.testing.assert_allclose(sums[c, d], np.sqrt(s.astype(np.float64)).sum(), rtol=1e-12)
        assert m["hausdorff_distance"][c] == ref["hausdorff_distance"]
        np.testing.assert_allclose(m["hausdorff_distance_95"][c], ref["hausdorff_distance_95"], rtol=RTOL)
        np.testing.assert_allclose(m["assd"][c], ref["assd"], rtol=RTOL)


# ------------------------------------------------------------------------------------------ C ABI
def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def test_abi_octm_contour2d_u8(cuda):
    """The one-call contour entry point of include/octm.h, workspace sized by the query call; with and without
    the label pass's first-occurrence table."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib
    lib = _lib.load()
    k, max_pts = 4, 1024
    yt, yp = synth.lesion_pair(3, 96, 128, k, seed=71, single_blob_interior=False)
    n, h, w = yt.shape
    t, p = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    nbytes = int(lib.octm_contour2d_workspace_bytes(n, h, w, k, max_pts))
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cuda)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for with_first in (False, True):
        first = None
        if with_first:
            fp = torch.empty((n, 2, k), dtype=torch.int32, device=cuda)
            _lib.call("octm_label_pass_u8", _p(t), _p(p), n, h, w, k, None, None, None, None, None, None, _p(fp), stream)
            first = _p(fp)
        i32 = dict(dtype=torch.int32, device=cuda)
        n_pts, flags = torch.empty((n, k, 2), **i32), torch.empty((n, k), **i32)
        max_sq, p95 = torch.empty((n, k, 2), **i32), torch.empty((n, k, 2, 2), **i32)
        sums = torch.empty((n, k, 2), dtype=torch.float64, device=cuda)
        _lib.call("octm_contour2d_u8", _p(t), _p(p), n, h, w, k, first, max_pts, _p(n_pts), _p(flags), _p(max_sq),
                  _p(p95), _p(sums), _p(ws), nbytes, stream)
        torch.cuda.synchronize()
        n_pts_h, max_h = n_pts.cpu().numpy().view(np.uint32), max_sq.cpu().numpy().view(np.uint32)
        p95_h, sums_h = p95.cpu().numpy().view(np.uint32), sums.cpu().numpy()
        for i in range(n):
            for c in range(k):
                ref = _oracle_contour_ints(yt[i] == c, yp[i] == c)
                if ref is None:
                    assert n_pts_h[i, c, 0] == 0 or n_pts_h[i, c, 1] == 0
                    continue
                assert tuple(n_pts_h[i, c]) == ref["n_pts"]
                for d in (0, 1):
                    mx, vlo, vhi, sm = ref[d]
                    assert max_h[i, c, d] == mx and tuple(p95_h[i, c, d]) == (vlo, vhi)
                    np.testing.assert_allclose(sums_h[i, c, d], sm, rtol=1e-12)
    # a workspace that is too small is refused, not overrun
    rc = lib.octm_contour2d_u8(_p(t), _p(p), n, h, w, k, None, max_pts, _p(n_pts), _p(flags), _p(max_sq), _p(p95),
                               _p(sums), _p(ws), nbytes - 1, stream)
    assert rc == -4 and b"workspace" in lib.octm_last_error()


def test_abi_octm_column_scan_u8(cuda):
    """K2 alone through the C ABI on a fast-path shape and on a ragged one (generic kernel), all outputs."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (n, h, w, k, seed) in ((3, 496, 512, 8, 81), (2, 37, 50, 11, 82), (2, 600, 256, 6, 83)):
        yt, yp = synth.layered_pair(n, h, w, k, seed=seed, noise=0.01)
        t, p = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
        i64 = dict(dtype=torch.int64, device=cuda)
        thick, bsq, babs = torch.empty((n, k), **i64), torch.empty((n, k - 1), **i64), torch.empty((n, k - 1), **i64)
        bt = torch.empty((n, k - 1, w), dtype=torch.int32, device=cuda)
        bp = torch.empty((n, k - 1, w), dtype=torch.int32, device=cuda)
        _lib.call("octm_column_scan_u8", _p(t), _p(p), n, h, w, k, _p(thick), _p(bsq), _p(babs), _p(bt), _p(bp), stream)
        torch.cuda.synchronize()
        for i in range(n):
            ref = lo.score_bscan_fast(yt[i], yp[i], k)
            assert np.array_equal(thick[i].cpu().numpy(), ref["thickness_absdiff"])
            assert np.array_equal(bsq[i].cpu().numpy(), ref["boundary_sq"])
            assert np.array_equal(babs[i].cpu().numpy(), ref["boundary_abs"])
            assert np.array_equal(bt[i].cpu().numpy(), ref["boundary_true"])
            assert np.array_equal(bp[i].cpu().numpy(), ref["boundary_pred"])
        # sums only (boundary rows NULL)
        thick2 = torch.empty((n, k), **i64)
        _lib.call("octm_column_scan_u8", _p(t), _p(p), n, h, w, k, _p(thick2), None, None, None, None, stream)
        assert torch.equal(thick, thick2)


# ------------------------------------------------------------------------------------------ input validation
@pytest.mark.parametrize("shape,k", [((2, 496, 512), 8), ((2, 33, 50), 8), ((2, 64, 128), 6), ((2, 31, 37), 12)])
@pytest.mark.parametrize("bad", [None, "k", 9, 255])
def test_label_at_or_above_num_classes_raises(cuda, shape, k, bad):
    """ADVICE r1 / VERDICT r1: an ignore label or a wrong class count must not silently alias into other classes.
    Both kernels (TMA strip kernel: W % 16 == 0 and K <= 8; generic: the rest) drop such pixels, the totals kernel
    counts the items whose confusion matrix does not add up to H * W, and reading the results raises."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    n, h, w = shape
    yt, yp = synth.layered_pair(n, h, w, k, seed=91)
    val = k if bad == "k" else bad
    if val is not None and val < k:
        pytest.skip("label is valid for this K")
    for which in (0, 1):
        a, b = yt.copy(), yp.copy()
        if val is not None:
            (a if which == 0 else b)[1, h // 2, w // 3] = val           # ONE bad pixel inside a uniform region
        t, p = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
        res = suite.evaluate(t, p, k, contours=False)
        if val is None:
            res.metrics()
            continue
        with pytest.raises(ValueError, match="label >= num_classes"):
            res.metrics()
        res = suite.evaluate(t, p, k, contours=False, validate=False)   # opt-out: the reference does not check either
        res.metrics()
    # host-array entry point
    if val is not None:
        a = yt.copy()
        a[0, 0, 0] = val
        with pytest.raises(ValueError, match="label >= num_classes"):
            suite.evaluate_host(a, yp, k, contours=False).metrics()


def test_empty_batch(cuda):
    """n_items == 0 must give zeroed totals, not uninitialised memory."""
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    from retinal_oct_image_segmentation_via_deep_learning_b200 import dist as odist
    e = torch.empty((0, 64, 128), dtype=torch.uint8, device=cuda)
    res = suite.evaluate(e, e, 6)
    vec = res.totals_host()
    assert vec[0] == 0 and np.all(vec[:odist.base_len(6)] == 0) and vec[-1] == 0
    tot = odist.dataset_totals(suite.evaluate(e, e, 6), 1)
    assert tot["n_items"] == 0 and tot["confusion"].sum() == 0


def test_cfg4_ragged_predicted_boundaries(cuda):
    """One-pixel teeth, overhangs and detached pixels along every predicted layer boundary (what an argmax looks like):
    the prediction is out of class order in most columns and its contours are walked; the numbers must still be the
    reference's."""
    import torch
    yt, yp = synth.ragged_pair_device(2, 496, 512, 8, seed=91, device=cuda)
    _check_suite_vs_oracle(yt.cpu().numpy(), yp.cpu().numpy(), 8, cuda)


@pytest.mark.parametrize("policy", [1, 2])
def test_both_stream_policies_give_the_reference_numbers(cuda, policy):
    """The order of the first contour-stage steps (and the seed source) follows the data; pinned either way, clean,
    lightly noisy, heavily noisy and ragged pairs must all come out as the oracle says."""
    from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib
    lib = _lib.load()
    before = lib.octm_label_pass_seed_policy(policy)
    try:
        for seed, noise in ((41, 0.0), (42, 2e-5), (43, 2e-3)):
            yt, yp = synth.layered_pair(2, 496, 512, 8, seed=seed, noise=noise)
            _check_suite_vs_oracle(yt, yp, 8, cuda)
        yt, yp = synth.ragged_pair_device(2, 496, 512, 8, seed=44, device=cuda)
        _check_suite_vs_oracle(yt.cpu().numpy(), yp.cpu().numpy(), 8, cuda)
        yt, yp = synth.layered_pair(2, 496, 512, 8, seed=45, noise=1e-4)
        yt[0], yp[1] = yp[0].copy(), yt[1].copy()             # the noisy map as y_true (item 0), both clean-ish (item 1)
        _check_suite_vs_oracle(yt, yp, 8, cuda)
    finally:
        lib.octm_label_pass_seed_policy(before)
