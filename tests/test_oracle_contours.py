"""CPU tier: invariants of the restated find_contours (parity unpinned: scikit-image is absent)."""
import numpy as np

from oracle import contours_oracle as co
from oracle import labelmap_oracle as lo
from oracle import metrics_oracle as mo


def _rand_masks(seed, n=60, lo_=2, hi=16):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        h, w = rng.integers(lo_, hi, size=2)
        yield (rng.random((h, w)) < rng.uniform(0.15, 0.85)).astype(np.int64)


def test_fast_scan_equals_literal_scan():
    for m in _rand_masks(1):
        assert co.contour_segments(m, 0.5) == co.contour_segments_literal(m, 0.5)


def test_vertices_are_exactly_the_cracks():
    """Union of all contours' vertices == midpoints between unequal 4-neighbours."""
    for m in _rand_masks(2):
        cs = co.find_contours(m)
        cracks = set(map(tuple, co.all_crack_vertices(m).tolist()))
        got = set()
        for c in cs:
            got |= set(map(tuple, co.to_lattice(c).tolist()))
        assert got == cracks


def test_closed_contours_repeat_their_first_vertex_and_open_ones_end_on_the_border():
    for m in _rand_masks(3):
        h, w = m.shape
        for c in co.find_contours(m):
            v = co.to_lattice(c)
            closed = len(v) > 2 and tuple(v[0]) == tuple(v[-1])
            body = v[:-1] if closed else v
            assert len(set(map(tuple, body.tolist()))) == len(body)          # simple path
            if not closed:
                for end in (v[0], v[-1]):                                   # ends lie on a border crack
                    assert end[0] in (0, 2 * h - 2) or end[1] in (0, 2 * w - 2)
            steps = np.abs(np.diff(v, axis=0)).sum(axis=1)
            assert np.all(steps == 2)                                        # consecutive vertices are adjacent cracks


def test_first_contour_contains_the_first_segment():
    for m in _rand_masks(4):
        segs = co.contour_segments(m, 0.5)
        if not segs:
            assert co.find_contours(m) == []
            continue
        first = co.find_contours(m)[0]
        pts = set(map(tuple, first.tolist()))
        assert tuple(segs[0][0]) in pts and tuple(segs[0][1]) in pts


def test_no_contour_raises_indexerror_like_the_reference():
    z = np.zeros((5, 6), np.int64)
    for a, b in ((z, z + 1), (z + 1, z), (z, z)):
        try:
            mo.hausdorff_distance(a, b)
            raise AssertionError("expected IndexError")
        except IndexError:
            pass
    try:
        co.find_contours(np.zeros((1, 5)))
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def test_sq_distances_equal_reference_float_distances_and_edt():
    """sqrt(D2/4) == the reference's float64 min-distance expression bit-for-bit; D2 == scipy EDT^2."""
    from scipy.ndimage import distance_transform_edt
    rng = np.random.default_rng(5)
    done = 0
    for m1 in _rand_masks(6, n=40, lo_=4, hi=24):
        m2 = np.roll(m1, 1, axis=1) ^ (rng.random(m1.shape) < 0.1)
        im = lo.contour_intermediates(m1, m2)
        if im is None:
            continue
        a = co.find_contours(m1)[0]
        b = co.find_contours(m2.astype(np.int64))[0]
        ref = np.array(mo._directed_min_distances(a, b))
        assert np.array_equal(np.sqrt(im["sq_pred_to_true"] / 4.0), ref)
        h, w = m1.shape
        grid = np.ones((2 * h - 1, 2 * w - 1), bool)
        grid[im["verts_true"][:, 0], im["verts_true"][:, 1]] = False
        edt2 = np.rint(distance_transform_edt(grid) ** 2).astype(np.int64)
        assert np.array_equal(im["sq_pred_to_true"], edt2[im["verts_pred"][:, 0], im["verts_pred"][:, 1]])
        done += 1
    assert done > 10


def test_contour_golden_regression(golden_dir):
    g = np.load(f"{golden_dir}/contours_golden.npz")
    assert "parity unpinned" in str(g["source"])
    for name in g["names"]:
        a, b = g[f"{name}/mask_true"], g[f"{name}/mask_pred"]
        im = lo.contour_intermediates(a, b)
        for key in ("verts_true", "verts_pred", "sq_pred_to_true", "sq_true_to_pred"):
            assert np.array_equal(im[key], g[f"{name}/{key}"]), (name, key)
        m = lo.contour_metrics_from_sq(im["sq_pred_to_true"], im["sq_true_to_pred"])
        ref = g[f"{name}/metrics"]
        assert m["hausdorff_distance"] == ref[0]
        np.testing.assert_allclose([m["hausdorff_distance_95"], m["assd"]], ref[1:], rtol=1e-12)


def test_derive_contour_metrics_from_integers():
    """The product epilogue (derive.contour_metrics) fed oracle integers == numpy percentile / mean path."""
    from retinal_oct_image_segmentation_via_deep_learning_b200 import derive
    rng = np.random.default_rng(9)
    for _ in range(50):
        n1, n2 = rng.integers(1, 60, size=2)
        d1 = rng.integers(0, 400, size=n1)     # pred -> true: n_pred values
        d2 = rng.integers(0, 400, size=n2)
        ref = lo.contour_metrics_from_sq(d1, d2)

        def stats(d):
            s = np.sort(d)
            pos = (len(s) - 1) * 0.95
            lo_i = int(np.floor(pos))
            return s[lo_i], s[min(lo_i + 1, len(s) - 1)]
        got = derive.contour_metrics(np.array([n2, n1]), np.array([d1.max(), d2.max()]),
                                     np.array([stats(d1), stats(d2)]),
                                     np.array([np.sqrt(d1 / 4.0).sum(), np.sqrt(d2 / 4.0).sum()]))
        assert got["hausdorff_distance"] == ref["hausdorff_distance"]
        np.testing.assert_allclose(got["hausdorff_distance_95"], ref["hausdorff_distance_95"], rtol=1e-12)
        np.testing.assert_allclose(got["assd"], ref["assd"], rtol=1e-12)
