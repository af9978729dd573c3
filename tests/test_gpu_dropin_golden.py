"""GPU parity of the REFERENCE-FACING surface: all 17 drop-in functions of ``<package>/Metrics/*.py`` against the
outputs of the EXECUTED reference stored in ``tests/golden`` (``counts_golden.npz``, ``auc_golden.npz``; generated
by ``oracle/make_golden.py`` from the unmodified ``Metrics/*.py``), the contour three against
``contours_golden.npz`` and the published scikit-image vectors, plus the reference's exception behaviour
(``Contour_based_metrics.py:15-16``: IndexError on a contour-less mask, ValueError for non-2-D input).

Counts-derived scalars must be EQUAL (0 ulp); floats derived from distances within 1e-6 relative."""
import importlib
import sys

import numpy as np
import pytest

from test_oracle_skimage_vectors import published_cases

pytestmark = pytest.mark.gpu
RTOL = 1e-6

# golden key -> (drop-in module, function name)
COUNT_FUNCS = {
    "accuracy": ("ConfusionMatrix_based_metrics", "accuracy"),
    "sensitivity": ("ConfusionMatrix_based_metrics", "sensitivity"),
    "cm_precision": ("ConfusionMatrix_based_metrics", "precision"),
    "specificity": ("ConfusionMatrix_based_metrics", "specificity"),
    "dice_coefficient": ("Region_based_metrics", "dice_coefficient"),
    "iou_score": ("Region_based_metrics", "iou_score"),
    "region_precision": ("Region_based_metrics", "precision"),
    "recall": ("Region_based_metrics", "recall"),
    "mean_squared_error": ("PixelError_based_metrics", "mean_squared_error"),
    "root_mean_squared_error": ("PixelError_based_metrics", "root_mean_squared_error"),
    "mad": ("Contour_based_metrics", "mad"),
    "vascularity_index": ("Biomarker_based_metrics", "vascularity_index"),
    "thickness_difference": ("Biomarker_based_metrics", "thickness_difference"),
}


@pytest.fixture(scope="module")
def dropin(cuda):
    """The five modules imported the way a user of the reference imports them: Metrics/ on sys.path."""
    from retinal_oct_image_segmentation_via_deep_learning_b200 import METRICS_DIR
    sys.path.insert(0, METRICS_DIR)
    try:
        mods = {}
        for name in ("ConfusionMatrix_based_metrics", "Region_based_metrics", "Contour_based_metrics",
                     "PixelError_based_metrics", "Biomarker_based_metrics"):
            sys.modules.pop(name, None)
            mods[name] = importlib.import_module(name)
            assert mods[name].__file__.startswith(METRICS_DIR)
        yield mods
    finally:
        sys.path.remove(METRICS_DIR)


def test_the_surface_is_the_references_17_functions(dropin):
    names = {m: sorted(n for n in dir(mod) if callable(getattr(mod, n)) and not n.startswith("_")
                       and getattr(getattr(mod, n), "__module__", "") == m) for m, mod in dropin.items()}
    assert names["ConfusionMatrix_based_metrics"] == ["accuracy", "auc_score", "precision", "sensitivity", "specificity"]
    assert names["Region_based_metrics"] == ["dice_coefficient", "iou_score", "precision", "recall"]
    assert names["Contour_based_metrics"] == ["assd", "hausdorff_distance", "hausdorff_distance_95", "mad"]
    assert names["PixelError_based_metrics"] == ["mean_squared_error", "root_mean_squared_error"]
    assert names["Biomarker_based_metrics"] == ["thickness_difference", "vascularity_index"]
    assert sum(map(len, names.values())) == 17


def test_13_count_functions_equal_the_executed_reference(dropin, golden_dir):
    g = np.load(f"{golden_dir}/counts_golden.npz")
    assert str(g["source"]).startswith("executed reference")
    checked = 0
    for name in g["names"]:
        yt, yp, k = g[f"{name}/y_true"], g[f"{name}/y_pred"], int(g[f"{name}/K"])
        for c in range(k):
            # the reference is called with int64 masks (uint8 wraps in thickness_difference) ...
            mt, mp = (yt == c).astype(np.int64), (yp == c).astype(np.int64)
            for key, (mod, fn) in COUNT_FUNCS.items():
                got = getattr(dropin[mod], fn)(mt, mp)
                want = g[f"{name}/{key}"][c]
                assert isinstance(got, np.floating), (name, key)
                assert np.array_equal(np.float64(got), want, equal_nan=True), (name, c, key, got, want)
                checked += 1
            # ... and bool masks give the same value (reference: golden dice_bool)
            assert dropin["Region_based_metrics"].dice_coefficient(yt == c, yp == c) == g[f"{name}/dice_bool"][c]
    assert checked >= 13 * 30


def test_count_functions_accept_uint8_masks_and_cuda_tensors(dropin, golden_dir, cuda):
    import torch
    g = np.load(f"{golden_dir}/counts_golden.npz")
    name = "layered48x64_0"
    yt, yp = g[f"{name}/y_true"], g[f"{name}/y_pred"]
    for c in range(int(g[f"{name}/K"])):
        mt8, mp8 = (yt == c).astype(np.uint8), (yp == c).astype(np.uint8)
        tt, tp = torch.from_numpy(mt8).to(cuda), torch.from_numpy(mp8).to(cuda)
        for key, (mod, fn) in COUNT_FUNCS.items():
            want = g[f"{name}/{key}"][c]
            assert np.array_equal(np.float64(getattr(dropin[mod], fn)(mt8, mp8)), want), (key, "uint8")
            assert np.array_equal(np.float64(getattr(dropin[mod], fn)(tt, tp)), want), (key, "cuda")


def test_pixel_error_on_non_binary_uint8_follows_astype_float(dropin):
    """uint8 arrays with values > 1 (label maps, boundary rows stored as uint8) are NOT masks; the
    reference computes astype(float) differences for any values (PixelError_based_metrics.py:14-17)."""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 8, size=(33, 50)).astype(np.uint8)
    b = rng.integers(0, 8, size=(33, 50)).astype(np.uint8)
    d = a.astype(float) - b.astype(float)
    assert dropin["PixelError_based_metrics"].mean_squared_error(a, b) == np.mean(d ** 2)
    assert dropin["PixelError_based_metrics"].root_mean_squared_error(a, b) == np.sqrt(np.mean(d ** 2))
    assert dropin["Contour_based_metrics"].mad(a, b) == np.mean(np.abs(d))
    bt = rng.integers(0, 250, size=(9, 64)).astype(np.uint8)        # boundary rows that happen to fit uint8
    bp = (bt.astype(int) + rng.integers(-3, 4, size=bt.shape)).clip(0, 255).astype(np.uint8)
    d = bt.astype(float) - bp.astype(float)
    assert dropin["PixelError_based_metrics"].mean_squared_error(bt, bp) == np.mean(d ** 2)
    assert dropin["Contour_based_metrics"].mad(bt, bp) == np.mean(np.abs(d))


def test_auc_score_equals_the_executed_reference(dropin, golden_dir):
    g = np.load(f"{golden_dir}/auc_golden.npz")
    assert str(g["source"]).startswith("executed reference")
    for name in g["names"]:
        got = dropin["ConfusionMatrix_based_metrics"].auc_score(g[f"{name}/y_true"], g[f"{name}/scores"])
        want = float(g[f"{name}/auc"])
        assert isinstance(got, float)
        if np.isnan(want):
            assert np.isnan(got), name
        else:
            assert abs(got - want) <= 4 * np.finfo(np.float64).eps, (name, got, want)


def test_contour_functions_against_contour_goldens(dropin, golden_dir):
    g = np.load(f"{golden_dir}/contours_golden.npz")
    C = dropin["Contour_based_metrics"]
    for name in g["names"]:
        a, b = g[f"{name}/mask_true"], g[f"{name}/mask_pred"]
        hd, hd95, assd = g[f"{name}/metrics"]
        got = C.hausdorff_distance(a, b)
        assert isinstance(got, float) and got == hd, name                            # bit-exact (sqrt of an integer / 4)
        np.testing.assert_allclose(C.hausdorff_distance_95(a, b), hd95, rtol=RTOL, err_msg=name)
        np.testing.assert_allclose(C.assd(a, b), assd, rtol=RTOL, err_msg=name)
        assert isinstance(C.assd(a, b), np.floating)


def test_contour_functions_on_published_skimage_vectors(dropin):
    """find_contours(mask, .5)[0] is fixed by the published vectors; the reference's expressions
    (Contour_based_metrics.py:19-22, 36-39, 53-56) evaluated verbatim on them give the expected scalars."""
    C = dropin["Contour_based_metrics"]
    cases = [c for c in published_cases() if c[0] != "test_float_5x5_radius"]       # binary images only
    (_, img_a, _, cont_a), (_, img_b, _, cont_b) = cases
    pad = np.zeros((8, 8))
    pad[:3, :3] = img_a                                                              # same contour, same coordinates
    A, B = cont_a[0], cont_b[0]
    for yt, yp, ca, cb in ((pad, img_b, A, B), (img_b, pad, B, A), (img_b, img_b, B, B)):
        d1 = [np.min(np.sqrt(np.sum((ca - p) ** 2, axis=1))) for p in cb]           # reference :19-20 verbatim
        d2 = [np.min(np.sqrt(np.sum((cb - p) ** 2, axis=1))) for p in ca]
        assert C.hausdorff_distance(yt, yp) == max(np.max(d1), np.max(d2))
        np.testing.assert_allclose(C.hausdorff_distance_95(yt, yp), max(np.percentile(d1, 95), np.percentile(d2, 95)), rtol=RTOL)
        np.testing.assert_allclose(C.assd(yt, yp), (np.mean(d1) + np.mean(d2)) / 2, rtol=RTOL)


def test_exception_behaviour_of_the_reference(dropin):
    C = dropin["Contour_based_metrics"]
    blob = np.zeros((8, 8), np.uint8)
    blob[2:5, 3:6] = 1
    for fn in (C.hausdorff_distance, C.hausdorff_distance_95, C.assd):
        with pytest.raises(IndexError):                      # find_contours(...)[0] on an empty list, :15-16
            fn(np.zeros((8, 8), np.uint8), blob)
        with pytest.raises(IndexError):
            fn(blob, np.ones((8, 8), np.uint8))              # a full mask has no contour either
        with pytest.raises(ValueError):                      # skimage: "Only 2D arrays are supported."
            fn(np.zeros((2, 8, 8), np.uint8), np.zeros((2, 8, 8), np.uint8))
        with pytest.raises(ValueError):                      # skimage: "Input array must be at least 2x2."
            fn(np.zeros((1, 8), np.uint8), np.zeros((1, 8), np.uint8))
    with pytest.raises(ValueError):                          # numpy broadcasting error in the reference
        dropin["Region_based_metrics"].dice_coefficient(np.zeros((4, 4), bool), np.zeros((4, 5), bool))
    with pytest.raises(ValueError):                          # not a mask: the reference is binary-only
        dropin["Region_based_metrics"].dice_coefficient(np.full((4, 4), 3, np.uint8), np.zeros((4, 4), np.uint8))
    # size-0 arrays: accuracy is nan in the reference (0/0), the epsilon forms give 0
    with np.errstate(all="ignore"):
        assert np.isnan(dropin["ConfusionMatrix_based_metrics"].accuracy(np.zeros((0, 4), bool), np.zeros((0, 4), bool)))
    assert dropin["Region_based_metrics"].dice_coefficient(np.zeros((0, 4), bool), np.zeros((0, 4), bool)) == 0.0
