"""CPU tier: the oracle and the product's float64 epilogue against outputs of the EXECUTED reference."""
import numpy as np
import pytest

from oracle import labelmap_oracle as lo
from oracle import metrics_oracle as mo
from oracle import ref_loader
from retinal_oct_image_segmentation_via_deep_learning_b200 import derive

FUNCS = lo.COUNT_METRICS + ("thickness_difference",)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(f"{golden_dir}/counts_golden.npz")


def test_golden_came_from_the_reference(golden):
    assert str(golden["source"]).startswith("executed reference")
    assert len(golden["names"]) >= 10


def test_oracle_matches_reference_outputs_bit_exact(golden):
    """metrics_oracle == the reference's own functions on the same int64 masks, 0 ulp (nan == nan)."""
    for name in golden["names"]:
        yt, yp, k = golden[f"{name}/y_true"], golden[f"{name}/y_pred"], int(golden[f"{name}/K"])
        for fn in FUNCS:
            got = np.array([getattr(mo, fn)((yt == c).astype(np.int64), (yp == c).astype(np.int64)) for c in range(k)])
            assert np.array_equal(got, golden[f"{name}/{fn}"], equal_nan=True), (name, fn)


def test_closed_forms_over_confusion_counts_bit_exact(golden):
    """The product's epilogue (derive.py) fed the oracle's K x K confusion matrix reproduces the reference."""
    for name in golden["names"]:
        yt, yp, k = golden[f"{name}/y_true"], golden[f"{name}/y_pred"], int(golden[f"{name}/K"])
        cm = lo.confusion_matrix(yt, yp, k)
        m = derive.count_metrics(*derive.class_counts(cm))
        for fn in lo.COUNT_METRICS:
            assert np.array_equal(m[fn], golden[f"{name}/{fn}"], equal_nan=True), (name, fn)
        if yt.ndim == 2:
            fast = lo.score_bscan_fast(yt, yp, k)
            td = derive.thickness_difference(fast["thickness_absdiff"], yt.shape[1])
            assert np.array_equal(td, golden[f"{name}/thickness_difference"]), name
        assert np.array_equal(m["dice_coefficient"], golden[f"{name}/dice_bool"]), name


def test_from_counts_helper(golden):
    name = golden["names"][0]
    yt, yp, k = golden[f"{name}/y_true"], golden[f"{name}/y_pred"], int(golden[f"{name}/K"])
    tp, fp, fn, tn, n = derive.class_counts(lo.confusion_matrix(yt, yp, k))
    for c in range(k):
        d = mo.from_counts(tp[c], fp[c], fn[c], tn[c])
        assert d["dice_coefficient"] == golden[f"{name}/dice_coefficient"][c]
        assert d["specificity"] == golden[f"{name}/specificity"][c]


def test_edge_behaviours_of_the_reference(golden):
    """both masks empty -> acc 1, sens 0, prec 0, dice 0, spec just under 1."""
    acc = golden["both_empty/accuracy"][1]
    assert acc == 1.0 and golden["both_empty/sensitivity"][1] == 0.0 and golden["both_empty/dice_coefficient"][1] == 0.0
    assert 0.999999 < golden["both_empty/specificity"][1] < 1.0


@pytest.mark.skipif(ref_loader.load() is None, reason="reference checkout not on this machine")
def test_oracle_against_live_reference_random():
    ref = ref_loader.load()
    rng = np.random.default_rng(8)
    for _ in range(25):
        h, w = rng.integers(1, 40, size=2)
        a = (rng.random((h, w)) < rng.random()).astype(np.int64)
        b = (rng.random((h, w)) < rng.random()).astype(np.int64)
        for fn in FUNCS + ("mad",):
            assert np.array_equal(getattr(mo, fn)(a, b), getattr(ref, fn)(a, b), equal_nan=True), fn
    # shape-agnostic pixel-error functions on boundary arrays
    bt = rng.integers(0, 496, size=(9, 64))
    bp = bt + rng.integers(-5, 6, size=bt.shape)
    assert mo.mean_squared_error(bt, bp) == ref.mean_squared_error(bt, bp)
    assert mo.root_mean_squared_error(bt, bp) == ref.root_mean_squared_error(bt, bp)
    d = (bt - bp).astype(np.int64)
    assert derive.boundary_errors((d * d).sum(), np.abs(d).sum(), bt.size)["boundary_mse"] == ref.mean_squared_error(bt, bp)


def test_suite_golden_is_consistent(golden_dir):
    g = np.load(f"{golden_dir}/suite_golden.npz")
    for name in g["names"]:
        yt, yp, k = g[f"{name}/y_true"], g[f"{name}/y_pred"], int(g[f"{name}/K"])
        fast = lo.score_bscan_fast(yt[0], yp[0], k)
        for key, v in fast.items():
            assert np.array_equal(v, g[f"{name}/{key}"][0]), (name, key)
        # boundaries are cumulative thickness: b_k = sum_{c<k} thickness_c
        assert np.array_equal(np.cumsum(fast["thickness_true"], axis=0)[:-1], fast["boundary_true"])
