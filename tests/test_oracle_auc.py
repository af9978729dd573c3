"""CPU tier: the auc_score oracle against the executed-reference goldens and against live scikit-learn."""
import warnings

import numpy as np
import pytest

from oracle import metrics_oracle as mo


def test_oracle_matches_executed_reference(golden_dir):
    g = np.load(f"{golden_dir}/auc_golden.npz")
    assert "executed reference" in str(g["source"])
    for name in g["names"]:
        ref = float(g[f"{name}/auc"])
        got = mo.auc_score(g[f"{name}/y_true"], g[f"{name}/scores"])
        if np.isnan(ref):
            assert np.isnan(got), name
        else:
            assert got == ref, name                     # same arithmetic, bit for bit


def test_rank_sum_form_equals_trapezoid():
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(2, 400))
        t = (rng.random(n) < rng.uniform(0.1, 0.9)).astype(np.uint8)
        if t.min() == t.max():
            t[0] = 1 - t[0]
        s = np.round(rng.normal(size=n), int(rng.integers(0, 4)))      # plenty of ties
        u2, npos, nneg = mo.auc_rank_sum(t, s)
        np.testing.assert_allclose(u2 / (2.0 * npos * nneg), mo.auc_score(t, s), rtol=1e-12, atol=1e-15)


def test_oracle_against_live_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(4)
    for _ in range(30):
        n = int(rng.integers(2, 2000))
        t = (rng.random(n) < 0.4).astype(np.uint8)
        if t.min() == t.max():
            t[0] = 1 - t[0]
        s = rng.random(n).astype(rng.choice([np.float32, np.float64]))
        if rng.random() < 0.5:
            s = np.round(s, 2)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            # roc_curve drops collinear points (drop_intermediate) before the trapezoid: same area, last bits differ
            np.testing.assert_allclose(mo.auc_score(t, s), sk.roc_auc_score(t, s), rtol=1e-13, atol=0)
