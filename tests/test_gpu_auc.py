"""GPU parity: auc_score (radix sort + tie-aware rank sums) against the executed-reference goldens and the oracle.
Tolerance 1e-6 relative (north_star); the exact integer rank sum makes the observed error a few ulp."""
import warnings

import numpy as np
import pytest

from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _gpu_auc(m, p, cuda, dtype=None):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    t = torch.from_numpy(np.ascontiguousarray(m)).to(cuda)
    s = torch.from_numpy(np.ascontiguousarray(p)).to(cuda)
    if dtype is not None:
        s = s.to(dtype)
    return suite.auc_scores(t[None], s[None]).cpu().numpy()[0]


def test_golden_auc(cuda, golden_dir):
    g = np.load(f"{golden_dir}/auc_golden.npz")
    for name in g["names"]:
        ref = float(g[f"{name}/auc"])
        got = _gpu_auc(g[f"{name}/y_true"], g[f"{name}/scores"], cuda)
        if np.isnan(ref):
            assert np.isnan(got), name
        else:
            np.testing.assert_allclose(got, ref, rtol=RTOL, atol=0, err_msg=str(name))
            assert abs(got - ref) <= 4 * np.finfo(np.float64).eps, name


def test_dropin_function_matches_goldens(cuda, golden_dir):
    from retinal_oct_image_segmentation_via_deep_learning_b200.Metrics import ConfusionMatrix_based_metrics as cmm
    g = np.load(f"{golden_dir}/auc_golden.npz")
    for name in g["names"]:
        ref = float(g[f"{name}/auc"])
        got = cmm.auc_score(g[f"{name}/y_true"], g[f"{name}/scores"])
        assert isinstance(got, float)
        assert (np.isnan(got) and np.isnan(ref)) or abs(got - ref) <= 1e-6 * abs(ref), name
    assert cmm.auc_score(np.zeros(4), np.zeros(5)) == 0.0           # length mismatch -> ValueError -> 0.0
    assert cmm.auc_score(np.array([0, 2, 2, 0]), np.array([.1, .9, .8, .2])) == 1.0   # any two label values


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "bfloat16"])
def test_random_batches_all_dtypes(cuda, dtype):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    rng = np.random.default_rng(hash(dtype) % 2**31)
    n, h, w = 5, 70, 93                                     # 6510 elements: ragged warp and thread chunks
    m = (rng.random((n, h, w)) < 0.3).astype(np.uint8)
    m[3] = 1                                                # single class
    s = rng.normal(size=(n, h, w))
    s[1] = np.round(s[1], 1)                                # ties
    s[2] = 0.25
    st = torch.from_numpy(s).to(getattr(torch, dtype))
    got = suite.auc_scores(torch.from_numpy(m).to(cuda), st.to(cuda)).cpu().numpy()
    s_exact = st.to(torch.float64).numpy()                  # what the kernel sorts
    for i in range(n):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = mo.auc_score(m[i], s_exact[i])
        if np.isnan(ref):
            assert np.isnan(got[i])
        else:
            np.testing.assert_allclose(got[i], ref, rtol=RTOL, atol=0)
            u2, npos, nneg = mo.auc_rank_sum(m[i], s_exact[i])
            assert got[i] == u2 / (2.0 * npos * nneg)       # the integer rank sum is exact


def test_full_size_item_and_many_items(cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    rng = np.random.default_rng(9)
    n = 160                                                 # more items than CTAs: persistent loop + workspace reuse
    m = (rng.random((n, 40, 50)) < 0.5).astype(np.uint8)
    s = rng.random((n, 40, 50)).astype(np.float32)
    got = suite.auc_scores(torch.from_numpy(m).to(cuda), torch.from_numpy(s).to(cuda)).cpu().numpy()
    for i in range(0, n, 7):
        u2, npos, nneg = mo.auc_rank_sum(m[i], s[i].astype(np.float64))
        assert got[i] == u2 / (2.0 * npos * nneg)
    m = (rng.random((2, 496, 512)) < 0.2).astype(np.uint8)
    s = (m * 0.3 + rng.random((2, 496, 512)) * 0.7).astype(np.float32)
    got = suite.auc_scores(torch.from_numpy(m).to(cuda), torch.from_numpy(s).to(cuda)).cpu().numpy()
    for i in range(2):
        u2, npos, nneg = mo.auc_rank_sum(m[i], s[i].astype(np.float64))
        assert got[i] == u2 / (2.0 * npos * nneg)
