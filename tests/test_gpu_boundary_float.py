"""GPU parity: boundary-position error on continuous (soft) positions and topology violations, against the
reference's own array expressions evaluated in numpy.  Tolerance 1e-6 relative (north_star); observed ~1e-15."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _ref_mse(a, b):          # Metrics/PixelError_based_metrics.py:14-17
    return np.mean((a.astype(float) - b.astype(float)) ** 2)


def _ref_mad(a, b):          # Metrics/Contour_based_metrics.py:68-71
    return np.mean(np.abs(a.astype(float) - b.astype(float)))


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "bfloat16"])
def test_soft_boundary_rows(cuda, dtype):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    rng = np.random.default_rng(5)
    n, kb, w = 7, 9, 1024                                    # cfg2 geometry: 9 boundaries x 1024 A-scans
    base = np.cumsum(rng.uniform(20, 50, size=(n, kb, 1)), axis=1) + 8 * np.sin(np.arange(w) / 90.0)
    bt = torch.from_numpy(base + rng.normal(0, 0.3, (n, kb, w))).to(getattr(torch, dtype))
    bp = torch.from_numpy(base + rng.normal(0, 1.5, (n, kb, w))).to(getattr(torch, dtype))
    m = suite.boundary_metrics(bt.to(cuda), bp.to(cuda))
    a, b = bt.to(torch.float64).numpy(), bp.to(torch.float64).numpy()
    for i in range(n):
        for k in range(kb):
            np.testing.assert_allclose(m["boundary_mse"][i, k].item(), _ref_mse(a[i, k], b[i, k]), rtol=RTOL, atol=0)
            np.testing.assert_allclose(m["boundary_rmse"][i, k].item(), np.sqrt(_ref_mse(a[i, k], b[i, k])), rtol=RTOL, atol=0)
            np.testing.assert_allclose(m["boundary_mad"][i, k].item(), _ref_mad(a[i, k], b[i, k]), rtol=RTOL, atol=0)


def test_topology_violations(cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    rng = np.random.default_rng(6)
    pos = np.cumsum(rng.uniform(0, 4, size=(3, 6, 300)), axis=1) + rng.normal(0, 2.0, (3, 6, 300))
    sv, nv = suite.topology_violations(torch.from_numpy(pos.astype(np.float32)).to(cuda))
    p32 = pos.astype(np.float32).astype(np.float64)
    viol = np.maximum(p32[:, :-1] - p32[:, 1:], 0.0)          # layer_engine.py:74-76
    np.testing.assert_allclose(sv.cpu().numpy(), viol.sum(-1), rtol=1e-12)
    np.testing.assert_array_equal(nv.cpu().numpy(), (viol > 0).sum(-1))


def test_dropin_functions_take_float_arrays(cuda):
    from retinal_oct_image_segmentation_via_deep_learning_b200.Metrics import Contour_based_metrics as cbm
    from retinal_oct_image_segmentation_via_deep_learning_b200.Metrics import PixelError_based_metrics as pem
    rng = np.random.default_rng(7)
    a = rng.normal(100, 20, (9, 1024)).astype(np.float32)
    b = a + rng.normal(0, 2, a.shape).astype(np.float32)
    np.testing.assert_allclose(pem.mean_squared_error(a, b), _ref_mse(a, b), rtol=RTOL, atol=0)
    np.testing.assert_allclose(pem.root_mean_squared_error(a, b), np.sqrt(_ref_mse(a, b)), rtol=RTOL, atol=0)
    np.testing.assert_allclose(cbm.mad(a, b), _ref_mad(a, b), rtol=RTOL, atol=0)
    ai, bi = np.rint(a).astype(np.int64), np.rint(b).astype(np.int64)
    assert pem.mean_squared_error(ai, bi) == _ref_mse(ai, bi)           # integer data stays exact
    assert cbm.mad(ai, bi) == _ref_mad(ai, bi)
