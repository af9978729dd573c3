"""GPU parity: the device float64 epilogue (octm_derive_metrics) against the host mirror (derive.py)
and the oracle; the device totals vector against the host packing used by the all-reduce."""
import numpy as np
import pytest

from oracle import labelmap_oracle as lo
from retinal_oct_image_segmentation_via_deep_learning_b200 import synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("case", ["layered", "lesion", "random", "absent"])
def test_device_metrics_equal_host_mirror_bit_for_bit(cuda, case):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    if case == "layered":
        yt, yp = synth.layered_pair(5, 96, 128, 7, seed=61, noise=0.01)
        k = 7
    elif case == "lesion":
        yt, yp = synth.lesion_pair(4, 96, 96, 4, seed=62, single_blob_interior=False)
        k = 4
    elif case == "random":
        yt, yp = synth.random_pair(3, 33, 50, 12, seed=63)
        k = 12
    else:
        yt = np.zeros((2, 16, 32), np.uint8)
        yp = np.zeros((2, 16, 32), np.uint8)
        yp[1, 4:9, 5:20] = 2
        k = 4
    res = suite.evaluate(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda), k)
    dev, host = res.metrics(), res.metrics_host()
    for name, v in host.items():
        assert _same(dev[name], v), name
    # and against the per-class reference-style oracle for one item
    ref = lo.score_bscan(yt[-1], yp[-1], k)
    for name in lo.COUNT_METRICS + ("thickness_difference", "boundary_mse", "boundary_rmse", "boundary_mad"):
        assert _same(dev[name][-1], ref[name]), name
    assert _same(dev["hausdorff_distance"][-1], ref["hausdorff_distance"])
    for name in ("hausdorff_distance_95", "assd"):
        np.testing.assert_allclose(dev[name][-1], ref[name], rtol=1e-6, equal_nan=True)


def test_device_totals_equal_host_packing(cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import dist as odist
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    k = 6
    yt, yp = synth.layered_pair(9, 64, 96, k, seed=64, noise=0.02)
    res = suite.evaluate(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda), k)
    vec = res.totals_host()
    host = odist.local_partials(res.integers(), res.metrics(), k)
    nb = odist.base_len(k)
    n_exact = nb - 3 * k - 2                 # [.. | sum hd, hd95, assd | n_overflow_items, n_bad_label_items]
    assert np.array_equal(vec[:n_exact], host[:n_exact])                 # integer sums: exact
    np.testing.assert_allclose(vec[n_exact:nb], host[n_exact:nb], rtol=1e-12)   # float sums: order only
    m = res.metrics()
    np.testing.assert_array_equal(vec[nb:nb + k], np.nanmax(m["hausdorff_distance"], axis=0))
    tot = odist.dataset_totals(res, 1, want_max=True)
    cm = sum(lo.confusion_matrix(yt[i], yp[i], k).astype(np.int64) for i in range(len(yt)))
    assert np.array_equal(tot["confusion"], cm) and tot["n_items"] == len(yt)


def test_evaluate_host_streams_chunks(cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    k = 5
    yt, yp = synth.layered_pair(11, 48, 64, k, seed=65, noise=0.02)
    a = suite.evaluate(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda), k).metrics()
    for pack in (True, False, "auto"):           # packed transfer (two labels per byte) and the plain copy agree
        b = suite.evaluate_host(yt, yp, k, device=cuda, chunk_items=4, pack=pack).metrics()
        for name in a:
            assert _same(a[name], b[name]), (name, pack)


def test_packed_transfer_odd_sizes(cuda):
    """host pack -> device unpack round trip on sizes that are not multiples of the vector widths"""
    import ctypes
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(66)
    for n in (1, 2, 31, 32, 33, 4097, 1 << 20, (1 << 20) + 17):
        src = rng.integers(0, 16, n, dtype=np.uint8)
        packed = np.zeros((n + 1) // 2, np.uint8)
        assert lib.octm_host_pack_nibbles(src.ctypes.data, packed.ctypes.data, n, 4) == 0
        dp = torch.from_numpy(packed).to(cuda)
        out = torch.full((n + 5,), 0xEE, dtype=torch.uint8, device=cuda)
        _lib.call("octm_unpack_nibbles_u8", ctypes.c_void_p(dp.data_ptr()), n, ctypes.c_void_p(out.data_ptr()),
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        got = out.cpu().numpy()
        np.testing.assert_array_equal(got[:n], src)
        assert (got[n:] == 0xEE).all()
