"""GPU parity: scores -> label map (argmax front end) against numpy argmax, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(scores, axis):
    # numpy argmax: first maximal element, NaN counts as maximal
    return np.argmax(scores.astype(np.float32), axis=axis).astype(np.uint8)


@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("shape", [(3, 8, 32, 64), (2, 11, 17, 23), (1, 4, 496, 512), (5, 1, 8, 8), (2, 200, 6, 10)])
def test_argmax_planes_and_generic(cuda, dtype, shape):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    g = torch.Generator().manual_seed(hash((dtype, shape)) % 2**31)
    x = torch.randn(shape, generator=g)
    x = (x * 4).round() / 4                       # many exact ties, also after rounding to 16 bits
    x = x.to(getattr(torch, dtype))
    got = suite.labels_from_scores(x.to(cuda)).cpu().numpy()
    ref = _ref(x.float().numpy(), 1)
    np.testing.assert_array_equal(got, ref)
    xl = x.permute(0, 2, 3, 1).contiguous()
    got = suite.labels_from_scores(xl.to(cuda), channels_last=True).cpu().numpy()
    np.testing.assert_array_equal(got, ref)


def test_argmax_nan_and_inf(cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    x = torch.zeros((1, 5, 4, 8))
    x[0, 3, 0, 0] = float("nan")
    x[0, 1, 0, 1] = float("inf")
    x[0, :, 0, 2] = float("-inf")
    x[0, 2, 0, 3] = float("nan"); x[0, 4, 0, 3] = float("nan")
    x[0, 4, 1, :] = 1e-30
    got = suite.labels_from_scores(x.to(cuda)).cpu().numpy()
    np.testing.assert_array_equal(got, _ref(x.numpy(), 1))
    assert got[0, 0, 0] == 3 and got[0, 0, 1] == 1 and got[0, 0, 2] == 0 and got[0, 0, 3] == 2


def test_evaluate_scores_matches_evaluate_on_labels(cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth
    yt, yp = synth.layered_pair(2, 64, 128, 6, seed=9, noise=0.01)
    onehot = np.eye(6, dtype=np.float32)[yp].transpose(0, 3, 1, 2) * 3.0 - 1.0       # scores whose argmax is yp
    a = suite.evaluate(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda), 6).integers()
    b = suite.evaluate_scores(torch.from_numpy(yt).to(cuda), torch.from_numpy(np.ascontiguousarray(onehot)).to(cuda)).integers()
    for key in a:
        np.testing.assert_array_equal(a[key], b[key], err_msg=key)
