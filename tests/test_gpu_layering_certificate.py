"""GPU parity: the label pass's layering certificate (``unsorted[i]`` bit m = some column of map m is not
non-decreasing from top to bottom) against numpy, on both kernels; it decides whether a contour may be measured from the
boundary rows alone, so a false "sorted" would be a silent wrong result."""
import numpy as np
import pytest

from retinal_oct_image_segmentation_via_deep_learning_b200 import synth

pytestmark = pytest.mark.gpu


def _expected(yt, yp, tall_limit=504, fast=True):
    out = np.zeros(len(yt), np.uint32)
    for i in range(len(yt)):
        for m, a in enumerate((yt[i], yp[i])):
            if (np.diff(a.astype(np.int16), axis=0) < 0).any():
                out[i] |= 1 << m
    if fast and yt.shape[1] > tall_limit:
        out[:] = 3                                  # the strip kernel does not certify items taller than 504 rows
    return out


def _run(yt, yp, k, cuda):
    import torch
    from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib, suite
    t, p = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    lp = suite.label_pass(t, p, k, seeds=True, boundaries=True, certify=True)
    fast = bool(_lib.load().octm_label_pass_path(yt.shape[1], yt.shape[2], k, t.data_ptr(), p.data_ptr()))
    return lp.unsorted.cpu().numpy().view(np.uint32), fast


@pytest.mark.parametrize("shape,k", [((6, 496, 512), 8), ((5, 124, 256), 6), ((4, 62, 128), 3), ((3, 496, 1024), 8),
                                     ((4, 63, 96), 8), ((5, 33, 50), 8), ((3, 64, 128), 12), ((2, 600, 256), 5)])
def test_certificate_matches_numpy(cuda, shape, k):
    n, h, w = shape
    rng = np.random.default_rng(h * w + k)
    yt, yp = synth.layered_pair(n, h, w, k, seed=11 + k)
    # item 0: clean.  others: single defects of every kind the two-part check has to catch
    if n > 1:
        yp[1, h // 2, w // 3] = (yp[1, h // 2, w // 3] + 1) % k                     # one stray pixel
    if n > 2:
        x = w - 5
        col = yt[2, :, x].copy()
        r = int(np.argmax(col > 0))                                                  # first row of class >= 1
        if 1 <= r < h - 2:
            yt[2, r, x], yt[2, r + 1, x] = col[r - 1], col[r]                         # ... 0 1 -> 0 0? keep order: no-op-safe
            yt[2, r - 1, x] = col[r]                                                 # swap across the boundary: 1 above 0
    if n > 3:
        yt[3, 0, 0] = k - 1                                                          # largest label in the very first row
    if n > 4:
        yp[4, h - 1, w - 1] = 0                                                      # smallest label in the very last row
    if n > 5:
        rows = rng.integers(1, h - 1, size=8)
        yt[5, rows, rng.integers(0, w, size=8)] = rng.integers(0, k, size=8)         # sparse salt in y_true only
    got, fast = _run(yt, yp, k, cuda)
    np.testing.assert_array_equal(got, _expected(yt, yp, fast=fast))
    assert got[0] == (3 if (fast and h > 504) else 0)


def test_interleaving_defects(cuda):
    """Defects that keep BOTH row parities in order and are only visible in how they interleave (part 2 of the
    check): rows 2r and 2r+1 exchanged around a boundary."""
    k, h, w = 8, 496, 512
    yt, yp = synth.layered_pair(4, h, w, k, seed=77)
    for i, x in ((1, 7), (2, 200), (3, 511)):
        col = yt[i, :, x]
        r = int(np.argmax(col >= 3))                      # first row of class >= 3: rows r-1 | r differ
        a, b = col[r - 1], col[r]
        yt[i, r - 1, x], yt[i, r, x] = b, a               # adjacent exchange: each parity chain stays monotone
    got, _ = _run(yt, yp, k, cuda)
    np.testing.assert_array_equal(got, _expected(yt, yp))
    assert got[0] == 0 and all(got[1:] & 1)


def test_random_maps_are_never_certified(cuda):
    yt, yp = synth.random_pair(3, 124, 256, 8, seed=5)
    got, _ = _run(yt, yp, 8, cuda)
    assert (got == 3).all()
    z = np.zeros((2, 64, 128), np.uint8)                  # constant maps are sorted
    got, _ = _run(z, z, 4, cuda)
    assert (got == 0).all()
