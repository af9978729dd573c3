"""Seeded synthetic OCT label maps of the shapes BASELINE.json names.

There is no network and the reference ships no data, so every parity case and benchmark runs
on these.  All outputs are C-contiguous ``uint8 [N, H, W]`` pairs ``(y_true, y_pred)``.

* ``layered_pair``        -- retinal-layer maps: K-1 wavy boundaries per B-scan, label = number of
                             boundaries at or above the pixel; the prediction jitters every
                             boundary per column and optionally salts random labels.
* ``lesion_pair``         -- RETOUCH-style fluid masks: a few ellipses per class per slice.
* ``random_pair``         -- uniform random labels (the histogram-contention worst case).
* ``layered_pair_device`` -- the layered generator with torch ops on a CUDA device, for sets too
                             large to upload (cfg4: 100k B-scans).
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (N, H, W, K)  -- BASELINE.json "configs" in order
    "cfg1_duke_dme": (61, 496, 768, 8),
    "cfg2_hcms": (49, 496, 1024, 10),
    "cfg3_retouch": (128, 512, 512, 4),
    "cfg4_suite": (100_000, 496, 512, 8),
    "cfg5_volume3d": (1, 1024, 1024 * 128, 11),   # (1024, 1024, 128) volume, see surface3d
}


def _layer_boundaries(rng, n, h, w, k, amp=0.04, min_gap=2):
    """float [n, k-1, w] boundary rows, ordered top to bottom with at least ``min_gap`` px between."""
    nb = k - 1
    x = np.arange(w, dtype=np.float64)[None, None, :]
    base = h * (0.18 + 0.64 * (np.arange(nb, dtype=np.float64) + 0.5) / nb)[None, :, None]
    lam = rng.uniform(0.6 * w, 2.5 * w, size=(n, 1, 1))
    phase = rng.uniform(0, 2 * np.pi, size=(n, 1, 1))
    tilt = rng.uniform(-0.03, 0.03, size=(n, 1, 1)) * (x - w / 2)
    wave = amp * h * np.sin(2 * np.pi * x / lam + phase)
    ripple = rng.uniform(0.0, 0.012 * h, size=(n, nb, 1)) * np.sin(
        2 * np.pi * x / rng.uniform(0.08 * w, 0.3 * w, size=(n, nb, 1)) + rng.uniform(0, 6.28, size=(n, nb, 1)))
    b = base + wave + tilt + ripple
    return _order_boundaries(b, h, min_gap)


def _order_boundaries(b, h, min_gap):
    b = np.sort(b, axis=1)
    nb = b.shape[1]
    gap = np.arange(nb, dtype=np.float64)[None, :, None] * min_gap
    b = np.maximum.accumulate(b - gap, axis=1) + gap          # enforce the gap, keep order
    return np.clip(b, 1, h - 1)


def _rasterise(b_int, h):
    """label[n, y, x] = #{k : b_k(x) <= y} for integer boundaries [n, nb, w]."""
    n, nb, w = b_int.shape
    y = np.arange(h, dtype=np.int32)[None, :, None]
    lab = np.zeros((n, h, w), dtype=np.uint8)
    for k in range(nb):
        lab += (b_int[:, k, None, :] <= y).astype(np.uint8)
    return lab


def layered_pair(n, h, w, num_classes, seed, jitter=1.5, noise=0.0, min_gap=2):
    """(y_true, y_pred) layered label maps.  ``noise`` = fraction of prediction pixels replaced by a
    uniform random class (fills the off-diagonal confusion bins; breaks column ordering)."""
    rng = np.random.default_rng(seed)
    bt = _layer_boundaries(rng, n, h, w, num_classes, min_gap=min_gap)
    bt_i = np.rint(bt).astype(np.int32)
    bp = bt + rng.uniform(-2.0, 2.0, size=(n, num_classes - 1, 1)) + rng.normal(0.0, jitter, size=bt.shape)
    bp_i = np.rint(_order_boundaries(bp, h, min_gap)).astype(np.int32)
    y_true = _rasterise(bt_i, h)
    y_pred = _rasterise(bp_i, h)
    if noise > 0:
        salt = rng.random(size=y_pred.shape) < noise
        y_pred[salt] = rng.integers(0, num_classes, size=int(salt.sum()), dtype=np.uint8)
    return np.ascontiguousarray(y_true), np.ascontiguousarray(y_pred)


def random_pair(n, h, w, num_classes, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, num_classes, size=(n, h, w), dtype=np.uint8),
            rng.integers(0, num_classes, size=(n, h, w), dtype=np.uint8))


def _paint_ellipses(lab, cls, cy, cx, ry, rx):
    h, w = lab.shape
    yy, xx = np.ogrid[:h, :w]
    for a, b, c, d in zip(cy, cx, ry, rx):
        lab[((yy - a) / c) ** 2 + ((xx - b) / d) ** 2 <= 1.0] = cls


def lesion_pair(n, h, w, num_classes, seed, single_blob_interior=True):
    """RETOUCH-style masks: class 0 background, classes 1..K-1 fluid blobs (ellipses).

    ``single_blob_interior=True``: exactly one blob per class, clear of the border and of the other
    classes, so contour ``[0]`` is the whole (closed) outline.  ``False``: 1-3 blobs per class that
    may overlap or touch the border (exercises the ``[0]`` selection and open contours).  Every
    class is present in both maps of every slice (else the reference raises IndexError)."""
    rng = np.random.default_rng(seed)
    y_true = np.zeros((n, h, w), dtype=np.uint8)
    y_pred = np.zeros((n, h, w), dtype=np.uint8)
    nf = num_classes - 1
    for i in range(n):
        for c in range(1, num_classes):
            if single_blob_interior:
                # one blob per class in its own horizontal band of the slice
                band = h / nf
                ry = rng.uniform(0.08, 0.30) * band
                rx = rng.uniform(0.03, 0.16) * w
                cy = np.array([band * (c - 0.5) + rng.uniform(-0.1, 0.1) * band])
                cx = np.array([rng.uniform(rx + 6, w - rx - 6)])
                ry, rx = np.array([max(ry, 3.0)]), np.array([max(rx, 3.0)])
            else:
                m = int(rng.integers(1, 4))
                ry = rng.uniform(8, 80, size=m) * h / 512
                rx = rng.uniform(8, 80, size=m) * w / 512
                cy = rng.uniform(0, h, size=m)
                cx = rng.uniform(0, w, size=m)
            _paint_ellipses(y_true[i], c, cy, cx, ry, rx)
            dy, dx = rng.uniform(-3, 3, size=cy.shape), rng.uniform(-3, 3, size=cy.shape)
            sy, sx = rng.uniform(0.9, 1.1, size=cy.shape), rng.uniform(0.9, 1.1, size=cy.shape)
            _paint_ellipses(y_pred[i], c, cy + dy, cx + dx, ry * sy, rx * sx)
        if not single_blob_interior:
            for lab in (y_true[i], y_pred[i]):           # later classes may have painted one out
                for c in range(1, num_classes):
                    if not (lab == c).any():
                        lab[2 + 3 * c:5 + 3 * c, 2:6] = c
    return y_true, y_pred


def layered_pair_device(n, h, w, num_classes, seed, device, jitter=1.5, noise=0.0, chunk=4096):
    """Layered maps generated with torch ops on ``device`` (same construction as ``layered_pair``;
    a different random stream).  Returns two ``uint8 [n, h, w]`` torch tensors on ``device``."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    nb = num_classes - 1
    y_true = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    y_pred = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    x = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
    yy = torch.arange(h, device=device, dtype=torch.int32)[None, :, None]
    base = h * (0.18 + 0.64 * (torch.arange(nb, device=device, dtype=torch.float32) + 0.5) / nb)[None, :, None]
    gap = torch.arange(nb, device=device, dtype=torch.float32)[None, :, None] * 2

    def uni(lo, hi, shape):
        return lo + (hi - lo) * torch.rand(shape, generator=g, device=device)

    def order(b):
        b = torch.sort(b, dim=1).values
        b = torch.cummax(b - gap, dim=1).values + gap
        return b.clamp(1, h - 1)

    def raster(bi, out):
        acc = torch.zeros((bi.shape[0], h, w), dtype=torch.uint8, device=device)
        for k in range(nb):
            acc += (bi[:, k, None, :] <= yy).to(torch.uint8)
        out.copy_(acc)

    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        lam = uni(0.6 * w, 2.5 * w, (m, 1, 1))
        wave = 0.04 * h * torch.sin(2 * torch.pi * x / lam + uni(0, 6.28, (m, 1, 1)))
        tilt = uni(-0.03, 0.03, (m, 1, 1)) * (x - w / 2)
        ripple = uni(0, 0.012 * h, (m, nb, 1)) * torch.sin(
            2 * torch.pi * x / uni(0.08 * w, 0.3 * w, (m, nb, 1)) + uni(0, 6.28, (m, nb, 1)))
        bt = order(base + wave + tilt + ripple)
        bp = order(bt + uni(-2, 2, (m, nb, 1)) + jitter * torch.randn(bt.shape, generator=g, device=device))
        raster(torch.round(bt).to(torch.int32), y_true[s:s + m])
        raster(torch.round(bp).to(torch.int32), y_pred[s:s + m])
        if noise > 0:
            salt = torch.rand((m, h, w), generator=g, device=device) < noise
            rnd = torch.randint(0, num_classes, (m, h, w), generator=g, device=device, dtype=torch.uint8)
            y_pred[s:s + m] = torch.where(salt, rnd, y_pred[s:s + m])
    return y_true, y_pred


def ragged_pair_device(n, h, w, num_classes, seed, device, frac=0.5, chunk=4096):
    """Layered pair whose PREDICTION has ragged boundaries, the way a network's argmax does: every predicted pixel takes,
    with probability ``frac / 2`` each, the label of its upper or of its lower neighbour.  Interiors stay intact (the
    neighbours agree there); along every boundary the path gets one-pixel teeth, overhangs and detached single pixels,
    so most columns of the prediction are out of class order and most predicted contours are not height functions."""
    import torch
    y_true, y_pred = layered_pair_device(n, h, w, num_classes, seed, device, chunk=chunk)
    g = torch.Generator(device=device)
    g.manual_seed(seed + 12345)
    for s in range(0, n, chunk):
        p = y_pred[s:s + chunk]
        r = torch.rand(p.shape, generator=g, device=device)
        up = torch.cat([p[:, :1], p[:, :-1]], dim=1)
        down = torch.cat([p[:, 1:], p[:, -1:]], dim=1)
        y_pred[s:s + chunk] = torch.where(r < frac / 2, up, torch.where(r < frac, down, p))
    return y_true, y_pred


def layered_volume_pair(d0, d1, d2, num_classes, seed, jitter=1.0):
    """(vol_true, vol_pred) ``uint8 [d0, d1, d2]`` label volumes with depth along axis 0: K-1 smooth
    surfaces ``b_k(x, z)`` stacked top to bottom; the prediction shifts and jitters every surface
    (BASELINE config 5: ``1024 x 1024 x 128``, 11 classes)."""
    rng = np.random.default_rng(seed)
    nb = num_classes - 1
    x = np.arange(d1, dtype=np.float64)[None, :, None]
    z = np.arange(d2, dtype=np.float64)[None, None, :]
    base = d0 * (0.15 + 0.7 * (np.arange(nb, dtype=np.float64) + 0.5) / nb)[:, None, None]
    wave = 0.04 * d0 * np.sin(2 * np.pi * x / rng.uniform(0.7 * d1, 2.0 * d1) + rng.uniform(0, 6.28)) * \
        np.cos(2 * np.pi * z / rng.uniform(1.0 * d2, 3.0 * d2) + rng.uniform(0, 6.28))
    bt = np.sort(base + wave, axis=0)
    gap = np.arange(nb, dtype=np.float64)[:, None, None] * 2
    bt = np.clip(np.maximum.accumulate(bt - gap, axis=0) + gap, 1, d0 - 1)
    bp = bt + rng.uniform(-1.5, 1.5, size=(nb, 1, 1)) + rng.normal(0.0, jitter, size=bt.shape)
    bp = np.clip(np.maximum.accumulate(np.sort(bp, axis=0) - gap, axis=0) + gap, 1, d0 - 1)
    y = np.arange(d0, dtype=np.int32)[:, None, None]

    def raster(b):
        vol = np.zeros((d0, d1, d2), dtype=np.uint8)
        bi = np.rint(b).astype(np.int32)
        for k in range(nb):
            vol += (bi[k][None] <= y).astype(np.uint8)
        return vol
    return raster(bt), raster(bp)
