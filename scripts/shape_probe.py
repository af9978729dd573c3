#!/usr/bin/env python
"""Per-kernel times of the full suite on one shape: python scripts/shape_probe.py items H W K [noise]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch      # noqa: E402
from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib, suite, synth   # noqa: E402

n, H, W, K = (int(v) for v in sys.argv[1:5])
noise = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
dev = torch.device("cuda", 0)
if noise < 0:          # negative "noise": ragged predicted boundaries
    yt, yp = synth.ragged_pair_device(n, H, W, K, seed=7003, device=dev)
else:
    yt, yp = synth.layered_pair_device(n, H, W, K, seed=7003, device=dev, noise=noise)
for _ in range(3):
    suite.evaluate(yt, yp, K).totals_host()
with _lib.kernel_profile() as prof:
    for _ in range(2):
        suite.evaluate(yt, yp, K).totals_host()
tot = 0.0
for name, (launches, ms) in sorted(prof.kernels.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name:32s} x{launches // 2:<3d} {ms / 2:8.3f} ms")
    tot += ms / 2
print(f"{n} x {H}x{W} K={K} noise {noise:g}: {tot:.3f} ms of kernels per step -> {n / tot * 1e3:,.0f} B-scans/s")
