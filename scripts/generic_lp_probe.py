#!/usr/bin/env python
"""Label-pass throughput on shapes the strip kernel does not take (K > 8, ragged widths): the generic kernel at scale.

    python scripts/generic_lp_probe.py [items]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch      # noqa: E402
from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


SHAPES = ((496, 1024, 10, n), (496, 1024, 10, 49), (496, 500, 8, n), (496, 512, 8, n))
if os.environ.get("GLP_ONLY"):
    SHAPES = (SHAPES[int(os.environ["GLP_ONLY"])],)
for (H, W, K, m) in SHAPES:
    yt, yp = synth.layered_pair_device(m, H, W, K, seed=5, device=dev, noise=0.0)
    for kw, name in ((dict(counts=True, columns=True), "counts+columns"),
                     (dict(counts=True, columns=True, seeds=True, boundaries=True, certify=True), "suite call")):
        ms = timed(lambda: suite.label_pass(yt, yp, K, **kw))
        print(f"{m:6d} x {H}x{W} K={K:2d} {name:15s}: {ms:8.3f} ms  {2 * m * H * W / ms / 1e6:8.1f} GB/s  {m / ms * 1e3:10.0f} B-scans/s")
