#!/usr/bin/env python
"""Debug build only (OCTM_NVCC_EXTRA=-DOCTM_WALK_STATS): the population trace_kernel walks and what a step costs.

    OCTM_NVCC_EXTRA=-DOCTM_WALK_STATS python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force
    python scripts/walk_stats.py 2e-5 4096
"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib, suite, synth

noise = float(sys.argv[1]) if len(sys.argv) > 1 else 2e-5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lib = ctypes.CDLL(_lib.LIB_PATH)
f = lib.octm_debug_walk_stats
f.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
dev = torch.device("cuda:0")
yt, yp = synth.layered_pair_device(n, 496, 512, 8, seed=7001, device=dev, noise=noise)
out = (ctypes.c_ulonglong * 8)()
for rep in range(2):
    f(out, 1)
    r = suite.evaluate(yt, yp, 8)
    torch.cuda.synchronize()
    f(out, 0)
v = list(out)
print(f"noise {noise:g} items {n}: walks {v[0]} ({v[0] / n:.2f}/item) vertices {v[1]} max {v[2]}; short(<64) {v[3]}, long(>=512) {v[4]} "
      f"({v[4] / n:.3f}/item) mean {v[6] / max(v[4], 1):.0f} vertices, {v[5] / max(v[6], 1):.0f} cycles/vertex; slowest walk {v[7]} cycles")
