#!/usr/bin/env python
"""Label pass cost by feature set (16,384 cfg4 items): plain, + seeds, + certificate, + both."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib, suite, synth
dev = torch.device("cuda", 0)
n, h, w, k = 16384, 496, 512, 8
yt, yp = synth.layered_pair_device(n, h, w, k, seed=1, device=dev)
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
i64 = dict(dtype=torch.int64, device=dev); i32 = dict(dtype=torch.int32, device=dev)
counts, thick = torch.empty((n, k, k), **i64), torch.empty((n, k), **i64)
bsq, bab = torch.empty((n, k - 1), **i64), torch.empty((n, k - 1), **i64)
bt, bp = torch.empty((n, k - 1, w), **i32), torch.empty((n, k - 1, w), **i32)
fp, uns = torch.empty((n, 2, k), **i32), torch.empty((n,), **i32)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, f, u in (("plain", None, None), ("seeds", fp, None), ("certificate", None, uns), ("seeds+certificate", fp, uns)):
    call = lambda: _lib.call("octm_label_pass_sorted_u8", P(yt), P(yp), n, h, w, k, P(counts), P(thick), P(bsq), P(bab), P(bt), P(bp), P(f), P(u), st)
    for _ in range(2):
        call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5):
        call()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%-18s %.3f ms  %.2f TB/s" % (name, ms, n * 2 * h * w / ms / 1e9))
