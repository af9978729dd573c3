#!/usr/bin/env python
"""Experiment: scores -> labels -> label pass with the predicted labels kept L2-resident by chunking
(argmax of a chunk into a small reused buffer, label pass of the chunk right behind it), against argmax alone and against
the unchunked two-kernel path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth

dev = torch.device("cuda", 0)
n, k, h, w = 4096, 8, 496, 512
scores = torch.randn((n, k, h, w), device=dev, dtype=torch.float16)
yt = torch.randint(0, k, (n, h, w), device=dev, dtype=torch.uint8)
lab = torch.empty((n, h, w), dtype=torch.uint8, device=dev)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

t_arg = timed(lambda: suite.labels_from_scores(scores, out=lab))
def two():
    suite.labels_from_scores(scores, out=lab)
    suite.label_pass(yt, lab, k, counts=True, columns=True)
t_two = timed(two)
gb = scores.numel() * 2 / 1e9
print("argmax alone %.3f ms (%.2f TB/s of scores)   argmax + label pass %.3f ms (+%.1f %%)" % (t_arg, gb / t_arg, t_two, 100 * (t_two / t_arg - 1)))
for c in (64, 128, 256, 512):
    bufs = [torch.empty((c, h, w), dtype=torch.uint8, device=dev) for _ in range(2)]
    def chunked():
        for i, s in enumerate(range(0, n, c)):
            e = min(n, s + c)
            b = bufs[i & 1][:e - s]
            suite.labels_from_scores(scores[s:e], out=b)
            suite.label_pass(yt[s:e], b, k, counts=True, columns=True)
    t = timed(chunked)
    print("chunk %4d items (%5.1f MB of labels): %.3f ms (+%.1f %% over argmax alone)" % (c, c * h * w / 1e6, t, 100 * (t / t_arg - 1)))
