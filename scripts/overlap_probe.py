#!/usr/bin/env python
"""Experiment: do the label pass (HBM + issue) and the fused contour kernel (issue, no DRAM) gain from running
CONCURRENTLY on the same SMs?  Two streams, two independent batches, resident CTAs per SM capped by the environment
variables OCTM_LP_CTAS / OCTM_LD_CTAS.  Prints sequential and concurrent times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth

dev = torch.device("cuda", 0)
n, h, w, k = 8192, 496, 512, 8
a0, p0 = synth.layered_pair_device(n, h, w, k, seed=1, device=dev)
a1, p1 = synth.layered_pair_device(n, h, w, k, seed=2, device=dev)
lp1 = suite.label_pass(a1, p1, k, seeds=True, boundaries=True, certify=True)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()

def label():
    return suite.label_pass(a0, p0, k, seeds=True, boundaries=True, certify=True)

def contour():
    return suite.contour_pass(a1, p1, k, lp1.first_pos, boundaries=(lp1.bnd_true, lp1.bnd_pred), unsorted=lp1.unsorted, check_overflow=False)

def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

def both():
    with torch.cuda.stream(sA):
        label()
    with torch.cuda.stream(sB):
        contour()

tl, tc = timed(label), timed(contour)
tb = timed(both)
print("LP_CTAS=%s LD_CTAS=%s  label %.3f ms  contour %.3f ms  sum %.3f  concurrent %.3f  gain %.1f %%" % (
    os.environ.get("OCTM_LP_CTAS", "-"), os.environ.get("OCTM_LD_CTAS", "-"), tl, tc, tl + tc, tb, 100 * (1 - tb / (tl + tc))))
