"""Diagnostic: pinned host->device copy bandwidth of this box next to the end-to-end rate of evaluate_host."""
import sys, time
import torch
sys.path.insert(0, ".")
from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth

dev = torch.device("cuda", 0)
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
print("pinned H2D GB/s", 4 * (1 << 30) / (time.perf_counter() - t0) / 1e9)
yt, yp = synth.layered_pair_device(4096, 496, 512, 8, seed=1, device=dev)
ht, hp = yt.cpu().pin_memory(), yp.cpu().pin_memory()
for contours in (True, False):
    for _ in range(2):
        suite.evaluate_host(ht, hp, 8, contours=contours, device=dev).metrics()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        suite.evaluate_host(ht, hp, 8, contours=contours, device=dev).metrics()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("contours", contours, "e2e B-scans/s", 4096 / dt, "GB/s", 4096 * 2 * 496 * 512 / dt / 1e9)
