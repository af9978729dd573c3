"""Diagnostic: host nibble packer rate, packed H2D rate and evaluate_host with and without packing on this box."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib, suite, synth

dev = torch.device("cuda", 0)
lib = _lib.load()
print("host cores", os.cpu_count())
yt, yp = synth.layered_pair_device(4096, 496, 512, 8, seed=1, device=dev)
ht, hp = yt.cpu().pin_memory(), yp.cpu().pin_memory()
n = ht.numel()
dst = torch.empty((n + 1) // 2, dtype=torch.uint8).pin_memory()
for threads in (0, 8, 16, 32):
    lib.octm_host_pack_nibbles(ht.data_ptr(), dst.data_ptr(), n, threads)
    t0 = time.perf_counter()
    for _ in range(3):
        lib.octm_host_pack_nibbles(ht.data_ptr(), dst.data_ptr(), n, threads)
    dt = (time.perf_counter() - t0) / 3
    print(f"pack threads={threads}: {n / dt / 1e9:.1f} GB/s of input")
for pack in (False, True):
    for _ in range(2):
        suite.evaluate_host(ht, hp, 8, device=dev, pack=pack).metrics()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        suite.evaluate_host(ht, hp, 8, device=dev, pack=pack).metrics()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("pack", pack, "e2e B-scans/s", round(4096 / dt), "input GB/s", round(4096 * 2 * 496 * 512 / dt / 1e9, 1))
