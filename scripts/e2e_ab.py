"""Diagnostic: evaluate_host end-to-end rate of the tree in the current directory (run from two checkouts to compare)."""
import sys, time
import torch
sys.path.insert(0, ".")
from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth
dev = torch.device("cuda", 0)
yt, yp = synth.layered_pair_device(4096, 496, 512, 8, seed=1, device=dev)
ht, hp = yt.cpu().pin_memory(), yp.cpu().pin_memory()
for rep in range(2):
    for contours in (True, False):
        for _ in range(2):
            suite.evaluate_host(ht, hp, 8, contours=contours, device=dev).metrics()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            suite.evaluate_host(ht, hp, 8, contours=contours, device=dev).metrics()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print(sys.argv[1] if len(sys.argv) > 1 else "", "contours", contours, "e2e B-scans/s %.0f" % (4096 / dt), "GB/s %.1f" % (4096 * 2 * 496 * 512 / dt / 1e9), flush=True)
