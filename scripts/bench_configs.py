#!/usr/bin/env python
"""Timings of BASELINE.json configs 1-3 (single volumes) on one GPU, device-resident, CUDA events.

    python scripts/bench_configs.py

cfg1: confusion + Dice/IoU/... on a Duke-DME-sized volume (61 x 496 x 768, 8 classes)
cfg2: boundary pixel error + thickness biomarkers on an HC-MS-shaped volume (49 x 496 x 1024, 9 boundaries)
cfg3: contour metrics on RETOUCH-style lesion masks (128 x 512 x 512, 4 classes)
One JSON line per config: ms per volume, B-scans/s, label GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch      # noqa: E402


def timed(fn, reps=20, warm=3):
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


def main():
    from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib, suite, synth
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    out = []
    # cfg1
    yt, yp = synth.layered_pair(61, 496, 768, 8, seed=1001, noise=0.01)
    yt, yp = torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)
    ms = timed(lambda: suite.derive_on_device(suite.label_pass(yt, yp, 8, counts=True, columns=False), None, 61))
    out.append(("cfg1 confusion + count metrics, 61x496x768 K=8", 61, 496 * 768, ms,
                "fast" if lib.octm_label_pass_path(496, 768, 8, yt.data_ptr(), yp.data_ptr()) else "generic"))
    # cfg2
    yt, yp = synth.layered_pair(49, 496, 1024, 10, seed=2002)
    yt, yp = torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)
    ms = timed(lambda: suite.label_pass(yt, yp, 10, counts=False, columns=True, boundaries=True))
    out.append(("cfg2 boundary error + thickness, 49x496x1024 K=10", 49, 496 * 1024, ms,
                "fast" if lib.octm_label_pass_path(496, 1024, 10, yt.data_ptr(), yp.data_ptr()) else "generic"))
    # cfg3
    yt, yp = synth.lesion_pair(128, 512, 512, 4, seed=3003, single_blob_interior=False)
    yt, yp = torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)
    ms = timed(lambda: suite.evaluate(yt, yp, 4).totals, reps=10)
    out.append(("cfg3 full suite incl. contour metrics, 128x512x512 K=4 lesions", 128, 512 * 512, ms, "fast"))
    for name, n, px, ms, path in out:
        print(json.dumps({"workload": name, "ms_per_volume": ms, "bscans_per_s": n / (ms / 1e3),
                          "label_gb_per_s": 2 * n * px / (ms / 1e3) / 1e9, "label_pass_path": path}))


if __name__ == "__main__":
    main()
