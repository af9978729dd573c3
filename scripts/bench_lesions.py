#!/usr/bin/env python
"""cfg3 at scale: the full suite on a large batch of lesion (blob) masks -- every contour is closed, so the layered
path rejects all of them and the general walk + counting-sort search run.  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch      # noqa: E402


def main():
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth
    dev = torch.device("cuda", 0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    yt, yp = synth.lesion_pair(256, 512, 512, 4, seed=3003, single_blob_interior=False)
    reps = (n + 255) // 256
    yt = torch.from_numpy(yt).to(dev).repeat(reps, 1, 1)[:n].contiguous()
    yp = torch.from_numpy(yp).to(dev).repeat(reps, 1, 1)[:n].contiguous()
    timers = {}
    for _ in range(3):
        suite.evaluate(yt, yp, 4).totals
    torch.cuda.synchronize()
    timers.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        suite.evaluate(yt, yp, 4, timers=timers).totals
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    per = {k: sum(x.elapsed_time(y) for x, y in v) / len(v) for k, v in timers.items()}
    print(json.dumps({"workload": f"cfg3 at scale: {n} x 512 x 512 lesion masks, K=4, full suite", "ms_per_batch": ms,
                      "slices_per_s": n / (ms / 1e3), "kernel_ms": per}))


if __name__ == "__main__":
    main()
