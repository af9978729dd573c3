#!/usr/bin/env python
"""One launch of every kernel family outside the cfg4 step, at sizes that fill the GPU, for an ncu capture:

    ncu --set full --clock-control none --import-source on -k regex:'argmax|auc_kernel|rasterise|near_|edt3_|layered_distance|trace_' \\
        -o gpurun_out/prof_aux python scripts/profile_kernels.py

Also prints each family's event-timed duration and throughput (run it once WITHOUT ncu for those numbers)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch      # noqa: E402


def timed(fn, reps=5, warm=2):
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
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth
    dev = torch.device("cuda", 0)
    quick = "--quick" in sys.argv          # under ncu: one launch each
    reps = dict(reps=1, warm=0) if quick else {}
    out = {}
    n, h, w, k = 2048, 496, 512, 8
    # argmax front end: [N, K, H, W] scores -> labels
    for dt, name in ((torch.float16, "f16"), (torch.float32, "f32")):
        s = torch.randn((512 if dt == torch.float32 else 1024, k, h, w), device=dev, dtype=dt)
        lab = torch.empty((s.shape[0], h, w), dtype=torch.uint8, device=dev)
        ms = timed(lambda: suite.labels_from_scores(s, out=lab), **reps)
        byts = s.numel() * s.element_size() + lab.numel()
        out[f"argmax_planes_kernel_{name}"] = {"ms": ms, "GB/s": byts / ms / 1e6, "bytes": byts}
        del s, lab
    # auc_score: one CTA per item
    yt = (torch.rand((512, h, w), device=dev) < 0.3).to(torch.uint8)
    sc = torch.rand((512, h, w), device=dev, dtype=torch.float32)
    ms = timed(lambda: suite.auc_scores(yt, sc), **reps)
    out["auc_kernel"] = {"ms": ms, "items/s": 512 / ms * 1e3, "GB/s": (yt.numel() + sc.numel() * 4) / ms / 1e6}
    del yt, sc
    # boundary rows -> label maps
    b = torch.sort(torch.rand((n, k - 1, w), device=dev) * (h - 40) + 20, dim=1).values.to(torch.int32)
    ms = timed(lambda: suite.labels_from_boundaries(b, h), **reps)
    out["rasterise_kernel"] = {"ms": ms, "labels_written_GB/s": n * h * w / ms / 1e6}
    # 3-D surface distances, near-field path
    vt, vp = synth.layered_volume_pair(512, 512, 128, 5, seed=5005)
    vt, vp = torch.from_numpy(vt).to(dev), torch.from_numpy(vp).to(dev)
    ms = timed(lambda: suite.surface_distance_3d(vt, vp, 5, units=(2, 4)), **reps)
    out["surface3d_near_512x512x128_2units"] = {"ms": ms, "voxels/s_per_unit": 2 * vt.numel() / ms * 1e3}
    del vt, vp
    # the suite on noisy predictions: second pass of the fused contour kernel, walk, vertex-list search
    a, p = synth.layered_pair_device(n, h, w, k, seed=7, device=dev, noise=2e-4)
    ms = timed(lambda: suite.evaluate(a, p, k).totals, **reps)
    out["suite_noise_2e-4"] = {"ms": ms, "B-scans/s": n / ms * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
