#!/usr/bin/env python
"""BASELINE config 5: 3-D surface-distance metrics on one synthetic 1024 x 1024 x 128 volume pair, 11 classes.

    python scripts/bench_cfg5.py                      # one GPU, all 22 (class, direction) units
    torchrun --nproc-per-node N scripts/bench_cfg5.py # units sharded over N GPUs, one small all-reduce
Prints one JSON line (time per volume pair, max over ranks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import torch            # noqa: E402


def main():
    from retinal_oct_image_segmentation_via_deep_learning_b200 import dist as odist
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite, synth
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    d0, d1, d2, k = (int(x) for x in os.environ.get("CFG5_SHAPE", "1024,1024,128,11").split(","))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    vt, vp = synth.layered_volume_pair(d0, d1, d2, k, seed=5005)
    gen_s = time.perf_counter() - t0
    vt, vp = torch.from_numpy(vt).to(dev), torch.from_numpy(vp).to(dev)
    for _ in range(1):
        odist.surface_distance_3d_sharded(vt, vp, k, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 2
    a.record()
    for _ in range(steps):
        ints = odist.surface_distance_3d_sharded(vt, vp, k, rank, world)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    m = suite.surface_metrics_3d(ints)
    if rank == 0:
        vox = d0 * d1 * d2
        print(json.dumps({"workload": f"cfg5: 3-D surface distances, {d0}x{d1}x{d2}, {k} classes, {2 * k} EDT units",
                          "n_gpus": world, "ms_per_volume_pair": float(ms.item()),
                          "voxel_transforms_per_s": 2 * k * vox / (float(ms.item()) / 1e3),
                          "host_generation_s": gen_s,
                          "hausdorff_distance": [float(x) for x in m["hausdorff_distance"]],
                          "assd": [float(x) for x in m["assd"]]}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
