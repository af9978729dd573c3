#!/usr/bin/env python
"""Benchmark: B-scans/s of the full metric suite on synthetic 496x512 8-class label maps (cfg4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--items M]

A "step" is one pass of the whole suite (fused label pass + contour trace + contour distances +
float64 epilogue) over this rank's batch of B-scans, inputs resident in HBM.  One process per GPU
(torchrun for N>1), B-scans are sharded across ranks with no data-path collective (weak scaling:
--items per GPU); a single small NCCL all-reduce merges the per-class counts for the dataset-level
numbers and is part of the timed step.  Steps are enqueued asynchronously and timed with CUDA events;
"e2e" is the same suite through suite.evaluate_host on pinned host arrays (H2D inside the timed region,
metrics read back).  Rank 0 prints ONE JSON line.

--impl reference times the reference's CPU path (the numpy oracle port of Metrics/*.py, per-class
per-function Python loops, all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, K = 496, 512, 8
BYTES_PER_BSCAN = 2 * H * W                      # algorithmic (compulsory) label bytes, SURVEY.md 8(d)
WORKLOAD = "cfg4: full metric suite, synthetic layered 496x512 B-scans, 8 classes"
# DRAM traffic of label_pass_fast per B-scan from the ncu --set full capture of this workload
# (profiles/r1_v10_ncu_summary.md: dram__bytes_read 1.041599 GB + dram__bytes_write 11.639 MB for 2048 B-scans)
NCU_TRAFFIC_BYTES_PER_BSCAN = (1.041599e9 + 11.638528e6) / 2048


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs.

    In-process NVML (nvidia_ml_py) from a thread, one cheap query every 20 ms: an `nvidia-smi -lms` child
    process holds the driver for milliseconds per query and was measured to stretch a 17 ms step to 18-27 ms.
    Falls back to that child process only if NVML cannot be imported."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.index, self.thread, self.stop_flag = [], None, index, None, False
        self.max_mhz, self.source = None, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        if os.environ.get("OCTM_BENCH_NOCLOCKS") == "1":      # diagnosis only: is the sampler disturbing the run?
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = [getattr(pynvml, n, 0) for n in ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksEventReasonHwThermalSlowdown",
                                                    "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksEventReasonSwPowerCap")]
            if not all(bits):
                bits = [0x8, 0x40, 0x20, 0x4]            # NVML ABI values of the four reasons above

            def loop():
                while not self.stop_flag:
                    try:
                        mhz = int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            r = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.perf_counter(), mhz, [bool(r & b) for b in bits]))
                    except Exception:
                        pass
                    time.sleep(0.02)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if c and c[0].isdigit():
                if len(c) > 1 and c[1].isdigit():
                    self.max_mhz = max(self.max_mhz or 0, int(c[1]))
                self.rows.append((time.perf_counter(), int(c[0]), [len(c) > 2 + i and c[2 + i] == "Active" for i in range(4)]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken inside [t0, t1] (perf_counter; all samples if none fall inside).
        Sampling starts BEFORE the warm-up so that its own initialisation is not in the timed region."""
        if self.thread is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.05)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for r in self.rows if t0 is None or t0 <= r[0] <= t1 + 0.05] or list(self.rows)
        sm = sorted(r[1] for r in rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2][i] for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------------ CPU arm
def _cpu_one(args):
    seed, contours = args
    from oracle import labelmap_oracle as lo
    from retinal_oct_image_segmentation_via_deep_learning_b200 import synth
    yt, yp = synth.layered_pair(1, H, W, K, seed=seed)
    t0 = time.perf_counter()
    lo.score_bscan(yt[0], yp[0], K, contours=contours)
    return time.perf_counter() - t0


def cpu_baseline(n_scans, cores, contours=True):
    """Oracle port of the reference (per-class, per-function numpy calls) on `cores` processes."""
    import multiprocessing as mp
    t0 = time.perf_counter()
    if cores == 1:
        per = [_cpu_one((5000 + i, contours)) for i in range(n_scans)]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            per = pool.map(_cpu_one, [(5000 + i, contours) for i in range(n_scans)])
    wall = time.perf_counter() - t0
    return n_scans / wall, sum(per) / len(per)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_scans = max(cores, 1)                      # one B-scan per core per step (~10-20 s of CPU work)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(min(cores, n_scans), cores)
    t0 = time.perf_counter()
    rates = [cpu_baseline(n_scans, cores)[0] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = n_scans * args.steps / wall
    sample = f"{n_scans} B-scans per step (one per core), {args.steps} steps, full suite incl. contour metrics"
    print(json.dumps({
        "impl": "reference", "metric": "bscans_per_sec_full_metric_suite", "value": value, "unit": "B-scans/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "height": H, "width": W, "num_classes": K},
        "cpu_baseline": {"value": value, "unit": "B-scans/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "B-scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "per_step_rates": rates,
    }))


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from retinal_oct_image_segmentation_via_deep_learning_b200 import _lib, suite, synth
    from retinal_oct_image_segmentation_via_deep_learning_b200 import dist as odist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    n = args.items
    yt, yp = synth.layered_pair_device(n, H, W, K, seed=4004 + rank, device=dev, noise=args.noise)
    torch.cuda.synchronize()

    timers = {}

    trace = os.environ.get("OCTM_BENCH_TRACE") == "1"

    def step():
        """One pass of the whole suite over this rank's batch, enqueued asynchronously: kernels, the totals
        kernel and the cross-rank all-reduce.  Results stay in HBM; they are read back after the timed region
        (a per-step read-back would put this host's scheduling jitter, not the GPU, on the clock)."""
        t0 = time.perf_counter()
        res = suite.evaluate(yt, yp, K, contours=not args.no_contours, timers=timers)
        pend = odist.dataset_totals_async(res, world)     # one small all-reduce on the device vector
        if trace:
            print(f"[rank {rank}] enqueue {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
        return pend

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)                      # let nvidia-smi finish initialising before anything is timed
    prev = None
    for _ in range(args.warmup):          # two evaluations in flight, like the timed loop: primes torch's allocator
        cur = step()
        if prev is not None:
            prev.result()
        prev = cur
    if prev is not None:
        prev.result()
    prev = cur = None
    barrier()
    timers.clear()
    gc.collect()
    gc.disable()          # a generational collection inside a ~15 ms step is a 10 ms host stall
    all_totals = torch.zeros((args.steps, int(_lib.load().octm_totals_len(K))), dtype=torch.float64, device=dev)
    launches0 = _lib.launch_count()
    t_begin = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pend = None
    for i in range(args.steps):
        pend = step()                     # the previous step's outputs go back to torch's caching allocator here:
        all_totals[i].copy_(pend.reduced) # no cudaMalloc in the timed region (one can stall the host for 50+ ms)
    ev1.record()
    barrier()
    tot = pend.result()
    tot_host = all_totals.cpu().numpy()
    assert all((tot_host[i] == tot_host[0]).all() for i in range(args.steps)), "steps disagree on the dataset totals"
    ms = ev0.elapsed_time(ev1)
    gc.enable()
    launches = _lib.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * args.steps / (ms / 1e3)

    # per-kernel device times (CUDA events on the launch stream, inside the timed region)
    kern = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in timers.items()}
    peak, peak_src = _peaks()
    lp_ms = kern.get("label_pass")
    roofline = None
    if lp_ms:
        achieved = n * BYTES_PER_BSCAN / (lp_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "label_pass_fast", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": n * NCU_TRAFFIC_BYTES_PER_BSCAN if not args.no_contours else None,
                    "traffic_source": "ncu dram__bytes_read+write per B-scan (profiles/r1_v10_ncu_summary.md) x items per launch",
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": n * BYTES_PER_BSCAN, "ms_per_launch": lp_ms,
                    "suite_compulsory_gbs_per_gpu": n * BYTES_PER_BSCAN * args.steps / (ms / 1e3) / 1e9}

    # end to end through the public host-array API: pinned host buffers, H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        m = min(n, args.e2e_items)
        ht, hp = yt[:m].cpu().pin_memory(), yp[:m].cpu().pin_memory()
        # this box's pinned host->device copy rate: the ceiling of the end-to-end number (boxes differ by 2-3x)
        probe_d = torch.empty_like(yt[:m])
        probe_d.copy_(ht, non_blocking=True)
        torch.cuda.synchronize()
        tp = time.perf_counter()
        for _ in range(2):
            probe_d.copy_(ht, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = 2 * ht.numel() / (time.perf_counter() - tp) / 1e9
        del probe_d
        for _ in range(max(1, min(args.warmup, 2))):
            suite.evaluate_host(ht, hp, K, contours=not args.no_contours, device=dev).metrics()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 6))
        d2h = 0
        for _ in range(e2e_steps):
            r = suite.evaluate_host(ht, hp, K, contours=not args.no_contours, device=dev)
            d2h = sum(v.nbytes for v in r.metrics().values()) + r.totals_host().nbytes
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * m * e2e_steps / float(t.item()), "unit": "B-scans/s",
               "h2d_bytes_per_step": int(m * BYTES_PER_BSCAN), "d2h_bytes_per_step": int(d2h),
               "items_per_step_per_gpu": m, "steps": e2e_steps,
               "pinned_h2d_gbs_this_box": round(h2d_gbs, 1),
               "h2d_gbs_achieved": round(world * m * e2e_steps * BYTES_PER_BSCAN / float(t.item()) / 1e9 / world, 1)}

    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None   # device-timed + e2e regions
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            ns = max(1, min(cores, 64))
            rate, per = cpu_baseline(ns, min(cores, ns))
            cpu = {"value": rate, "unit": "B-scans/s", "cores": min(cores, ns), "kind": "port",
                   "sample": f"{ns} B-scans of the same workload, oracle port of Metrics/*.py called per class "
                             f"per function; {per:.1f} s per B-scan per core"}
        print(json.dumps({
            "metric": "bscans_per_sec_full_metric_suite", "value": value, "unit": "B-scans/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD + (f", {args.noise:g} of the predicted pixels randomised" if args.noise else ""),
                       "items_per_gpu": n, "height": H, "width": W, "num_classes": K,
                       "contours": not args.no_contours, "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2"
                       % (n * BYTES_PER_BSCAN / 1e9), "sharding": f"items x{world}, one NCCL all-reduce of totals",
                       "results": "left in HBM during the timed region; every step's dataset totals are read back and compared after it"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "kernel_ms_per_step": kern,
            "dataset_dice": [float(x) for x in tot["dice_coefficient"]] if tot else None,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--items", type=int, default=16384, help="B-scans per GPU per step")
    ap.add_argument("--noise", type=float, default=0.0,
                    help="fraction of predicted pixels replaced by a random class (default 0: the contract's clean layered maps)")
    ap.add_argument("--e2e-items", type=int, default=4096)
    ap.add_argument("--no-contours", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
