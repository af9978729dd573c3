#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) as markdown: one block of headline metrics per kernel launch.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls] > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    stalls = "--stalls" in sys.argv
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    for r in data:
        print(f"## {r[col['Kernel Name']]}  (launch id {r[col['ID']]})\n")
        for m in METRICS:
            if m in col:
                print(f"- `{m}` = {r[col[m]]} {units[col[m]]}")
        if stalls:
            st = []
            for h, i in col.items():
                if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled"):
                    try:
                        st.append((float(r[i].replace(",", "")), h))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            for v, h in st[:6]:
                print(f"- stall `{h}` = {v:.3f}")
        print()


if __name__ == "__main__":
    main()
