#!/bin/bash
# Tuning sweep of the label-pass ring geometry (rows per stage x stages) on the B200 box.
mkdir -p gpurun_out
: > gpurun_out/lp_sweep.txt
for rows in 8 12 16 24; do for st in 2 3 4; do
  OCTM_LP_ROWS=$rows OCTM_LP_STAGES=$st python bench.py --items 8192 --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rows $rows stages $st', 'label_pass_ms', round(d['kernel_ms_per_step']['label_pass'],4), 'frac', round(d['roofline']['frac'],4))" >> gpurun_out/lp_sweep.txt
done; done
cat gpurun_out/lp_sweep.txt
