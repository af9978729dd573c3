#!/bin/bash
# Tuning sweep of the label-pass ring geometry (rows per stage x stages) on the B200 box.
mkdir -p gpurun_out
: > gpurun_out/lp_sweep.txt
for g in "8 2" "8 3" "8 4" "12 2" "12 3" "16 2" "16 3" "24 2" "32 2"; do set -- $g
  OCTM_LP_ROWS=$1 OCTM_LP_STAGES=$2 python bench.py --items 8192 --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rows $1 stages $2', 'label_pass_ms', round(d['kernel_ms_per_step']['label_pass'],4), 'frac', round(d['roofline']['frac'],4))" >> gpurun_out/lp_sweep.txt
done
cat gpurun_out/lp_sweep.txt
