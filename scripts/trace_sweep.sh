#!/bin/bash
# Tuning sweep: resident CTAs per SM of the contour trace kernel.
mkdir -p gpurun_out
: > gpurun_out/trace_sweep.txt
for c in 2 3 4 5 6 8 10 12; do
  OCTM_TRACE_CTAS=$c python bench.py --items 8192 --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ctas/SM $c', 'trace_ms', round(d['kernel_ms_per_step']['contour_trace'],4))" >> gpurun_out/trace_sweep.txt
done
cat gpurun_out/trace_sweep.txt
