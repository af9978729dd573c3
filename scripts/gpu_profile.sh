#!/bin/bash
# Run on the B200 box (through gpurun): plain bench, then the ncu launch list and one full capture
# of each hot kernel.  Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
SMALL="--items 2048 --steps 2 --warmup 1 --no-e2e --no-cpu"
python bench.py $SMALL > gpurun_out/bench_small_$TAG.json 2> gpurun_out/bench_small_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $SMALL > gpurun_out/ncu_launches_$TAG.log 2>&1
python bench.py $SMALL > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'label_pass_fast|trace_layered_kernel|trace_kernel|distance_column_kernel|distance_search_kernel|distance_select_kernel|derive_kernel' -s 6 -c 6 \
    -o gpurun_out/prof_$TAG python bench.py $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
