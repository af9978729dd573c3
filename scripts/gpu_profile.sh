#!/bin/bash
# Run on the B200 box (through gpurun): plain bench, then the ncu launch list and one full capture
# of the label-pass and distance kernels.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
SMALL="--items 2048 --steps 2 --warmup 1 --no-e2e --no-cpu"
python bench.py $SMALL > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py $SMALL > gpurun_out/ncu_launches.log 2>&1
python bench.py $SMALL > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'label_pass_fast|distance_coop_kernel|trace_kernel' -s 4 -c 4 \
    -o gpurun_out/prof_r1 python bench.py $SMALL > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
