#!/bin/bash
# Run on the B200 box (through gpurun): plain bench lines of every configuration, the ncu launch list of the default
# workload at reduced size and one ncu --set full capture of each hot kernel.  Outputs land in gpurun_out/ (copy what
# should be judged into profiles/).  Usage: bash scripts/gpu_profile.sh r2
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
SMALL="--items 2048 --steps 3 --warmup 3 --no-e2e --no-cpu --no-secondary"
python bench.py > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err
for c in cfg1 cfg2 cfg3 cfg5; do python bench.py --config $c --steps 5 > gpurun_out/bench_${c}_$TAG.json 2> gpurun_out/bench_${c}_$TAG.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
python scripts/profile_kernels.py > gpurun_out/aux_kernels_$TAG.json 2> gpurun_out/aux_kernels_$TAG.err
python bench.py $SMALL > gpurun_out/bench_small_$TAG.json 2> gpurun_out/bench_small_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $SMALL > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'label_pass_fast|layered_distance_kernel|derive_kernel|totals_kernel' -s 15 -c 5 \
    -o gpurun_out/prof_$TAG python bench.py $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'argmax_planes|auc_kernel|rasterise_kernel|near_|layered_distance_kernel|trace_|distance_column' -c 24 \
    -o gpurun_out/prof_aux_$TAG python scripts/profile_kernels.py --quick > gpurun_out/ncu_aux_$TAG.log 2>&1
# the K <= 16 label pass and other geometries
python scripts/generic_lp_probe.py 2048 > gpurun_out/label_pass_shapes_$TAG.txt 2>&1
for s in "2048 496 1024 10" "2048 496 768 8" "2048 1024 512 8" "2048 496 500 8" "16384 496 512 8 2e-5" "16384 496 512 8 2e-3"; do
    python scripts/shape_probe.py $s >> gpurun_out/suite_shapes_$TAG.txt 2>&1
done
GLP_ONLY=0 ncu --set full --clock-control none --import-source on -k regex:label_pass_wide -s 3 -c 1 \
    -o gpurun_out/prof_wide_$TAG python scripts/generic_lp_probe.py 512 > gpurun_out/ncu_wide_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_full_$TAG.log gpurun_out/ncu_aux_$TAG.log
