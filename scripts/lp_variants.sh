#!/bin/bash
# Tuning: label-pass launch bounds (registers) x ring geometry.
for m in 2 3; do
  OCTM_NVCC_EXTRA="-DOCTM_LP_MINB=$m" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  for g in "8 2" "8 3" "12 2" "16 2"; do set -- $g
    OCTM_LP_ROWS=$1 OCTM_LP_STAGES=$2 python bench.py --items 8192 --steps 4 --warmup 2 --no-e2e --no-cpu --no-contours 2>/dev/null |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minb $m rows $1 stages $2', 'label_pass_ms', round(d['kernel_ms_per_step']['label_pass'],4), 'frac', round(d['roofline']['frac'],4))"
    OCTM_LP_ROWS=$1 OCTM_LP_STAGES=$2 python bench.py --items 8192 --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   with seeds:', 'label_pass_ms', round(d['kernel_ms_per_step']['label_pass'],4), 'frac', round(d['roofline']['frac'],4))"
  done
done
