#!/bin/bash
# Tuning: query-group size of the column search kernel (tests + timing per variant).
for g in ${@:-16 8 32}; do
  OCTM_NVCC_EXTRA="-DOCTM_COL_GROUP=$g" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python -m pytest tests/test_gpu_distance_modes.py tests/test_gpu_contours.py -m gpu -x -q 2>&1 | tail -1
  python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('col group $g', d['value'], d['kernel_ms_per_step'])"
done
