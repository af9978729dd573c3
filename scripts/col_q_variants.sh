#!/bin/bash
# Tuning: queries per lane / group size / launch bounds of the column search kernel.  Args: "Q G MINB" triples.
for v in "$@"; do
  set -- $v
  OCTM_NVCC_EXTRA="-DOCTM_COL_Q=$1 -DOCTM_COL_GROUP=$2 -DOCTM_COL_MINB=$3" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python -m pytest tests/test_gpu_distance_modes.py tests/test_gpu_contours.py -m gpu -x -q 2>&1 | tail -1
  python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('Q G MINB = $v', d['value'], d['kernel_ms_per_step']['contour_distance'])"
done
