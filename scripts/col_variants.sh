#!/bin/bash
# Tuning: rebuild the library with different launch bounds of the column search kernel and time the suite.
for m in ${@:-5 4 6}; do
  OCTM_NVCC_EXTRA="-DOCTM_COL_MINB=$m" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('col minb $m', d['value'], d['kernel_ms_per_step'])"
done
