#!/bin/bash
# Tuning: launch bounds of the layered contour kernel.
for m in ${@:-10 12 8}; do
  OCTM_NVCC_EXTRA="-DOCTM_LAYERED_MINB=$m" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('layered minb $m', d['value'], d['kernel_ms_per_step'])"
done
