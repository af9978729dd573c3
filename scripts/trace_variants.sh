#!/bin/bash
# Tuning: registers (launch bounds) of the contour walk.
for m in ${@:-8 10 12 16}; do
  OCTM_NVCC_EXTRA="-DOCTM_TRACE_MINB=$m" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('minb $m', d['value'], d['kernel_ms_per_step']['contour_trace'])"
done
