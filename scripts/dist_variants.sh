#!/bin/bash
# Tuning: rebuild the library with different launch bounds of the search kernel and time the suite.
for m in 4 5 6; do
  OCTM_NVCC_EXTRA="-DOCTM_SEARCH_MINB=$m" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minb $m', d['value'], d['kernel_ms_per_step'])"
done
