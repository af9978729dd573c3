#!/bin/bash
# Tuning: variants of the contour walk's label loads (extra nvcc flags per variant).
run() {
  OCTM_NVCC_EXTRA="$1" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['kernel_ms_per_step']['contour_trace'])"
}
run ""
run "-DOCTM_TRACE_NO_PREFETCH"
run "-DOCTM_TRACE_MINB=9"
run "-DOCTM_TRACE_MINB=10"
