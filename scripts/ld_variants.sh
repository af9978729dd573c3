#!/bin/bash
# Tuning run on the B200 box: rebuild liboctm.so with different launch bounds of layered_distance_kernel and print the
# per-kernel times of a 16,384-item cfg4 step (octm_profile_*).  Usage: bash scripts/ld_variants.sh "8 10 12"
set -u
for m in ${1:-8 10 12}; do
  OCTM_NVCC_EXTRA="-DOCTM_LD_MINB=$m" python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
  python bench.py --items 16384 --steps 5 --warmup 2 --no-e2e --no-cpu --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('MINB=$m', 'step %.3f ms' % d['ms_per_step'], ' '.join('%s=%.3f' % (k['kernel'], k['ms_per_step']) for k in d['roofline']['kernels'][:5]))"
done
python -m retinal_oct_image_segmentation_via_deep_learning_b200.csrc.build --force > /dev/null
