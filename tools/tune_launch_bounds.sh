#!/bin/bash
# Runs ON THE GPU BOX (nvcc is in the image): rebuild the library with different __launch_bounds__ min-blocks for the
# three register-heavy kernel families and print the per-kernel times of a short bench run.
set -e
for v in "8 8 8" "10 8 8" "10 10 10" "12 10 10" "12 12 12" "16 12 12"; do
  set -- $v
  HK_NVCC_EXTRA="-DHK_MINB_INDIRECT=$1 -DHK_MINB_DIRECT=$2 -DHK_MINB_SPATIAL=$3" python bevy_hikari_b200/build.py > /dev/null
  python bench.py --steps 12 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_ms']
print('minb(indirect,direct,spatial)=$v', 'ms/frame %.3f' % d['ms_per_step'], ' '.join('%s=%.3f' % (n, k[n]) for n in ('gbuffer','direct','emissive','emissive_spatial','indirect','indirect_spatial')))
"
done
