#!/bin/bash
set -e
for v in 1 8 10 12 16; do
  HK_NVCC_EXTRA="-DHK_MINB_DENOISE=$v" python bevy_hikari_b200/build.py > /dev/null
  python bench.py --steps 12 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_ms']
print('minb_denoise=$v', 'ms/frame %.3f' % d['ms_per_step'], ' '.join('%s=%.3f' % (n, k[n]) for n in k))
"
done
