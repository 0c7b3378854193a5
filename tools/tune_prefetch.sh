#!/bin/bash
set -e
for v in "-DHK_NO_PREFETCH" "-DHK_PREFETCH_ON"; do
  HK_NVCC_EXTRA="$v" python bevy_hikari_b200/build.py > /dev/null
  for cfg in cornell_1080p city_4k; do
  python bench.py --config $cfg --steps 12 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_ms']
print('$v $cfg', 'ms/frame %.3f' % d['ms_per_step'], ' '.join('%s=%.3f' % (n, k[n]) for n in k if n in ('gbuffer','direct','emissive','indirect')))
"
  done
done
