#!/bin/bash
# Runs ON THE GPU BOX: times every prebuilt tuning variant (tools/build_variants.py) with a short bench run.
for so in bevy_hikari_b200/variants/*.so; do
  python bench.py --lib $PWD/$so --steps 16 --warmup 4 --no-cpu-baseline 2>/dev/null | grep "^{" | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_ms']
print('%-14s' % '$(basename $so .so)', 'ms/frame %.3f' % d['ms_per_step'], ' '.join('%s=%.3f' % (n[:8], k[n]) for n in k))
"
done
