#!/bin/bash
# Round 2, second gpurun call (1 GPU): the pooled indirect kernel (kc_indirect) on the device — parity, then A/B of its tuning
# knobs (traversal warps per CTA, CTAs per SM) and of the per-pixel kernel it replaces, on three configs; ncu of the new kernel.
mkdir -p gpurun_out
O=gpurun_out
T=r2c2
echo "== device parity of the pooled kernel"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_examples.py tests/test_gpu_zz_full_resolution.py tests/test_gpu_dynamic.py tests/test_gpu_context_state.py -m gpu -q -x 2>&1 | tail -4 | tee $O/${T}_pytest.txt
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  %-18s (no bench line: %s)" % (sys.argv[2], e)); sys.exit(0)
k = d.get("kernel_ms", {})
print("  %-18s %-14s ms/frame %.3f e2e %.3f | indirect %.3f direct %.3f emissive %.3f ind_spa %.3f gbuffer %.3f" % (
    sys.argv[2], d["config"]["workload"].split(":")[0], d["ms_per_step"], d["e2e"]["ms_per_step"], k.get("indirect", 0), k.get("direct", 0),
    k.get("emissive", 0), k.get("indirect_spatial", 0), k.get("gbuffer", 0)))
PY
}
run() {  # name lib-or-empty config steps warmup
  local libarg=""; [ -n "$2" ] && libarg="--lib $PWD/$2"
  timeout 600 python bench.py $libarg --config $3 --steps $4 --warmup $5 --no-cpu-baseline 2> $O/${T}_$1_$3.err | grep "^{" > $O/${T}_$1_$3.json
  short $O/${T}_$1_$3.json $1
}
echo "== A/B"
for cfg in "cornell_1080p 16 4" "scene_1080p 8 4" "city_4k 6 3"; do
  set -- $cfg
  run pooled "" $1 $2 $3
  for v in perpixel pool_ntw2 pool_ntw8 pool_minb2 pool_minb4 pool_ntw2_minb4; do
    [ -f bevy_hikari_b200/variants/$v.so ] && run $v bevy_hikari_b200/variants/$v.so $1 $2 $3
  done
done
echo "== ncu --set full of kc_indirect (cornell 1080p, city 4K, scene 1080p): one warm launch each"
for cfg in cornell_1080p city_4k scene_1080p; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:kc_indirect' --launch-skip 3 --launch-count 1 \
      -o $O/${T}_full_kc_indirect_$cfg -f python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_full_$cfg.log 2>&1
done
ls -la $O | grep $T | tail -40
