#!/bin/bash
# Round 2, fifth gpurun call (1 GPU): kc_spatial and kc_denoise (TMA-tiled spatial reuse / a-trous levels) on the device — parity + tolerance gate, A/B against the
# gather form (same library, HK_TUNE_TILED_SPATIAL=0) and of the CTAs-per-SM knob, ncu of the new kernel.
mkdir -p gpurun_out
O=gpurun_out
T=r2c5
echo "== device tests"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tolerance.py tests/test_gpu_zz_full_resolution.py tests/test_gpu_upscale.py tests/test_gpu_zz_halo.py tests/test_gpu_context_state.py tests/test_gpu_variants.py -m gpu -q 2>&1 | tail -12 | tee $O/${T}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  %-12s (no bench line: %s)" % (sys.argv[2], e)); sys.exit(0)
k = d.get("kernel_ms", {})
dn = sum(v for n, v in k.items() if n.startswith("denoise") or n == "demodulation")
print("  %-10s %-14s ms/frame %.3f e2e %.3f | gbuf %.3f direct %.3f emis %.3f emis_spa %.3f indirect %.3f ind_spa %.3f denoise %.3f" % (
    sys.argv[2], d["config"]["workload"].split(":")[0], d["ms_per_step"], d["e2e"]["ms_per_step"], k.get("gbuffer", 0), k.get("direct", 0),
    k.get("emissive", 0), k.get("emissive_spatial", 0), k.get("indirect", 0), k.get("indirect_spatial", 0), dn))
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
  run tiled "" $1 $2 $3
  HK_TUNE_TILED_SPATIAL=0 run gather_sp "" $1 $2 $3
  HK_TUNE_TILED_DENOISE=0 run gather_dn "" $1 $2 $3
  for v in sp_ind2 sp_emi4 sp_emi6 mi6 mi5 md6; do
    [ -f bevy_hikari_b200/variants/$v.so ] && run $v bevy_hikari_b200/variants/$v.so $1 $2 $3
  done
done
echo "== ncu --set full of kc_spatial (cornell 1080p: both pipelines; city 4K: indirect)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:kc_spatial|kc_denoise' --launch-skip 18 --launch-count 6 \
    -o $O/${T}_full_kc_spatial_cornell -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_full_cornell.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:kc_spatial' --launch-skip 3 --launch-count 1 \
    -o $O/${T}_full_kc_spatial_city -f python bench.py --config city_4k --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_full_city.log 2>&1
ls $O | grep -c ${T}
