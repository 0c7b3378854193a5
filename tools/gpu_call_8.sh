#!/bin/bash
# Round 2, eighth gpurun call (1 GPU): the image-exact traversal mode (4-wide trees, ordered walk; csrc/hk_wide.cuh) on the device —
# its gate tests, the whole device suite on the new default, smoke, A/B against the reference's fixed-order walk on the same library
# (HK_TUNE_WIDE_TRAVERSAL=0) for BASELINE configs 2-5, tuning variants of the walk, and an ncu capture of the walking kernels.
mkdir -p gpurun_out
O=gpurun_out
T=r2c8
echo "== device tests"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $O/${T}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  %-12s (no bench line: %s)" % (sys.argv[2], e)); sys.exit(0)
k = d.get("kernel_ms", {})
dn = sum(v for n, v in k.items() if n.startswith("denoise") or n == "demodulation")
print("  %-10s %-14s ms/frame %.3f e2e %.3f | gbuf %.3f direct %.3f emis %.3f emis_spa %.3f indirect %.3f ind_spa %.3f denoise %.3f | Mrays/s %.0f" % (
    sys.argv[2], d["config"]["workload"].split(":")[0], d["ms_per_step"], d["e2e"]["ms_per_step"], k.get("gbuffer", 0), k.get("direct", 0),
    k.get("emissive", 0), k.get("emissive_spatial", 0), k.get("indirect", 0), k.get("indirect_spatial", 0), dn, d["value"]))
PY
}
run() {  # name lib-or-empty config steps warmup
  local libarg=""; [ -n "$2" ] && libarg="--lib $PWD/$2"
  timeout 600 python bench.py $libarg --config $3 --steps $4 --warmup $5 --no-cpu-baseline 2> $O/${T}_$1_$3.err | grep "^{" > $O/${T}_$1_$3.json
  short $O/${T}_$1_$3.json $1
}
echo "== A/B: wide (default) against the fixed-order walk, same library"
for cfg in "cornell_1080p 16 4" "scene_1080p 8 4" "city_4k 6 3" "city_8k 3 3"; do
  set -- $cfg
  run wide "" $1 $2 $3
  HK_TUNE_WIDE_TRAVERSAL=0 run fixed "" $1 $2 $3
done
echo "== variants of the walk"
for cfg in "cornell_1080p 16 4" "scene_1080p 8 4" "city_4k 6 3"; do
  set -- $cfg
  for v in w_noinl w_mb8 w_mb5 w_mb4 w_smem; do
    [ -f bevy_hikari_b200/variants/$v.so ] && run $v bevy_hikari_b200/variants/$v.so $1 $2 $3
  done
done
echo "== default bench line (cpu baseline included) and the reference arm"
timeout 600 python bench.py > $O/${T}_default.json 2> $O/${T}_default.err; short $O/${T}_default.json default
echo "== ncu: launch list + full capture of the walking kernels, city 4K and cornell"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${T}_launches_city4k.csv python bench.py --config city_4k --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_launches_city.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_indirect|k_gbuffer|k_direct' --launch-skip 9 --launch-count 4 \
    -o $O/${T}_full_city -f python bench.py --config city_4k --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_full_city.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_indirect' --launch-skip 3 --launch-count 1 \
    -o $O/${T}_full_cornell -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_full_cornell.log 2>&1
ls $O | grep -c ${T}
