#!/bin/bash
# Round 2, call 15 (1 GPU, 10.8 GPU-minutes left): after the staging-alignment fix of hk_scene_update_transforms (adf6464).
# Self-limiting: every step is skipped once the call has used its share, so that the call ends well inside the budget.
#   1. the device tests that have never passed on hardware (device-side scene rebuild incl. the soups; CUDA path against the WGSL fixtures)
#   2. the default bench line of the final tree (roofline.traffic from profiles/r2_dram_traffic.json)
#   3. the other four BASELINE configs on one GPU in the final build (short runs, no CPU baseline)
#   4. the whole device suite in the time that is left (verbose log, so that a cut-off run still shows how far it got)
mkdir -p gpurun_out
O=gpurun_out
T=r2c15
START=$(date +%s)
LIMIT=${HK_CALL_LIMIT:-430}
left() { echo $(( LIMIT - ( $(date +%s) - START ) )); }
echo "== 1. new device tests (one process per file: a sticky fault in one must not hide the other)"
timeout 100 python -m pytest tests/test_gpu_wgsl_golden.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee $O/${T}_new_tests_wgsl.txt
timeout 100 python -m pytest tests/test_gpu_scene_update.py -m gpu -q -x -p no:cacheprovider > $O/${T}_new_tests_scene.txt 2>&1
tail -8 $O/${T}_new_tests_scene.txt
if ! grep -q " passed" $O/${T}_new_tests_scene.txt || grep -q "failed" $O/${T}_new_tests_scene.txt; then
    echo "== 1b. scene-update tests under compute-sanitizer (memcheck)"
    timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_scene_update.py -m gpu -q -x -p no:cacheprovider -k "soups" > $O/${T}_memcheck.txt 2>&1
    grep -A25 "=========" $O/${T}_memcheck.txt | head -80
fi
echo "== 2. default bench line ($(left) s left)"
timeout 120 python bench.py > $O/${T}_default.json 2> $O/${T}_default.err; tail -c 400 $O/${T}_default.err
echo "== 3. the other configs ($(left) s left)"
for c in cornell_256 scene_1080p city_4k city_8k; do
    if [ $(left) -gt 200 ]; then
        timeout 70 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline 2> $O/${T}_$c.err | grep "^{" > $O/${T}_$c.json
    fi
done
python - <<'PY'
import json
for n in ("default", "cornell_256", "scene_1080p", "city_4k", "city_8k"):
    try:
        d = json.loads(open(f"gpurun_out/r2c15_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/frame", d.get("ms_per_step"), "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "roofline", (d.get("roofline") or {}).get("frac"),
              "traffic", (d.get("roofline") or {}).get("traffic"))
        print("   kernel_ms", d.get("kernel_ms"))
    except Exception as e:
        print(n, "no line:", e)
PY
R=$(left)
echo "== 4. whole device suite ($R s left)"
if [ $R -gt 40 ]; then
    timeout -s INT $(( R - 15 )) python -m pytest tests -m gpu -v -p no:cacheprovider --durations=25 > $O/${T}_pytest_full.txt 2>&1
    grep -c PASSED $O/${T}_pytest_full.txt; grep -E "FAILED|ERROR" $O/${T}_pytest_full.txt | head -20; tail -32 $O/${T}_pytest_full.txt
fi
echo "== done after $(( $(date +%s) - START )) s"
