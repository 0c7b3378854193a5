#!/bin/bash
# Round 2, last gpurun call (1 GPU, ~14 GPU-minutes left): what was written after the round's measurements, on the device —
#   1. the new device tests first: the device-side scene rebuild (hk_scene_update_transforms) and the CUDA path against the fixtures
#      computed from the reference's own WGSL;
#   2. the final default bench line (cornell 1080p, with cpu_baseline and the scene_update block: host path against device rebuild),
#      the city 4K line (the scene where the 4-wide TLAS is re-derived after a device rebuild) and the reference arm;
#   3. the ncu launch list of the default bench command + one full capture of the dominant kernel (roofline.traffic);
#   4. the whole device suite with durations (the oracle used to run 128 OpenMP threads under a 16-CPU quota: 683 s in call 8).
mkdir -p gpurun_out
O=gpurun_out
T=r2c14
echo "== new device tests"
timeout 240 python -m pytest tests/test_gpu_scene_update.py tests/test_gpu_wgsl_golden.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | tee $O/${T}_new_tests.txt
echo "== bench lines (not under ncu)"
timeout 300 python bench.py > $O/${T}_default.json 2> $O/${T}_default.err; tail -c 600 $O/${T}_default.err
timeout 200 python bench.py --config city_4k --steps 6 --warmup 3 --no-cpu-baseline 2> $O/${T}_city_4k.err | grep "^{" > $O/${T}_city_4k.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 3 2> $O/${T}_reference.err | grep "^{" > $O/${T}_reference.json
python - <<'PY'
import json
for n in ("default", "city_4k", "reference"):
    try:
        d = json.loads(open(f"gpurun_out/r2c14_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/frame", d.get("ms_per_step"), "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "roofline", (d.get("roofline") or {}).get("frac"))
        print("   kernel_ms", d.get("kernel_ms"))
        print("   scene_update", d.get("scene_update"))
    except Exception as e:
        print(n, "no line:", e)
PY
echo "== ncu: launch list of the default command, full capture of the dominant kernel"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_cornell_1080p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_indirect' --launch-skip 3 --launch-count 1 -o $O/${T}_full_k_indirect -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_full.log 2>&1
ncu -i $O/${T}_full_k_indirect.ncu-rep --page raw --csv 2>/dev/null > $O/${T}_full_k_indirect_raw.csv
echo "== smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== whole device suite, with durations"
timeout -s INT 400 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=25 2>&1 | tail -45 | tee $O/${T}_pytest.txt
ls $O | grep -c ${T}
