#!/bin/bash
# Round 2, first gpurun call (1 GPU, ~20 min): the state of the CURRENT kernels under the corrected measurement —
#   device suite (incl. the new 1080p parity tests), the five BASELINE configs at N = 1, the reference arm, the tuning variants left
#   from round 1, launch lists of three configs and a full ncu capture of the kernels that carry the frame.
# Nothing printed under ncu is a bench number.     gpurun --timeout 1800 -- bash tools/gpu_call_1.sh
mkdir -p gpurun_out
O=gpurun_out
T=r2c1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/${T}_gpu.txt 2>&1
echo "== device suite"
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $O/${T}_pytest.txt

short() { python - "$1" <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  (no bench line:", e, ")"); sys.exit(0)
k = d.get("kernel_ms", {})
r = d.get("roofline", {})
print("  %-14s ms/frame %.3f  e2e %.3f  Mrays/s %.0f  dom %s %.3f ms frac %.3f  frame-frac %.3f  agree %s" % (
    d["config"]["workload"].split(":")[0], d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], r.get("kernel"), r.get("kernel_ms", 0),
    r.get("frac", 0), r.get("frame", {}).get("frac", 0), d.get("value_vs_e2e", {}).get("relative_difference")))
print("   ", " ".join("%s=%.3f" % (n[:12], v) for n, v in k.items()))
EOF
}

echo "== bench, five configs, N = 1"
run_cfg() {   # config steps warmup
  timeout 900 python bench.py --config $1 --steps $2 --warmup $3 --no-cpu-baseline --print-frame-hash 2> $O/${T}_bench_$1.err | grep "^{" > $O/${T}_bench_$1_1gpu.json
  short $O/${T}_bench_$1_1gpu.json
}
run_cfg cornell_1080p 32 8
run_cfg cornell_256 64 8
run_cfg scene_1080p 16 6
run_cfg city_4k 12 4
run_cfg city_8k 6 3
echo "== the driver's two lines (default flags)"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2> $O/${T}_ref.err | grep "^{" > $O/${T}_bench_reference_default.json
python -c "import json; d=json.loads(open('$O/${T}_bench_reference_default.json').read()); print('  reference: %.3f Mrays/s, %s ms/step, rendered %sx%s, cores %s' % (d['value'], d['ms_per_step'], d['config']['rendered_width'], d['config']['rendered_height'], d['cpu_baseline']['cores']))"
timeout 600 python bench.py --steps 20 --warmup 5 2> $O/${T}_default.err | grep "^{" > $O/${T}_bench_default.json
short $O/${T}_bench_default.json

echo "== tuning variants left from round 1 (cornell 1080p)"
for so in bevy_hikari_b200/variants/*.so; do
  [ -f "$so" ] || continue
  timeout 300 python bench.py --lib $PWD/$so --steps 16 --warmup 4 --no-cpu-baseline 2>/dev/null | grep "^{" > $O/${T}_variant_$(basename $so .so).json
  echo "  variant $(basename $so .so):"; short $O/${T}_variant_$(basename $so .so).json
done

echo "== ncu launch lists (per-launch times are cold-cache and serialised: shares only)"
for cfg in cornell_1080p scene_1080p city_4k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches_$cfg.csv \
      python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_launches_$cfg.log 2>&1
  echo "  $cfg: $(grep -c k_ $O/${T}_launches_$cfg.csv) kernel rows"
done
echo "== ncu --set full: one warm launch of each heavy kernel, cornell 1080p and city 4K"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_indirect|k_spatial|k_denoise|k_direct|k_gbuffer|k_demodulation' \
    --launch-skip 42 --launch-count 14 -o $O/${T}_full_cornell -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_full_cornell.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_indirect|k_spatial' \
    --launch-skip 6 --launch-count 2 -o $O/${T}_full_city -f python bench.py --config city_4k --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_full_city.log 2>&1
ls -la $O | tail -40
