#!/bin/bash
# Round 2, twelfth gpurun call (8 GPUs): BASELINE configs[4] (city 7680x4320, 4 bounces, full ReSTIR + denoise) at its GPU count, cornell 1080p
# at 8, the reservoir-halo exchange under a moving camera on 8 physical GPUs, and the same-process peer frame target with its traceback.
#     gpurun --gpus 8 --timeout 600 -- bash tools/gpu_call_12_8gpu.sh
mkdir -p gpurun_out
O=gpurun_out
T=r2c12
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $O/${T}_gpus.txt | wc -l
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  %-24s (no bench line: %s)" % (sys.argv[2], e)); sys.exit(0)
fc = d.get("frame_check") or {}
print("  %-24s N=%d ms/frame %.3f e2e %.3f agree %.3f | frame min/med/max %.3f/%.3f/%.3f | Mrays/s %.0f | identical=%s differing=%s | %s" % (
    sys.argv[2], d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["value_vs_e2e"]["relative_difference"], d["frame_ms"]["min"],
    d["frame_ms"]["median"], d["frame_ms"]["max"], d["value"], fc.get("identical"), fc.get("differing_pixels"), d["config"]["tiles"][:3]))
PY
}
run() {  # name nproc extra-args...
  local name=$1 n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --no-cpu-baseline "$@" \
      2> $O/${T}_$name.err | grep "^{" > $O/${T}_$name.json
  short $O/${T}_$name.json $name
  grep -i "error\|Traceback" $O/${T}_$name.err | head -3 | cut -c1-300
}
echo "== same-process peer frame target (2 of the 8 GPUs)"
timeout 200 python -m pytest tests/test_gpu_frame_assembly.py -m gpu -q --tb=short 2>&1 | tail -30 | cut -c1-220 | tee $O/${T}_pytest.txt
echo "== city 8K on 8 GPUs (configs[4])"
run city8k_n8 8 --config city_8k --steps 4 --warmup 3
echo "== cornell 1080p on 8 GPUs: static camera, moving camera with the halo exchange"
run cornell1080p_n8 8 --steps 20 --warmup 5
run cornell_moving_halo8_n8 8 --steps 20 --warmup 5 --moving-camera --halo-margin 8
ls $O | grep -c ${T}
