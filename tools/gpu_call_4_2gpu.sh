#!/bin/bash
# Round 2, fourth gpurun call (2 GPUs): the multi-GPU path of bench.py after its rework, before the 8-GPU call pays for it —
# peer / IPC tests on two physical GPUs, N = 2 with the static camera (assembled frame must equal the unsharded one), with a moving
# camera without and with the reservoir-halo exchange (--halo-margin): the exchange must restore bit-equality, and its cost is the
# difference between the two lines.
mkdir -p gpurun_out
O=gpurun_out
T=r2c4
nvidia-smi --query-gpu=index,name --format=csv | tee $O/${T}_gpus.txt
echo "== multi-GPU device tests"
timeout 900 python -m pytest tests/test_gpu_frame_assembly.py tests/test_gpu_zz_halo.py -m gpu -q 2>&1 | tail -6 | tee $O/${T}_pytest.txt
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  %-22s (no bench line: %s)" % (sys.argv[2], e)); sys.exit(0)
fc = d.get("frame_check", {})
print("  %-22s N=%d ms/frame %.3f e2e %.3f agree %.3f | frame min/med/max %.3f/%.3f/%.3f | frame_check identical=%s differing=%s | tiles %s" % (
    sys.argv[2], d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["value_vs_e2e"]["relative_difference"], d["frame_ms"]["min"],
    d["frame_ms"]["median"], d["frame_ms"]["max"], fc.get("identical"), fc.get("differing_pixels"), d["config"]["tiles"]))
PY
}
run() {  # name nproc extra-args...
  local name=$1 n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --no-cpu-baseline "$@" \
      2> $O/${T}_$name.err | grep "^{" > $O/${T}_$name.json
  short $O/${T}_$name.json $name
  tail -3 $O/${T}_$name.err | cut -c1-300
}
run n1_static 1 --steps 20 --warmup 5
run n2_static 2 --steps 20 --warmup 5
run n2_moving_nohalo 2 --steps 20 --warmup 5 --moving-camera
run n2_moving_halo8 2 --steps 20 --warmup 5 --moving-camera --halo-margin 8
run n2_static_halo8 2 --steps 20 --warmup 5 --halo-margin 8
run n2_city4k 2 --config city_4k --steps 6 --warmup 3
