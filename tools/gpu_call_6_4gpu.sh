#!/bin/bash
# Round 2, sixth gpurun call (4 GPUs): the multi-GPU path on physical GPUs after the round-2 rework of bench.py — peer frame assembly and
# the two-process halo exchange (device tests), cornell 1080p at N = 2 / 4, city 4K at its BASELINE GPU count (configs[3]: 4 GPUs),
# scene.rs at 4, and the reservoir-halo exchange under a moving camera without / with --halo-margin (frame_check is the bit-equality).
#     gpurun --gpus 4 --timeout 900 -- bash tools/gpu_call_6_4gpu.sh
mkdir -p gpurun_out
O=gpurun_out
T=r2c6
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
    d["frame_ms"]["median"], d["frame_ms"]["max"], d["value"], fc.get("identical"), fc.get("differing_pixels"), d["config"]["tiles"]))
PY
}
run() {  # name nproc extra-args...
  local name=$1 n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --no-cpu-baseline "$@" \
      2> $O/${T}_$name.err | grep "^{" > $O/${T}_$name.json
  short $O/${T}_$name.json $name
  grep -i "error\|Traceback" $O/${T}_$name.err | head -3 | cut -c1-300
}
echo "== multi-GPU device tests (peer frame assembly, two-process halo exchange)"
timeout 400 python -m pytest tests/test_gpu_frame_assembly.py tests/test_gpu_zz_halo.py -m gpu -q 2>&1 | tail -4 | tee $O/${T}_pytest.txt
echo "== cornell 1080p (configs[1]) strong scaling"
run cornell1080p_n2 2 --steps 20 --warmup 5
[ -s $O/${T}_cornell1080p_n2.json ] || { echo "N=2 produced no bench line"; tail -30 $O/${T}_cornell1080p_n2.err | cut -c1-400; }
run cornell1080p_n4 4 --steps 20 --warmup 5
echo "== city 4K on 4 GPUs (configs[3]), scene.rs 1080p on 4"
run city4k_n4 4 --config city_4k --steps 6 --warmup 3
run scene1080p_n4 4 --config scene_1080p --steps 8 --warmup 4
echo "== reservoir-halo exchange on physical GPUs: moving camera without / with --halo-margin 8"
run cornell_moving_nohalo_n4 4 --steps 20 --warmup 5 --moving-camera
run cornell_moving_halo8_n4 4 --steps 20 --warmup 5 --moving-camera --halo-margin 8
ls $O | grep -c ${T}
