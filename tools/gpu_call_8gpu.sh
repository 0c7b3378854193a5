#!/bin/bash
# Round 2, the 8-GPU call: every N of the scaling table on one box, the sharded BASELINE configs at their GPU counts, and the
# reservoir-halo exchange on real GPUs (moving camera).  Device-timed max over ranks; the assembled frame is checked against the
# unsharded render inside every N > 1 run (frame_check).     gpurun --gpus 8 --timeout 1500 -- bash tools/gpu_call_8gpu.sh
mkdir -p gpurun_out
O=gpurun_out
T=r2c8
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $O/${T}_gpus.txt | wc -l
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  %-26s (no bench line: %s)" % (sys.argv[2], e)); sys.exit(0)
fc = d.get("frame_check", {})
print("  %-26s N=%d ms/frame %.3f e2e %.3f agree %.3f | frame min/med/max %.3f/%.3f/%.3f | Mrays/s %.0f | identical=%s differing=%s | %d tiles %s" % (
    sys.argv[2], d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["value_vs_e2e"]["relative_difference"], d["frame_ms"]["min"],
    d["frame_ms"]["median"], d["frame_ms"]["max"], d["value"], fc.get("identical"), fc.get("differing_pixels"), len(d["config"]["tiles"]), d["config"]["tiles"][:2]))
PY
}
run() {  # name nproc extra-args...
  local name=$1 n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --no-cpu-baseline "$@" 2> $O/${T}_$name.err | grep "^{" > $O/${T}_$name.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --no-cpu-baseline "$@" \
        2> $O/${T}_$name.err | grep "^{" > $O/${T}_$name.json
  fi
  short $O/${T}_$name.json $name
  grep -i "error\|Traceback" $O/${T}_$name.err | head -3
}
echo "== multi-GPU device tests (peer frame assembly, two-process halo exchange)"
timeout 600 python -m pytest tests/test_gpu_frame_assembly.py tests/test_gpu_zz_halo.py -m gpu -q 2>&1 | tail -3 | tee $O/${T}_pytest.txt
echo "== cornell 1080p (BASELINE configs[1]), strong scaling; N = 2 first: if the multi-GPU path is broken, stop before 8 GPUs idle"
run cornell1080p_n2 2 --steps 20 --warmup 5
[ -s $O/${T}_cornell1080p_n2.json ] || { echo "N=2 produced no bench line: aborting"; tail -20 $O/${T}_cornell1080p_n2.err; exit 1; }
for n in 1 4 8; do run cornell1080p_n$n $n --steps 20 --warmup 5; done
echo "== city 4K on 4 GPUs (configs[3]), city 8K on 8 (configs[4]), scene.rs 1080p on 8"
run city4k_n4 4 --config city_4k --steps 6 --warmup 3
run city8k_n8 8 --config city_8k --steps 4 --warmup 3
run scene1080p_n8 8 --config scene_1080p --steps 8 --warmup 4
echo "== reservoir-halo exchange on 8 physical GPUs: moving camera without / with --halo-margin 8 (bit-equality is frame_check)"
run cornell_moving_nohalo_n8 8 --steps 20 --warmup 5 --moving-camera
run cornell_moving_halo8_n8 8 --steps 20 --warmup 5 --moving-camera --halo-margin 8
ls $O | grep -c ${T}
