#!/bin/bash
# Second gpurun call of the next round (one GPU, about 6 GPU-minutes): the profile evidence of the CURRENT default build, in the form
# profiles/ wants it.  Nothing printed under ncu is a bench number.
#   gpurun --timeout 900 -- tools/next_round_profile.sh
# Afterwards, on the CPU box:
#   ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > /tmp/raw.csv          (per-kernel metrics -> profiles/r2_ncu_full_summary.txt)
#   python tools/ncu_lines.py gpurun_out/r2_full.ncu-rep k_indirect 40         (hot lines -> profiles/r2_ncu_k_indirect_hot_lines.txt)
set -x
mkdir -p gpurun_out
# 1. launch list of one bench command (per-launch times are cold-cache and serialised: the SHARE of each kernel is what must
#    agree with the live per-kernel CUDA-event times of the bench line)
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
# 2. full capture of the kernels that carry the frame, warm temporal state (frame 4), one launch each
ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_indirect|k_spatial|k_denoise|k_direct|k_gbuffer|k_demodulation' \
    --launch-skip 42 --launch-count 14 -o gpurun_out/r2_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_full_bench.log 2>&1
# 3. the same for the textured workload (generic instantiations; BASELINE configs[2] scene at a quarter of the pixels to stay short)
ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_indirect|k_spatial' \
    --launch-skip 42 --launch-count 4 -o gpurun_out/r2_full_scene -f python bench.py --config scene_1080p --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2_full_scene_bench.log 2>&1
# 4. the bench lines themselves (NOT under ncu)
python bench.py --steps 32 --warmup 8 2>/dev/null | grep "^{" > gpurun_out/r2_bench_cornell1080p_1gpu.json
python bench.py --config scene_1080p --steps 16 --warmup 8 2>/dev/null | grep "^{" > gpurun_out/r2_bench_scene1080p_1gpu.json
python bench.py --config city_4k --steps 8 --warmup 4 2>/dev/null | grep "^{" > gpurun_out/r2_bench_city4k_1gpu.json
ls -la gpurun_out | tail -12
