#!/bin/bash
# First gpurun call of the next round (about 8 GPU-minutes).  Everything below was written after round 1's GPU budget was
# spent and is validated only on the emulated kernels (tests/emu/); this confirms it on the device and times the two prepared
# tuning variants.  Before calling (on the CPU box, nvcc cross-compiles):
#     python tools/build_variants.py denoise='-DHK_DENOISE_BRANCHFREE=1' spatial='-DHK_SPATIAL_EAGER_LOAD=1' \
#            fastdiv='-DHK_SPATIAL_FAST_DIV=1' all='-DHK_DENOISE_BRANCHFREE=1 -DHK_SPATIAL_EAGER_LOAD=1 -DHK_SPATIAL_FAST_DIV=1' \
#            surface_loop='-DHK_SURFACE_LOOP=1'   (textured scenes only: time it with --config city_4k / scene_1080p) \
#            generic_direct='-DHK_NOVAL_VARIANT=0' \
#            generic_texture_path='-DHK_NO_TEXTURE_VARIANT=0'      (the round-1 kernels: the default now picks NO_TEXTURE variants for cornell)
#     gpurun --timeout 1200 -- tools/next_round_first_call.sh
mkdir -p gpurun_out
echo "== device suite, newest tests first"
python -m pytest tests/test_gpu_zz_*.py -m gpu -q 2>&1 | tail -8
python -m pytest tests -m gpu -q $(for f in tests/test_gpu_zz_*.py; do echo --deselect $f; done) 2>&1 | tail -4
echo "== tuning variants (per-kernel ms; make a variant the default if its kernels drop)"
python bench.py --steps 16 --warmup 4 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernel_ms']
print('default       ', 'ms/frame %.3f' % d['ms_per_step'], ' '.join('%s=%.3f' % (n[:8], k[n]) for n in k))"
[ -d bevy_hikari_b200/variants ] && tools/sweep_variants.sh
echo "== BASELINE configs[2] (examples/scene.rs), never timed on a device so far"
python bench.py --config scene_1080p --steps 16 --warmup 4 2>/dev/null | grep "^{" | tee gpurun_out/bench_r2_scene1080p.json | cut -c1-400
echo "== bench line"
python bench.py --steps 32 --warmup 8 2>/dev/null | grep "^{" | tee gpurun_out/bench_r2_first.json | cut -c1-400
