"""Parity at the BENCHMARK geometry (the largest frame compared on a device was 176x144): BASELINE configs[1] —
cornell 1920x1080, 2 bounces, temporal + emissive + indirect spatial ReSTIR, denoise — for 3 frames,
  * the whole frame on one context against the CPU oracle, every plane, bit for bit, and
  * the 8-strip partition bench.py uses at 8 GPUs (cost-balanced cuts from the coverage probe, 36-px ghosts) against the same
    oracle frames: every strip's tone-mapped image, radiance planes and indirect reservoirs.
The oracle needs ~2 s per 1080p frame on 16 host threads.  With HK_EMULATE_KERNELS=1 (no GPU) the geometry is scaled down 4x."""
import numpy as np
import pytest

import bench
from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import EMULATED, Bench
from tests.test_gpu_parity import ALL_PLANES, DENOISED, compare_all, mismatch

pytestmark = pytest.mark.gpu

W, H = (480, 270) if EMULATED else (1920, 1080)
FRAMES = 3


@pytest.fixture(scope="module")
def oracle_frames():
    """the oracle's planes for frames 1..3 of the benchmark configuration (computed once for both tests)"""
    b = Bench("cornell", W, H, config="cornell_1080p")
    orc = b.oracle()
    keep = (L.OUT_TONE_MAPPED, L.OUT_RENDER_DIRECT, L.OUT_RENDER_EMISSIVE, L.OUT_RENDER_INDIRECT, L.OUT_RESERVOIR_0 + 8,
            L.OUT_RESERVOIR_0 + 9, L.OUT_GBUFFER_INSTANCE_MATERIAL)
    frames = []
    full = b.device()
    full.set_keep_intermediates(True)
    for f in range(1, FRAMES + 1):
        inp = b.inputs(f)
        orc.render_frame(inp)
        full.render_frame(inp)
        compare_all(full, orc, ALL_PLANES + DENOISED, f)          # test 1: the unsharded frame, every plane
        frames.append({k: orc.readback(k).copy() for k in keep})
    full.close()
    return b, frames


def test_full_frame_at_benchmark_resolution_bit_exact(oracle_frames):
    b, frames = oracle_frames
    ids = frames[-1][L.OUT_GBUFFER_INSTANCE_MATERIAL]
    assert ids.shape[:2] == (H, W) and (ids[..., 0] > 0).mean() > 0.2      # the box covers a good part of the 16:9 frame


def test_the_eight_strips_of_the_bench_at_benchmark_resolution(oracle_frames):
    b, frames = oracle_frames
    # bench.py's own plan: quarter-resolution coverage probe -> cost-balanced strips
    probe = plugin.HikariPlugin(max(W // 4, 1), max(H // 4, 1))
    probe.upload_scene(b.world)
    pv, ppv, pl = b.scene.view_inputs(max(W // 4, 1), max(H // 4, 1))
    probe.prepass(plugin.make_frame_inputs(b.settings, 1, pv, ppv, pl))
    coverage = probe.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)[..., 0] > 0
    probe.close()
    tiles = bench.plan_tiles(W, H, 8, coverage)
    assert len(tiles) == 8 and sorted(set(t[0] for t in tiles) | {t[1] for t in tiles})[0] == 0
    devs = [b.device(r0, r1, x0, x1) for (x0, x1, r0, r1) in tiles]
    for f in range(1, FRAMES + 1):
        inp = b.inputs(f)
        for d in devs:
            d.render_frame(inp)
        for k, whole in frames[f - 1].items():
            for d, (x0, x1, r0, r1) in zip(devs, tiles):
                assert mismatch(d.readback(k), whole[r0:r1, x0:x1]) == 0, (f, k, (x0, x1, r0, r1))
