"""GPU parity for the row after the hot path: rendering below the output resolution
(`Upscale::ratio` in (1, 2], light.rs:622-624 — jittered deferred look-ups, render-size reservoirs) and the temporal
upscalers that follow tone mapping (smaa_tu4x + smaa_tu4x_extrapolate, taa_jasmine; post_process.rs:1236-1277).
Same bar as the rest of the path: every plane bit-exact against the CPU oracle, tolerance 0."""
import numpy as np
import pytest

from bevy_hikari_b200 import _ffi
from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_gpu_parity import ALL_PLANES, DENOISED, compare_all, mismatch

pytestmark = pytest.mark.gpu


def scaled(size, ratio):
    return int(np.ceil(np.float32(1.0) / np.float32(ratio) * np.float32(size)))


def upscaled(size, ratio):
    """extent of upscale_output after SMAA TU4x: create_texture(format, ratio.recip() * 2.0) = ceil(size * scale) in f32
    (post_process.rs:663-667,711,717) — NOT always 2 * scaled(size, ratio): 50 rows at ratio 1.5 render 34 rows and upscale to 67"""
    return int(np.ceil(np.float32(size) * (np.float32(1.0) / np.float32(ratio) * np.float32(2.0))))


@pytest.mark.parametrize("ratio", [2.0, 1.5, 1.3])
def test_upscale_ratio_bit_exact(ratio):
    """Light + denoise + tone mapping at render size ceil(size / ratio), G-buffer at full size; moving camera so the
    reprojection through jittered_deferred_uv and the scatter writes are exercised."""
    b = Bench("cornell", 112, 80, config="cornell_1080p", upscale_ratio=ratio)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    for f in range(1, 9):
        inp = b.moving_inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)
    assert dev.readback(L.OUT_TONE_MAPPED).shape[:2] == (scaled(80, ratio), scaled(112, ratio))
    assert dev.readback(L.OUT_GBUFFER_POSITION).shape[:2] == (80, 112)
    assert dev.readback(L.OUT_RESERVOIR_0 + 4).shape[:2] == (scaled(80, ratio), scaled(112, ratio))


def test_upscale_ratio_city_textured():
    b = Bench("city", 128, 72, config="city_4k", upscale_ratio=1.7)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    for f in range(1, 6):
        inp = b.moving_inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)


UPSCALER_CASES = [
    # scene, config, size, settings, planes produced
    ("cornell", "cornell_1080p", (112, 80), dict(taa=plugin.TAA_JASMINE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=1.0),
     [L.OUT_UPSCALED, L.OUT_TAA]),                                           # Upscale::SMAA_TU_1_0 + Taa::Jasmine
    ("cornell", "cornell_1080p", (112, 80), dict(taa=plugin.TAA_JASMINE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=2.0),
     [L.OUT_UPSCALED, L.OUT_TAA]),                                           # HikariSettings::default(): Upscale::SMAA_TU_2_0 + Taa::Jasmine (lib.rs:435-455,494-497)
    ("cornell", "cornell_1080p", (90, 50), dict(taa=plugin.TAA_NONE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=1.5),
     [L.OUT_UPSCALED]),
    ("cornell", "cornell_1080p", (90, 50), dict(taa=plugin.TAA_JASMINE, upscale_kind=plugin.UPSCALE_FSR1, upscale_ratio=1.0),
     [L.OUT_TAA]),                                                           # TAA on the tone-mapped image (FSR pass not run)
    ("city", "city_4k", (128, 72), dict(taa=plugin.TAA_JASMINE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=2.0),
     [L.OUT_UPSCALED, L.OUT_TAA]),
]


@pytest.mark.parametrize("scene,config,size,settings,outputs", UPSCALER_CASES)
def test_temporal_upscalers_bit_exact(scene, config, size, settings, outputs):
    b = Bench(scene, size[0], size[1], config=config, **settings)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    ratio = settings["upscale_ratio"]
    rw, rh = scaled(size[0], ratio), scaled(size[1], ratio)
    for f in range(1, 10):
        inp = b.moving_inputs(f) if f != 5 else b.moving_inputs(f, step=(0.0, 0.0, 0.0))   # one frame of zero velocity too
        inp.temporal_upscalers = 1
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED + outputs, f)
    if L.OUT_UPSCALED in outputs:
        up = dev.readback(L.OUT_UPSCALED)
        assert up.shape[:2] == (upscaled(size[1], ratio), upscaled(size[0], ratio))
        assert np.isfinite(up.astype(np.float32)).all() and float(up[..., :3].astype(np.float32).max()) > 0.05
    if L.OUT_TAA in outputs:
        smaa = settings["upscale_kind"] == plugin.UPSCALE_SMAA_TU4X
        th, tw = (upscaled(size[1], ratio), upscaled(size[0], ratio)) if smaa else (rh, rw)
        taa = dev.readback(L.OUT_TAA)
        assert taa.shape[:2] == (th, tw)
        ptr, nbytes = dev.output_device_pointer(L.OUT_TAA)
        assert ptr and nbytes == th * tw * 8


def test_upscalers_nodes_one_by_one_and_plugin_switch():
    """hk_prepass_run + hk_light_run + hk_post_process_run == hk_render_frame with the upscalers on, and the host mirror's
    HikariPlugin.set_temporal_upscalers drives the same path through run_frame."""
    b = Bench("cornell", 80, 48, config="cornell_1080p", taa=plugin.TAA_JASMINE, upscale_ratio=1.5)
    a, c, d = b.device(), b.device(), b.device()
    d.set_temporal_upscalers(True)
    for f in range(1, 6):
        inp = b.inputs(f)
        inp.temporal_upscalers = 1
        a.render_frame(inp)
        c.prepass(inp); c.light(inp); c.post_process(inp)
        d.run_frame(b.settings, b.view, b.previous_view, b.lights)
        assert d.frame_counter == f
        for k in (L.OUT_TONE_MAPPED, L.OUT_UPSCALED, L.OUT_TAA):
            assert mismatch(a.readback(k), c.readback(k)) == 0, (f, k)
            assert mismatch(a.readback(k), d.readback(k)) == 0, (f, k)


def test_tiles_reject_scaled_rendering():
    b = Bench("cornell", 64, 64, config="cornell_256")
    tile = b.device(0, 32)
    inp = b.inputs(1)
    inp.temporal_upscalers = 1
    with pytest.raises(_ffi.HikariError, match="full-frame"):
        tile.render_frame(inp)
    inp = Bench("cornell", 64, 64, config="cornell_256", upscale_ratio=2.0).inputs(1)
    with pytest.raises(_ffi.HikariError, match="full-frame"):
        tile.render_frame(inp)
    inp = b.inputs(1)
    inp.frame.upscale_ratio = 2.5
    with pytest.raises(_ffi.HikariError, match="upscale_ratio"):
        b.device().render_frame(inp)


def test_committed_fixtures_from_cuda():
    """The CUDA path reproduces both committed fixtures (tests/golden/, written by tools/make_golden.py from the oracle)
    byte for byte — the check that needs nothing but the repository on the GPU box."""
    import os
    from tests.conftest import ROOT
    z = np.load(os.path.join(ROOT, "tests", "golden", "cornell_48x48_cfg2_frames1-6.npz"))
    b = Bench("cornell", 48, 48, config="cornell_1080p")
    dev = b.device()
    for f in range(1, 7):
        dev.render_frame(b.inputs(f))
    for name, which in (("tone_mapped", L.OUT_TONE_MAPPED), ("position", L.OUT_GBUFFER_POSITION), ("instance_material", L.OUT_GBUFFER_INSTANCE_MATERIAL),
                        ("reservoir9", L.OUT_RESERVOIR_0 + 9), ("render_indirect", L.OUT_RENDER_INDIRECT)):
        assert np.array_equal(np.ascontiguousarray(dev.readback(which)).view(np.uint8).reshape(-1), z[name]), name
    z = np.load(os.path.join(ROOT, "tests", "golden", "cornell_48x40_ratio1.5_smaa_taa_frames1-6.npz"))
    b = Bench("cornell", 48, 40, config="cornell_1080p", taa=plugin.TAA_JASMINE, upscale_ratio=1.5)
    dev = b.device()
    for f in range(1, 7):
        inp = b.moving_inputs(f)
        inp.temporal_upscalers = 1
        dev.render_frame(inp)
    for name, which in (("tone_mapped", L.OUT_TONE_MAPPED), ("upscaled", L.OUT_UPSCALED), ("taa", L.OUT_TAA), ("reservoir9", L.OUT_RESERVOIR_0 + 9)):
        assert np.array_equal(np.ascontiguousarray(dev.readback(which)).view(np.uint8).reshape(-1), z[name]), name
