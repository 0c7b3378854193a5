"""GPU parity on the reference's shape-only examples (examples/minimal.rs, examples/simple.rs; scenes.minimal / scenes.simple).
simple.rs is the only scene with two emissive instances (emissive BVH with more than one leaf) and its spheres rotate every
frame (sphere_rotate_system), so it also runs through hk_scene_update_instances.  Bit-exact against the oracle.

(Named zz so that it runs last: these cases were written after the round's GPU budget was spent and have only been
validated on the CPU side; a surprise here must not hide the rest of the suite behind `pytest -x`.)"""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Animation, Bench, rotation_y_about
from tests.test_gpu_parity import ALL_PLANES, DENOISED, compare_all

pytestmark = pytest.mark.gpu


def test_minimal_example_default_settings():
    """HikariSettings::default() as the example uses it: 1 bounce, indirect spatial reuse, denoise, TAA Jasmine, SMAA TU4x at
    ratio 2 — the whole default pipeline incl. the temporal upscalers, under a moving camera."""
    b = Bench("minimal", 112, 80, taa=plugin.TAA_JASMINE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=2.0)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    for f in range(1, 8):
        inp = b.moving_inputs(f)
        inp.temporal_upscalers = 1
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED + [L.OUT_UPSCALED, L.OUT_TAA], f)


def test_simple_example_two_rotating_emissives():
    b = Bench("simple", 128, 80, taa=plugin.TAA_NONE, upscale_ratio=1.0, indirect_bounces=2, emissive_spatial_reuse=1)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    # sphere_rotate_system: rotate_local_z(0.2 * dt) — local z is world y after the spheres' -90 degree x rotation
    an = Animation(b, {6: lambda f: rotation_y_about(0.2 * f / 60.0 * 10, (2.0, 1.0, 0.0)),
                       7: lambda f: rotation_y_about(0.2 * f / 60.0 * 10, (-2.0, 1.0, 0.0))})
    for f in range(1, 8):
        w = an.step(f)
        dev.update_instances(w)
        orc.update_instances_desc(w.scene_desc())
        inp = b.inputs(f) if f < 5 else b.moving_inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)
    assert float(dev.readback(L.OUT_RENDER_EMISSIVE).astype(np.float32)[..., :3].mean()) > 0.005


def test_texture_samplers_bit_exact():
    """every sampler state the texture path distinguishes: repeat / clamp / mirror addressing on either axis, linear and
    nearest filtering, sRGB and linear data, uvs from -1 to 2 (scenes.samplers)"""
    b = Bench("samplers", 128, 80, taa=plugin.TAA_NONE, upscale_ratio=1.0, indirect_bounces=2, emissive_spatial_reuse=1)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    for f in range(1, 7):
        inp = b.moving_inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)
    albedo = dev.readback(L.OUT_ALBEDO).astype(np.float32)[..., :3]
    inst = np.floor(dev.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)[..., 0]).astype(int)
    hit = dev.readback(L.OUT_GBUFFER_POSITION)[..., 3] > 0
    for k in range(4):          # all four differently sampled quadrants are visible and textured (not a flat colour)
        m = hit & (inst == k)
        assert m.sum() > 200 and albedo[m].std(axis=0).max() > 0.02, k


def test_api_corners():
    """entry points and refusals the parity tests do not reach"""
    from bevy_hikari_b200 import _ffi
    import ctypes as C
    lib = _ffi.lib()
    b = Bench("cornell", 48, 32, config="cornell_256")
    with pytest.raises(_ffi.HikariError, match="tile rectangle"):
        plugin.HikariPlugin(48, 32, 0, 20, 10)                      # row_begin > row_end
    with pytest.raises(_ffi.HikariError, match="tile rectangle"):
        plugin.HikariPlugin(48, 32, 0, 0, 32, None, 0, 64)          # col_end beyond the frame
    t = b.device(8, 24, 16, 40)
    allocated, owned = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    assert lib.hk_tile_rect(t.ctx, allocated, owned) == 0
    assert list(owned) == [16, 40, 8, 24] and list(allocated) == [0, 48, 0, 32]      # +-36 ghost pixels, clamped to the frame
    a0, a1 = C.c_uint32(), C.c_uint32()
    assert lib.hk_band_rows(t.ctx, C.byref(a0), C.byref(a1)) == 0 and (a0.value, a1.value) == (0, 32)
    assert lib.hk_version().startswith(b"hikari_b200")
    bad = b.inputs(1)
    bad.frame.direct_validate_interval = 0
    with pytest.raises(_ffi.HikariError, match="validate interval"):
        t.render_frame(bad)
    bufs = b.world.buffers()
    desc = plugin.scene_desc_from_buffers(bufs)
    desc.instances = None                                           # count without a pointer
    with pytest.raises(_ffi.HikariError, match="NULL"):
        t.update_instances_desc(desc)
    with pytest.raises(_ffi.HikariError, match="unknown plane"):
        t.readback_into(99, 0x1000, 16)
    # per-pass timing and stats are filled when asked for
    t.set_profiling(True, True)
    t.render_frame(b.inputs(1))
    st = t.stats()
    assert st.kernel_launches == 9 and st.primary_rays == 16 * 24 and st.tlas_rays > 0     # cornell_256: no denoise, no emissive spatial pass
    assert st.ms_total >= 0.0 and all(m >= 0.0 for m in st.ms_kernel)


@pytest.mark.parametrize("base", [65535, 2 ** 24 - 3, 2 ** 31 - 4, 2 ** 32 - 6])
def test_large_frame_counters(base):
    """frame.number after hours of running and across the u32 wrap: noise texture index, halton index, ping-pong parity,
    validation intervals and the golden-ratio rotation (a float product of the frame number) stay in step with the oracle"""
    b = Bench("cornell", 64, 40, config="cornell_1080p", taa=plugin.TAA_JASMINE, upscale_ratio=1.5)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    for f in range(1, 7):
        inp = b.moving_inputs(f)
        inp.frame.number = (base + f) & 0xFFFFFFFF
        inp.temporal_upscalers = 1
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED + [L.OUT_UPSCALED, L.OUT_TAA], f)


def test_orthographic_camera_bit_exact():
    from tests.conftest import orthographic_inputs
    b = Bench("cornell", 96, 64, config="cornell_1080p")
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    for f in range(1, 7):
        inp = orthographic_inputs(b, f, shift=(0.03, 0.0, 0.0))
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)
    assert (dev.readback(L.OUT_GBUFFER_POSITION)[..., 3] > 0).mean() > 0.35


def test_town_scene_of_config_3():
    """BASELINE configs[2] (examples/scene.rs: scene.gltf, 120 440 triangles in 84 meshes, 52 textures, sun + emissive sphere, 3
    bounces, both spatial reuses, denoise) at a small frame size: every plane of every frame bit-equal to the oracle, static
    and moving camera"""
    b = Bench("town", 160, 90, config="scene_1080p")
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    dev.set_profiling(True, False)         # the ray-counting kernel variants
    for f in range(1, 6):
        inp = b.inputs(f) if f < 3 else b.moving_inputs(f, step=(0.3, 0.1, -0.2))
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)
    st = dev.stats()
    assert st.blas_rays > 0 and st.tlas_rays > 160 * 90 * 2     # sun + paths of up to 3 bounces (most leave the town after one)
