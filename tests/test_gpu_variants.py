"""GPU parity across settings, sizes and API edge cases (all bit-exact against the oracle)."""
import ctypes as C

import numpy as np
import pytest

from bevy_hikari_b200 import _ffi
from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_gpu_parity import ALL_PLANES, DENOISED, compare_all, mismatch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("settings", [
    dict(indirect_bounces=0, denoise=1),                                        # third denoise dropped (post_process.rs:949-954)
    dict(indirect_bounces=1, denoise=1, emissive_spatial_reuse=1),
    dict(indirect_bounces=3, denoise=0),
    dict(indirect_bounces=4, denoise=1, indirect_spatial_reuse=0),
    dict(indirect_bounces=2, temporal_reuse=0, denoise=0),                      # reservoirs never stored (light.wgsl:1229-1231)
    dict(indirect_bounces=2, max_reservoir_lifetime=1.0, emissive_spatial_reuse=1, denoise=0),   # lifetime limit -> F32_MAX
    dict(indirect_bounces=2, direct_validate_interval=1, emissive_validate_interval=2, denoise=0),  # validation every frame
    dict(indirect_bounces=2, max_temporal_reuse_count=2, max_spatial_reuse_count=3, max_indirect_luminance=0.5, denoise=1,
         emissive_spatial_reuse=1),
    dict(indirect_bounces=2, solar_angle=0.5, clear_color=(0.1, 0.2, 0.3, 1.0), denoise=1),
])
def test_settings_variants_bit_exact(settings):
    b = Bench("cornell", 72, 56, **settings)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    planes = ALL_PLANES + (DENOISED if b.settings.denoise else [])
    if b.settings.denoise and b.settings.indirect_bounces == 0:
        planes = [p for p in planes if p != L.OUT_DENOISED_INDIRECT]
    for f in range(1, 8):
        inp = b.inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, planes, f)


@pytest.mark.parametrize("size", [(1, 1), (7, 3), (50, 37), (129, 17), (16, 8)])
def test_ragged_sizes_bit_exact(size):
    """Sizes that are not multiples of the 16x8 CTA tile, down to a single pixel."""
    b = Bench("cornell", size[0], size[1], config="cornell_1080p")
    dev, orc = b.device(), b.oracle()
    for f in range(1, 5):
        inp = b.inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES, f)


def test_sun_only_scene_without_emissives_and_empty_scene():
    from bevy_hikari_b200 import scenes
    city = scenes.city()
    # drop the emissive sphere: no emissive BVH at all (emissive_node_count == 0)
    keep = [i for i in range(len(city.inst_mesh)) if i != 1]
    city.inst_mesh = [city.inst_mesh[i] for i in keep]
    city.inst_material = [city.inst_material[i] for i in keep]
    city.inst_transform = [city.inst_transform[i] for i in keep]
    b = Bench.__new__(Bench)
    b.scene, b.width, b.height = city, 96, 54
    b.world = city.populate(plugin.World())
    assert len(b.world.buffers()["emissives"]) == 0
    b.view, b.previous_view, b.lights = city.view_inputs(96, 54)
    b.settings = scenes.config_settings("city_4k")
    dev, orc = b.device(), b.oracle()
    for f in range(1, 5):
        inp = b.inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES, f)
    # empty world: every pixel is background -> clear colour, and nothing crashes
    empty = plugin.World()
    empty.prepare()
    p = plugin.HikariPlugin(40, 24)
    p.upload_scene(empty)
    p.render_frame(b.inputs(1))
    tm = p.readback(L.OUT_TONE_MAPPED).astype(np.float32)
    assert np.allclose(tm[..., :3], 0.4, atol=2e-4) and np.all(tm[..., 3] == 1.0)


def test_single_passes_from_identical_uploaded_state():
    """Per-node parity from identical inputs: run 3 frames on the oracle, upload its whole state into a fresh device
    context, then run only the light node / only the post-process node on both and compare what that node writes."""
    b = Bench("cornell", 64, 48, config="cornell_1080p")
    orc = b.oracle()
    for f in range(1, 4):
        orc.render_frame(b.inputs(f))
    inp = b.inputs(4)
    orc.prepass(inp)
    dev = b.device()
    for k in [L.OUT_GBUFFER_POSITION, L.OUT_GBUFFER_NORMAL, L.OUT_GBUFFER_DEPTH_GRADIENT, L.OUT_GBUFFER_INSTANCE_MATERIAL,
              L.OUT_GBUFFER_VELOCITY_UV, L.OUT_ALBEDO] + [L.OUT_RESERVOIR_0 + i for i in range(10)]:
        dev.upload_state(k, orc.readback(k))
        assert mismatch(dev.readback(k), orc.readback(k)) == 0, k        # upload/readback round trip incl. reservoir re-layout
    dev.light(inp)
    orc.light(inp)
    compare_all(dev, orc, [p for p in ALL_PLANES if p != L.OUT_TONE_MAPPED], 4)
    dev.post_process(inp)
    orc.post_process(inp)
    compare_all(dev, orc, [L.OUT_TONE_MAPPED] + DENOISED, 4)


def test_resize_and_reset_zero_the_temporal_state():
    b = Bench("cornell", 48, 40, config="cornell_1080p")
    dev = b.device()
    for f in range(1, 4):
        dev.render_frame(b.inputs(f))
    assert any(dev.readback(L.OUT_RESERVOIR_0 + i).tobytes().strip(b"\0") for i in range(10))
    dev.reset_temporal_state()
    assert not any(dev.readback(L.OUT_RESERVOIR_0 + i).tobytes().strip(b"\0") for i in range(10))
    # resize: planes re-allocated and zeroed like ReservoirCache does when size.x*size.y changes (light.rs:342-363)
    _ffi.check(_ffi.lib().hk_context_resize(dev.ctx, 32, 24, 0, 24), dev.ctx)
    dev.width, dev.height, dev.row_begin, dev.row_end, dev.col_begin, dev.col_end = 32, 24, 0, 24, 0, 32
    b2 = Bench("cornell", 32, 24, config="cornell_1080p")
    fresh = b2.device()
    for f in range(1, 4):
        dev.render_frame(b2.inputs(f))
        fresh.render_frame(b2.inputs(f))
    for k in (L.OUT_TONE_MAPPED, L.OUT_RESERVOIR_0 + 9):
        assert mismatch(dev.readback(k), fresh.readback(k)) == 0


def test_scene_upload_rejects_out_of_range_references():
    b = Bench("cornell", 16, 16, config="cornell_256")
    bufs = b.world.buffers()
    bufs["instances"]["material"][3] = 1000
    dev = plugin.HikariPlugin(16, 16)
    with pytest.raises(_ffi.HikariError, match="out of bounds"):
        dev.upload_scene_desc(plugin.scene_desc_from_buffers(bufs))
    bad_size = np.zeros(10, np.uint8)
    assert _ffi.lib().hk_readback(dev.ctx, L.OUT_TONE_MAPPED, bad_size.ctypes.data, bad_size.size) == _ffi.HK_ERR_INVALID_ARGUMENT
    assert _ffi.lib().hk_readback(dev.ctx, 999, bad_size.ctypes.data, bad_size.size) == _ffi.HK_ERR_INVALID_ARGUMENT
    assert _ffi.lib().hk_render_frame(dev.ctx, None) == _ffi.HK_ERR_INVALID_ARGUMENT


def test_pipelined_readback_equals_blocking():
    """hk_readback_async / hk_readback_wait: frame n's image lands in pinned host memory while frame n + 1 renders, and is
    byte-identical to the blocking read-back of the same frame; the next frame's tone-map write waits for the copy."""
    from tests.conftest import needs_real_gpu
    needs_real_gpu()
    import torch
    b = Bench("cornell", 160, 96, config="cornell_1080p")
    a, c = b.device(), b.device()
    nbytes = 160 * 96 * 8
    bufs = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    expected = []
    for f in range(1, 8):
        inp = b.inputs(f)
        c.render_frame(inp)
        expected.append(np.ascontiguousarray(c.readback(L.OUT_TONE_MAPPED)).view(np.uint8).reshape(-1).copy())
        a.render_frame(inp)
        a.readback_wait()                                   # frame f - 1 is complete on the host ...
        if f > 1:
            assert np.array_equal(bufs[f & 1].numpy(), expected[f - 2]), f - 1
        a.readback_async(L.OUT_TONE_MAPPED, bufs[(f + 1) & 1].data_ptr(), nbytes)   # ... while frame f is being copied
    a.readback_wait()
    assert np.array_equal(bufs[(7 + 1) & 1].numpy(), expected[6])
    from bevy_hikari_b200 import _ffi
    with pytest.raises(_ffi.HikariError, match="final images"):
        a.readback_async(L.OUT_ALBEDO, bufs[0].data_ptr(), nbytes)
    with pytest.raises(_ffi.HikariError, match="size mismatch"):
        a.readback_async(L.OUT_TONE_MAPPED, bufs[0].data_ptr(), nbytes - 8)
