"""Context life-cycle rules of the C ABI: a resize to the same size keeps the temporal state, a real resize leaves
no stale derived state behind, HikariPlugin.run_frame runs the upscalers the settings select, malformed scene buffers are refused."""
import ctypes as C

import numpy as np
import pytest

from bevy_hikari_b200 import _ffi
from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench

pytestmark = pytest.mark.gpu


def test_resize_to_the_same_size_keeps_reservoirs_and_a_real_resize_clears_them():
    """light.rs:342-363 re-allocates the reservoir cache only when size.x * size.y changes; prepare_light_textures runs every frame."""
    b = Bench("cornell", 64, 48, config="cornell_1080p")
    d = b.device()
    for f in (1, 2, 3):
        d.render_frame(b.inputs(f))
    before = d.readback(L.OUT_RESERVOIR_0 + 7).copy()
    assert np.any(before.view(np.uint8) != 0)
    _ffi.check(_ffi.lib().hk_context_resize(d.ctx, 64, 48, 0, 48), d.ctx)
    assert d.readback(L.OUT_RESERVOIR_0 + 7).tobytes() == before.tobytes()
    # same sequence on a context that was never resized: frame 4 must agree bit for bit
    ref = b.device()
    for f in (1, 2, 3, 4):
        ref.render_frame(b.inputs(f))
    d.render_frame(b.inputs(4))
    assert d.readback(L.OUT_TONE_MAPPED).tobytes() == ref.readback(L.OUT_TONE_MAPPED).tobytes()
    # a real resize: zeroed planes, extents follow the new size at once (no frame needed), frame target dropped
    _ffi.check(_ffi.lib().hk_context_resize(d.ctx, 32, 24, 0, 24), d.ctx)
    assert d.output_extent(L.OUT_RENDER_DIRECT) == (32, 24) and d.output_extent(L.OUT_RESERVOIR_0) == (32, 24)
    d.width, d.height = 32, 24
    assert not np.any(d.readback(L.OUT_RESERVOIR_0 + 7).view(np.uint8))
    small = Bench("cornell", 32, 24, config="cornell_1080p")
    fresh = small.device()
    for f in (1, 2):
        d.render_frame(small.inputs(f)); fresh.render_frame(small.inputs(f))
    assert d.readback(L.OUT_TONE_MAPPED).tobytes() == fresh.readback(L.OUT_TONE_MAPPED).tobytes()


def test_shrink_after_scaled_rendering_reports_new_extents():
    """last_render_w/h of a ratio-1.5 frame must not survive a resize (they size read-backs of render planes)."""
    b = Bench("cornell", 96, 64, config="cornell_1080p", upscale_ratio=1.5)
    d = b.device()
    d.render_frame(b.inputs(1))
    assert d.output_extent(L.OUT_RENDER_DIRECT) == (64, 43)
    _ffi.check(_ffi.lib().hk_context_resize(d.ctx, 48, 32, 0, 32), d.ctx)
    assert d.output_extent(L.OUT_RENDER_DIRECT) == (48, 32)
    d.width, d.height = 48, 32
    assert d.readback(L.OUT_RENDER_DIRECT).shape[:2] == (32, 48)


def test_run_frame_runs_the_upscalers_the_settings_select():
    """HikariSettings::default() = Taa::Jasmine + SmaaTu4x{2.0}: the reference runs smaa_tu4x, extrapolate and taa_jasmine
    (post_process.rs:1236-1277); so does HikariPlugin::run_frame unless the caller opts out."""
    b = Bench("cornell", 64, 48)
    b.settings = plugin.HikariSettings(indirect_bounces=1)          # defaults otherwise
    d, e = b.device(), b.device()
    for f in range(1, 4):
        d.run_frame(b.settings, b.view, b.previous_view, b.lights)
        inp = b.inputs(f)
        inp.temporal_upscalers = 1
        e.render_frame(inp)
    assert d.output_extent(L.OUT_TONE_MAPPED) == (32, 24) and d.output_extent(L.OUT_TAA) == (64, 48)
    assert d.readback(L.OUT_TAA).tobytes() == e.readback(L.OUT_TAA).tobytes()
    assert np.any(d.readback(L.OUT_TAA).view(np.uint16))
    # opting out ends the path at the tone-mapped image and drops the jitter the upscalers would have resolved
    o, p = b.device(), b.device()
    o.set_temporal_upscalers(False)
    o.run_frame(b.settings, b.view, b.previous_view, b.lights)
    inp = b.inputs(1)
    inp.taa_jitter = 0
    p.render_frame(inp)
    assert o.readback(L.OUT_GBUFFER_POSITION).tobytes() == p.readback(L.OUT_GBUFFER_POSITION).tobytes()
    assert o.readback(L.OUT_TONE_MAPPED).tobytes() == p.readback(L.OUT_TONE_MAPPED).tobytes()


def test_malformed_scene_buffers_are_refused():
    b = Bench("cornell", 32, 32, config="cornell_256")
    bufs = b.world.buffers()
    d = plugin.HikariPlugin(32, 32)

    def upload(**changed):
        c = {k: v.copy() for k, v in bufs.items()}
        for k, f in changed.items():
            f(c[k])
        with pytest.raises(_ffi.HikariError, match="out of bounds|out of its mesh"):
            d.upload_scene_desc(plugin.scene_desc_from_buffers(c))

    def bad_tlas_leaf(a):
        leaves = np.nonzero(a["entry_index"] >= 0x80000000)[0]
        a["entry_index"][leaves[0]] = 0x80000000 + 1000
    upload(instance_nodes=bad_tlas_leaf)

    def bad_blas_leaf(a):
        leaves = np.nonzero(a["entry_index"] >= 0x80000000)[0]
        a["entry_index"][leaves[-1]] = 0x80000000 + 100000
    upload(asset_nodes=bad_blas_leaf)

    def bad_vertex(a):
        a["vertices"]["index"][0, 0] = 1 << 30
    upload(primitives=bad_vertex)

    def bad_alias(a):
        a["index"][0] = 1 << 20
    upload(alias_table=bad_alias)

    def bad_emissive(a):
        a["instance"][0] = 999
    upload(emissives=bad_emissive)
    d.upload_scene_desc(plugin.scene_desc_from_buffers(bufs))       # the untouched buffers are fine
    d.render_frame(b.inputs(1))


def test_import_gbuffer_from_device_pointers():
    """A host that rasterises its own prepass hands the five render targets over as device pointers with arbitrary row pitches
    (hk_import_gbuffer instead of hk_prepass_run): light + post-process on them equal the oracle's on the same G-buffer, frame after frame
    (the current <-> previous swap of the position / velocity planes included)."""
    from tests.conftest import EMULATED
    from tests.test_gpu_parity import ALL_PLANES, compare_all
    b = Bench("cornell", 72, 48, config="cornell_1080p")
    dev, orc = b.device(), b.oracle()
    planes = (L.OUT_GBUFFER_POSITION, L.OUT_GBUFFER_NORMAL, L.OUT_GBUFFER_DEPTH_GRADIENT, L.OUT_GBUFFER_INSTANCE_MATERIAL, L.OUT_GBUFFER_VELOCITY_UV)
    keep = []
    for f in range(1, 5):
        inp = b.moving_inputs(f)
        orc.prepass(inp)
        desc = []
        for k in planes:
            a = np.ascontiguousarray(orc.readback(k))
            row = a.reshape(a.shape[0], -1).view(np.uint8)
            padded = np.zeros((row.shape[0], row.shape[1] + 48), np.uint8)          # a pitch wider than the row, as an imported image has
            padded[:, :row.shape[1]] = row
            if EMULATED:
                keep.append(padded)
                desc.append((padded.ctypes.data, padded.shape[1]))
            else:
                import torch
                t = torch.from_numpy(padded).cuda()
                keep.append(t)
                desc.append((t.data_ptr(), padded.shape[1]))
        if not EMULATED:
            import torch
            torch.cuda.synchronize()
        dev.import_gbuffer(desc)
        dev.light(inp); dev.post_process(inp)
        orc.light(inp); orc.post_process(inp)
        compare_all(dev, orc, ALL_PLANES, f)
