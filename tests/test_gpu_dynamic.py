"""GPU parity with animated instances: every frame the host mirror rebuilds instances / TLAS / emissives / alias tables
(instance.rs:352-437) and hands them over through hk_scene_update_instances; the G-buffer's motion vectors use the
previous model matrices (prepass.wgsl:52,99).  Bit-exact against the oracle on every plane, as everywhere else."""
import numpy as np
import pytest

from bevy_hikari_b200 import _ffi
from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench, city_animation, cornell_animation
from tests.test_gpu_parity import ALL_PLANES, DENOISED, compare_all, mismatch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene,config,size,animation", [("cornell", "cornell_1080p", (112, 80), cornell_animation),
                                                        ("city", "city_4k", (128, 72), city_animation)])
def test_animated_instances_bit_exact(scene, config, size, animation):
    b = Bench(scene, size[0], size[1], config=config)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    an = animation(b)
    moved = 0
    for f in range(1, 10):
        w = an.step(f)
        dev.update_instances(w)
        orc.update_instances_desc(w.scene_desc())
        inp = b.moving_inputs(f) if f > 5 else b.inputs(f)      # instances move under a static, then under a moving camera
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)
        if f <= 5:
            moved += int((np.abs(dev.readback(L.OUT_GBUFFER_VELOCITY_UV)[..., :2]) > 0).any(axis=2).sum())
    assert moved > 100


def test_animated_instances_with_upscalers():
    """default HikariSettings pipeline on a dynamic scene: scaled rendering + SMAA TU4x + TAA consume the motion vectors"""
    b = Bench("cornell", 96, 64, config="cornell_1080p", taa=plugin.TAA_JASMINE, upscale_ratio=2.0)
    dev, orc = b.device(), b.oracle()
    an = cornell_animation(b)
    for f in range(1, 8):
        w = an.step(f)
        dev.update_instances(w)
        orc.update_instances_desc(w.scene_desc())
        inp = b.inputs(f)
        inp.temporal_upscalers = 1
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + [L.OUT_UPSCALED, L.OUT_TAA], f)


def test_update_instances_equals_full_upload():
    b = Bench("city", 96, 54, config="city_4k")
    a, c = b.device(), b.device()
    an = city_animation(b)
    for f in range(1, 6):
        w = an.step(f)
        a.update_instances(w)
        c.upload_scene(w)
        inp = b.inputs(f)
        a.render_frame(inp); c.render_frame(inp)
        for k in ALL_PLANES:
            assert mismatch(a.readback(k), c.readback(k)) == 0, (f, k)


def test_update_instances_errors():
    b = Bench("cornell", 32, 32, config="cornell_256")
    p = plugin.HikariPlugin(32, 32)
    with pytest.raises(_ffi.HikariError, match="hk_scene_upload"):
        p.update_instances(b.world)                  # no scene yet
    p.upload_scene(b.world)
    bufs = b.world.buffers()
    bad = bufs["instances"].copy()
    bad["material"][3] = 1000
    bufs["instances"] = bad
    with pytest.raises(_ffi.HikariError, match="out of bounds"):
        p.update_instances_desc(plugin.scene_desc_from_buffers(bufs))
    p.render_frame(b.inputs(1))                      # the scene of the last good upload is still in place
    p.sync()
    # tiles accept updates too (each rank updates its replica of the scene)
    t = b.device(0, 16)
    an = cornell_animation(b)
    t.update_instances(an.step(1))
    t.render_frame(b.inputs(1))
    t.sync()


def test_animated_materials_bit_exact():
    """A StandardMaterial modified every frame (the light's emissive colour pulses, a wall changes colour and roughness):
    materials, emissive list and alias tables are rebuilt on the host and replaced through hk_scene_update_instances."""
    b = Bench("cornell", 96, 64, config="cornell_1080p")
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    w = b.world
    mats = b.scene.materials.copy()
    for f in range(1, 9):
        light = mats[4].copy()
        light["emissive"] = (1.0, 0.6 + 0.05 * f, 0.3, 0.5 + 0.06 * f) if f != 5 else (0.0, 0.0, 0.0, 1.0)   # frame 5: the light is switched off
        wall = mats[3].copy()
        wall["base_color"] = (0.2 + 0.1 * f, 0.9 - 0.1 * f, 0.3, 1.0)
        wall["perceptual_roughness"] = 0.1 * f
        w.set_material(4, light); w.set_material(3, wall)
        w.prepare_materials(); w.previous_transform_system(); w.prepare_instances()
        assert len(w.buffers()["emissives"]) == (0 if f == 5 else 1)
        dev.update_instances(w)
        orc.update_instances_desc(w.scene_desc())
        inp = b.inputs(f)
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED, f)
