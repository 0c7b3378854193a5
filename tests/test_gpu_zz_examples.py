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
