"""Regressions found after the round's last full device run, by the randomised campaign of tools/check_all.sh (fresh seeds) on the
emulated kernels; host-side logic of the context only (sorts last in the device suite like the other late additions; green on a B200
in call 19, the round's last GPU seconds)."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench, cornell_animation
from tests.test_gpu_parity import ALL_PLANES, DENOISED, compare_all

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,ratio", [((3, 4), 1.25), ((4, 2), 1.3), ((1, 1), 2.0)])
def test_scaled_rendering_whose_render_size_equals_the_frame(size, ratio):
    """ceil(3 / 1.25) = 3: the frame runs the scaled path (tight render-size planes, jittered look-ups) at a render size equal to
    the frame size.  hk_readback used to tell the two layouts apart by comparing sizes and read such planes with the deferred pitch
    (seed 229086 of tools/fuzz_parity.py); it now remembers which path the frame took."""
    b = Bench("cornell", size[0], size[1], config="cornell_1080p", upscale_ratio=ratio, upscale_kind=plugin.UPSCALE_SMAA_TU4X, taa=plugin.TAA_NONE)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    anim = cornell_animation(b)
    for f in range(1, 5):
        anim.step(f)
        dev.update_instances(b.world); orc.update_instances_desc(b.world.scene_desc())
        inp = b.inputs(f)
        inp.temporal_upscalers = 1
        dev.render_frame(inp); orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED + [L.OUT_UPSCALED], f)
    assert dev.readback(L.OUT_TONE_MAPPED).shape[:2] == (size[1], size[0])
