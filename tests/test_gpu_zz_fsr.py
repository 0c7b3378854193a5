"""GPU parity for Upscale::Fsr1: FSR 1.0 EASU + RCAS after tone mapping / TAA
(post_process.rs:1279-1308; algorithm = src/shaders/fsr/source.zip, the sources of the reference's SPIR-V blobs).  Same bar
as the rest of the path: the EASU image (upscale_output[0]) and the RCAS image (upscale_output[1], what the overlay
presents) bit-equal to the CPU oracle, next to every plane upstream.

(Named zz so that it runs last: written after the round's GPU budget was spent and validated on the emulated kernels
only; a surprise here must not hide the rest of the suite behind `pytest -x`.)"""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_gpu_parity import ALL_PLANES, DENOISED, compare_all

pytestmark = pytest.mark.gpu

FSR_CASES = [
    # scene, config, size, taa, ratio, sharpness
    ("cornell", "cornell_1080p", (112, 80), plugin.TAA_NONE, 1.5, 0.0),       # sharpest
    ("cornell", "cornell_1080p", (112, 80), plugin.TAA_JASMINE, 2.0, 0.2),    # TAA at render size feeds EASU
    ("cornell", "cornell_1080p", (90, 50), plugin.TAA_JASMINE, 1.3, 1.0),     # odd sizes: ceil(size / ratio)
    ("cornell", "cornell_1080p", (64, 48), plugin.TAA_NONE, 1.0, 2.0),        # ratio 1: EASU resamples on the texel grid
    ("city", "city_4k", (128, 72), plugin.TAA_JASMINE, 1.7, 0.5),             # textured, sun
]


@pytest.mark.parametrize("scene,config,size,taa,ratio,sharpness", FSR_CASES)
def test_fsr1_bit_exact(scene, config, size, taa, ratio, sharpness):
    b = Bench(scene, size[0], size[1], config=config, taa=taa, upscale_kind=plugin.UPSCALE_FSR1, upscale_ratio=ratio,
              upscale_sharpness=sharpness)
    dev, orc = b.device(), b.oracle()
    dev.set_keep_intermediates(True)
    outputs = [L.OUT_UPSCALED, L.OUT_FSR_SHARPENED] + ([L.OUT_TAA] if taa == plugin.TAA_JASMINE else [])
    for f in range(1, 8):
        inp = b.moving_inputs(f)
        assert inp.fsr1 == 1 and inp.smaa_tu4x == 0 and abs(inp.fsr_sharpness - sharpness) < 1e-7
        inp.temporal_upscalers = 1
        dev.render_frame(inp)
        orc.render_frame(inp)
        compare_all(dev, orc, ALL_PLANES + DENOISED + outputs, f)
    for which in (L.OUT_UPSCALED, L.OUT_FSR_SHARPENED):
        img = dev.readback(which)
        assert img.shape[:2] == (size[1], size[0])                           # the camera target size
        f32 = img.astype(np.float32)
        assert np.isfinite(f32).all() and float(f32[..., :3].max()) > 0.05 and (f32[..., 3] == 1.0).all()
    st = dev.stats()
    assert st.kernel_launches >= 16      # 14 of the hot path + EASU + RCAS (+ TAA)


def test_fsr1_rejected_where_it_cannot_run():
    """FSR1 together with smaa_tu4x (two variants of one enum) and FSR1 on a tile are errors, not silent fall-backs"""
    b = Bench("cornell", 64, 48, config="cornell_1080p", taa=plugin.TAA_NONE, upscale_kind=plugin.UPSCALE_FSR1, upscale_ratio=1.0)
    dev = b.device()
    inp = b.inputs(1)
    inp.temporal_upscalers = 1
    inp.smaa_tu4x = 1
    with pytest.raises(RuntimeError):
        dev.render_frame(inp)
    tile = b.device(row_begin=0, row_end=24)
    inp = b.inputs(1)
    inp.temporal_upscalers = 1
    with pytest.raises(RuntimeError):
        tile.render_frame(inp)
    inp.temporal_upscalers = 0           # without the upscalers the flag is inert: the hot path ends at tone mapping
    tile.render_frame(inp)
