"""The contract of the PRODUCT library (libhikari_b200.so; tolerance build of the translation units that trace no rays — spatial
reuse, demodulation, denoise, tone mapping — bevy_hikari_b200/build.py) against the CPU oracle.  SURVEY 8(c) / BASELINE north_star:

  * integer outputs bit-exact: G-buffer instance / material ids, every plane the tolerance units do not write (G-buffer, albedo,
    sun radiance, the temporal reservoir buffers of all three signals — reservoirs carry visible_instance) — over whole sequences,
    because the temporal chain never reads what the tolerance units write.  The product walks the 4-wide trees by default
    (HK_TUNE_WIDE_TRAVERSAL, csrc/hk_wide.cuh): a reservoir record whose sample carries no radiance may then hold the position of a
    different occluder (test_gpu_parity.reservoir_mismatch); every other byte is the reference walk's;
  * every pass of a tolerance unit FROM IDENTICAL INPUTS (the oracle's planes uploaded before the pass): Rgba16Float outputs within
    1 f16 ulp, fewer than 1e-4 of the pixels outside that (a discrete decision — reservoir replacement, a rejection threshold — that
    falls the other way under a 1e-7 perturbation), fp32 reservoir fields within 4 ulp on the agreeing pixels;
  * bounded drift: 64 frames free-running (the product's own spatial history, never re-synchronised) against the oracle — PSNR of
    the tone-mapped frame stated and floored.
With HK_EMULATE_KERNELS=1 the "product" is the exact emulation and every bound is met with zero difference (checks the harness)."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from tests.conftest import Bench
from tests.test_gpu_parity import mismatch, reservoir_mismatch

pytestmark = pytest.mark.gpu

EXACT_PLANES = ([L.OUT_GBUFFER_POSITION, L.OUT_GBUFFER_NORMAL, L.OUT_GBUFFER_DEPTH_GRADIENT, L.OUT_GBUFFER_INSTANCE_MATERIAL,
                 L.OUT_GBUFFER_VELOCITY_UV, L.OUT_ALBEDO, L.OUT_RENDER_DIRECT, L.OUT_VARIANCE_DIRECT] +
                [L.OUT_RESERVOIR_0 + i for i in (0, 1, 2, 3, 6, 7)])          # temporal reservoirs: direct, emissive, indirect
STATE = ([L.OUT_GBUFFER_POSITION, L.OUT_GBUFFER_NORMAL, L.OUT_GBUFFER_DEPTH_GRADIENT, L.OUT_GBUFFER_INSTANCE_MATERIAL, L.OUT_GBUFFER_VELOCITY_UV,
          L.OUT_ALBEDO, L.OUT_RENDER_DIRECT, L.OUT_RENDER_EMISSIVE, L.OUT_RENDER_INDIRECT, L.OUT_VARIANCE_DIRECT, L.OUT_VARIANCE_EMISSIVE,
          L.OUT_VARIANCE_INDIRECT] + [L.OUT_RESERVOIR_0 + i for i in range(10)])


def f16_ulp_distance(a, b):
    """per element: distance in f16 representable values between two float16 arrays (inf for NaN mismatches)"""
    def key(x):
        u = np.ascontiguousarray(x).view(np.uint16).astype(np.int32)
        return np.where(u & 0x8000, 0x8000 - u, u)          # monotone integer key of the f16 value (-0 == +0)
    d = np.abs(key(a) - key(b)).astype(np.float64)
    nan_a, nan_b = np.isnan(a.astype(np.float32)), np.isnan(b.astype(np.float32))
    d[nan_a & nan_b] = 0
    d[nan_a ^ nan_b] = np.inf
    return d


def outlier_pixels(dev_plane, orc_plane, ulps=1):
    d = f16_ulp_distance(dev_plane, orc_plane)
    return int((d > ulps).any(axis=-1).sum()), float(np.where(np.isfinite(d), d, 0).max())


def sync_state(dev, orc, planes=STATE):
    for k in planes:
        dev.upload_state(k, orc.readback(k))


@pytest.mark.parametrize("scene,config,size", [("cornell", "cornell_1080p", (192, 128)), ("city", "city_4k", (160, 96))])
def test_tolerance_units_per_pass_from_identical_inputs(scene, config, size):
    W, H = size
    b = Bench(scene, W, H, config=config, emissive_spatial_reuse=1)
    dev, orc = b.device(flavor="product"), b.oracle()
    dev.set_keep_intermediates(True)
    npix = W * H
    budget = max(1, int(1e-4 * npix) + 1)          # < 1e-4 of the pixels (+1 so that tiny test frames may have a single one)
    worst = {}
    for f in range(1, 9):
        inp = b.inputs(f) if f < 5 else b.moving_inputs(f)
        # the exact units run on the device from the oracle's state of the previous frame and must reproduce it bit for bit
        dev.prepass(inp); orc.prepass(inp)
        orc.run_pass(inp, 0); dev.run_pass(inp, 0)
        for p in (1, 2):
            orc.run_pass(inp, p); dev.run_pass(inp, p)
        for k in EXACT_PLANES[:8]:
            assert mismatch(dev.readback(k), orc.readback(k)) == 0, (f, "exact unit", k)
        for k in [L.OUT_RESERVOIR_0 + i for i in (0, 1, 2, 3)]:
            assert reservoir_mismatch(dev.readback(k), orc.readback(k)) == 0, (f, "exact unit", k)
        # ---- pass 3: spatial reuse (emissive), from identical inputs
        sync_state(dev, orc)
        orc.run_pass(inp, 3); dev.run_pass(inp, 3)
        n, m = outlier_pixels(dev.readback(L.OUT_RENDER_EMISSIVE), orc.readback(L.OUT_RENDER_EMISSIVE))
        worst[(f, "spatial emissive render")] = (n, m)
        assert n <= budget, (f, "emissive spatial", n, m)
        sync_state(dev, orc)
        orc.run_pass(inp, 4); dev.run_pass(inp, 4)
        for k in (L.OUT_RENDER_INDIRECT, L.OUT_VARIANCE_INDIRECT):
            assert mismatch(dev.readback(k), orc.readback(k)) == 0, (f, "indirect (exact unit)", k)
        for k in (L.OUT_RESERVOIR_0 + 6, L.OUT_RESERVOIR_0 + 7):
            assert reservoir_mismatch(dev.readback(k), orc.readback(k)) == 0, (f, "indirect (exact unit)", k)
        # ---- pass 5: spatial reuse (indirect)
        sync_state(dev, orc)
        orc.run_pass(inp, 5); dev.run_pass(inp, 5)
        n, m = outlier_pixels(dev.readback(L.OUT_RENDER_INDIRECT), orc.readback(L.OUT_RENDER_INDIRECT))
        worst[(f, "spatial indirect render")] = (n, m)
        assert n <= budget, (f, "indirect spatial", n, m)
        # the spatial reservoir it wrote: same sample chosen (bit-identical record) on all but the outlier pixels
        cur = (f & 1)
        rd, ro = dev.readback(L.OUT_RESERVOIR_0 + 8 + (1 - cur)), orc.readback(L.OUT_RESERVOIR_0 + 8 + (1 - cur))
        same_sample = (rd["sample_position"] == ro["sample_position"]).all(axis=-1) if rd.dtype.names else None
        if same_sample is not None:
            assert (~same_sample).sum() <= budget, (f, "spatial reservoir picked another sample", int((~same_sample).sum()))
        # ---- pass 6 + 7: denoise chain and tone mapping
        sync_state(dev, orc)
        for sgl in range(3):
            orc.run_pass(inp, 6, sgl)
        dev.run_pass(inp, 6)
        for k in (L.OUT_DENOISED_DIRECT, L.OUT_DENOISED_EMISSIVE, L.OUT_DENOISED_INDIRECT):
            n, m = outlier_pixels(dev.readback(k), orc.readback(k), ulps=4)       # four a-trous levels: 1 f16 ulp each
            worst[(f, "denoised", k)] = (n, m)
            assert n <= budget, (f, "denoise chain", k, n, m)
        orc.run_pass(inp, 7); dev.run_pass(inp, 7)
        n, m = outlier_pixels(dev.readback(L.OUT_TONE_MAPPED), orc.readback(L.OUT_TONE_MAPPED), ulps=5)
        assert n <= budget, (f, "tone mapping", n, m)
        sync_state(dev, orc)          # next frame starts from the oracle's state on both sides
    print({str(k): v for k, v in worst.items() if v[0]})


def test_ids_and_temporal_chain_stay_bit_exact_over_a_free_running_sequence():
    """16 frames, static then moving camera, the product renders on its own (no re-synchronisation): everything outside the
    tolerance units' outputs equals the oracle bit for bit in every frame."""
    b = Bench("cornell", 160, 96, config="cornell_1080p")
    dev, orc = b.device(flavor="product"), b.oracle()
    for f in range(1, 17):
        inp = b.inputs(f) if f < 7 else b.moving_inputs(f)
        dev.render_frame(inp); orc.render_frame(inp)
        for k in EXACT_PLANES:
            same = reservoir_mismatch if k >= L.OUT_RESERVOIR_0 else mismatch
            assert same(dev.readback(k), orc.readback(k)) == 0, (f, k)


def test_drift_over_64_frames_is_bounded():
    """64 frames free-running at the oracle's side and the product's side: the tone-mapped frame's PSNR stays above 50 dB and the
    share of pixels beyond 2 f16 ulp stays small (the product's spatial history differs from the oracle's by its own rounding and by
    the rare discrete decision that fell the other way; the estimator is the same)."""
    b = Bench("cornell", 160, 96, config="cornell_1080p")
    dev, orc = b.device(flavor="product"), b.oracle()
    psnr, beyond = [], []
    for f in range(1, 65):
        inp = b.inputs(f)
        dev.render_frame(inp); orc.render_frame(inp)
        if f % 8 == 0:
            a, o = dev.readback(L.OUT_TONE_MAPPED)[..., :3].astype(np.float64), orc.readback(L.OUT_TONE_MAPPED)[..., :3].astype(np.float64)
            mse = float(((a - o) ** 2).mean())
            psnr.append(99.0 if mse == 0 else 10 * np.log10(1.0 / mse))
            n, _ = outlier_pixels(dev.readback(L.OUT_TONE_MAPPED), orc.readback(L.OUT_TONE_MAPPED), ulps=2)
            beyond.append(n / (160 * 96))
    print("PSNR (dB) every 8th frame:", [round(p, 1) for p in psnr], "share beyond 2 ulp:", [round(x, 5) for x in beyond])
    assert min(psnr) > 50.0, psnr
    assert max(beyond) < 0.02, beyond
