"""Demodulation, the four a-trous denoise levels and tone mapping, pinned from the outside:
a SECOND, independent restatement of src/shaders/denoise.wgsl:135-319 and tone_mapping.wgsl:21-32 — whole-image numpy
float32 arithmetic written from the WGSL, not from oracle/hk_oracle.cpp — fed with the planes the oracle renders for real
frames (cornell, city), must reproduce the oracle's denoised planes and tone-mapped image.  The two differ only in the
accuracy of exp / pow (the oracle and the kernels use the polynomial routines of hk_math.h, numpy uses libm), which the
Rgba16Float stores absorb almost everywhere.  Measured: 99.78 - 100 % of the texels bit-identical, never more than 1 f16 ulp
apart; demanded: >= 99.7 % identical and at most 4 texels per plane beyond 1 ulp (room for a threshold flip of the
firefly clamp).  CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench

F = np.float32
F32_MAX = F(3.402823466e38)
F32_EPSILON = F(1.1920929e-7)
KERNEL = np.array([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]], F)   # view.rs:125-129, symmetric


def f16(a):
    return a.astype(np.float16).astype(F)


def luminance(v):                                      # utils.wgsl: dot(v, (0.2126, 0.7152, 0.0722))
    return v[..., 0] * F(0.2126) + v[..., 1] * F(0.7152) + v[..., 2] * F(0.0722)


class Frame:
    """textures of one frame as float32 arrays + the addressing rules of denoise.wgsl"""

    def __init__(self, orc, number, ratio):
        self.number, self.ratio = number, F(ratio)
        self.position = orc.readback(L.OUT_GBUFFER_POSITION)
        self.normal = np.maximum(orc.readback(L.OUT_GBUFFER_NORMAL).astype(F) / F(127.0), F(-1.0))   # Rgba8Snorm
        self.depth_gradient = orc.readback(L.OUT_GBUFFER_DEPTH_GRADIENT)
        self.instance_material = orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)
        self.albedo = orc.readback(L.OUT_ALBEDO).astype(F)
        self.render = [orc.readback(L.OUT_RENDER_DIRECT + i).astype(F) for i in range(3)]
        self.variance = [orc.readback(L.OUT_VARIANCE_DIRECT + i) for i in range(3)]
        self.H, self.W = self.position.shape[:2]
        self.RH, self.RW = self.render[0].shape[:2]
        ys, xs = np.meshgrid(np.arange(self.RH), np.arange(self.RW), indexing="ij")
        self.xs, self.ys = xs, ys

    def coords_to_uv(self, x, y):                       # utils.wgsl coords_to_uv on the render-size output
        return (x.astype(F) + F(0.5)) / F(self.RW), (y.astype(F) + F(0.5)) / F(self.RH)

    def jittered_deferred_uv(self, u, v):               # denoise.wgsl:37-41
        sel = F(-0.5) if (self.number & 1) == 0 else F(0.5)
        return u + sel * (F(1.0) / F(self.W)) * (self.ratio - F(1.0)), v + sel * (F(1.0) / F(self.H)) * (self.ratio - F(1.0))

    def nearest_deferred(self, tex, u, v):              # nearest sampler, clamp-to-edge, full-size texture
        x = np.clip(np.floor(u * F(self.W)).astype(np.int64), 0, self.W - 1)
        y = np.clip(np.floor(v * F(self.H)).astype(np.int64), 0, self.H - 1)
        return tex[y, x]


def demodulation(fr, signal):                           # denoise.wgsl:135-162
    u, v = fr.coords_to_uv(fr.xs, fr.ys)
    du, dv = fr.jittered_deferred_uv(u, v)
    albedo = fr.nearest_deferred(fr.albedo, du, dv)[..., :3]
    irradiance = fr.render[signal][..., :3]
    with np.errstate(all="ignore"):
        irradiance = np.where(albedo < F(0.01), F(0.0), irradiance / albedo)
    internal0 = np.concatenate([irradiance, np.ones_like(irradiance[..., :1])], axis=-1)
    sum_variance = np.zeros((fr.RH, fr.RW), F)
    var = fr.variance[signal]
    for ox in (-1, 0, 1):                               # the order of the nine accumulate_variance calls
        for oy in (-1, 0, 1):
            su, sv = u + F(ox) / F(fr.RW), v + F(oy) / F(fr.RH)
            inside = ~((su < 0) | (sv < 0) | (su > 1) | (sv > 1))
            sx = np.clip(np.floor(su * F(fr.RW)).astype(np.int64), 0, fr.RW - 1)
            sy = np.clip(np.floor(sv * F(fr.RH)).astype(np.int64), 0, fr.RH - 1)
            s = var[sy, sx]
            ok = inside & ~(s > F32_MAX)
            sum_variance = np.where(ok, sum_variance + KERNEL[oy + 1, ox + 1] * np.fmax(s, F(0.0)), sum_variance)
    return f16(internal0), sum_variance


def denoise_level(fr, inp, variance, level, firefly):   # denoise.wgsl:164-319
    step = (8, 4, 2, 1)[level]
    u, v = fr.coords_to_uv(fr.xs, fr.ys)
    du, dv = fr.jittered_deferred_uv(u, v)
    depth = fr.nearest_deferred(fr.position, du, dv)[..., 3]
    gradient = fr.nearest_deferred(fr.depth_gradient, du, dv)
    with np.errstate(all="ignore"):
        n = fr.nearest_deferred(fr.normal, du, dv)[..., :3]
        normal = n / np.sqrt(n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1] + n[..., 2] * n[..., 2])[..., None]
        instance = fr.nearest_deferred(fr.instance_material, du, dv)[..., 0]
        irradiance = inp[..., :3]
        bad = np.isnan(irradiance).any(-1) | (irradiance > F32_MAX).any(-1)
        irradiance = np.where(bad[..., None], F(0.0), irradiance)
        sum_irr = np.where(bad[..., None], F(0.0), irradiance * KERNEL[1, 1])
        sum_w = np.where(bad, F(0.0), KERNEL[1, 1]).astype(F)
        lum = luminance(irradiance)
        m1 = np.zeros_like(lum); m2 = np.zeros_like(lum); cnt = np.zeros_like(lum)
        for ox, oy in ((-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)):
            sx, sy = fr.xs + ox * step, fr.ys + oy * step
            su, sv = fr.coords_to_uv(sx, sy)
            sdu, sdv = fr.jittered_deferred_uv(su, sv)
            inside = ~((su < 0) | (sv < 0) | (su > 1) | (sv > 1))
            cx, cy = np.clip(sx, 0, fr.RW - 1), np.clip(sy, 0, fr.RH - 1)
            s_irr = inp[cy, cx][..., :3]
            ok = inside & ~(np.isnan(s_irr).any(-1) | (s_irr > F32_MAX).any(-1))
            sn = fr.nearest_deferred(fr.normal, sdu, sdv)[..., :3]
            sn = sn / np.sqrt(sn[..., 0] * sn[..., 0] + sn[..., 1] * sn[..., 1] + sn[..., 2] * sn[..., 2])[..., None]
            s_depth = fr.nearest_deferred(fr.position, sdu, sdv)[..., 3]
            s_inst = fr.nearest_deferred(fr.instance_material, sdu, sdv)[..., 0]
            s_lum = luminance(s_irr)
            ndot = normal[..., 0] * sn[..., 0] + normal[..., 1] * sn[..., 1] + normal[..., 2] * sn[..., 2]
            w_normal = np.power(np.fmax(F(0.0), ndot), F(16.0))
            w_depth = np.exp((-np.abs(depth - s_depth)) / (np.abs(gradient[..., 0] * F(ox) + gradient[..., 1] * F(oy)) + F(0.01)))
            w_inst = np.fmax(F(0.0), F(1.0) - np.abs(instance - s_inst))
            w_lum = np.exp((-np.abs(lum - s_lum)) / (F(4.0) * np.power(variance, F(0.25)) + F(0.001)))
            w = np.fmin(np.fmax(w_normal * w_depth * w_inst * w_lum, F(0.0)), F(1.0)) * KERNEL[oy + 1, ox + 1]
            sum_irr = np.where(ok[..., None], sum_irr + s_irr * w[..., None], sum_irr)
            sum_w = np.where(ok, sum_w + w, sum_w)
            m1 = np.where(ok, m1 + s_lum, m1); m2 = np.where(ok, m2 + s_lum * s_lum, m2); cnt = np.where(ok, cnt + F(1.0), cnt)
        out = np.where((sum_w < F(0.0001))[..., None], F(0.0), sum_irr / sum_w[..., None])
        if firefly:
            mean = m1 / cnt
            var_ff = m2 / cnt - mean * mean
            clamp = lum > mean + F(3.0) * np.sqrt(var_ff)
            out = np.where(clamp[..., None], (mean / lum)[..., None] * out, out)
        color = np.concatenate([out, np.ones_like(out[..., :1])], axis=-1)
        if level == 3:
            color = color * fr.nearest_deferred(fr.albedo, du, dv)
        color = np.where((depth < F32_EPSILON)[..., None], F(0.0), color)
    return f16(color)


def tone_mapping(denoised, clear_color):                # tone_mapping.wgsl:21-32
    color = denoised[0] + denoised[1]
    color = color + denoised[2]
    rgb = np.fmax(color[..., :3], F(0.0039))
    l_old = luminance(rgb)
    with np.errstate(all="ignore"):
        rgb = rgb * ((l_old / (F(1.0) + l_old)) / l_old)[..., None]      # bevy_core_pipeline reinhard_luminance
    out = np.concatenate([rgb, color[..., 3:]], axis=-1)
    return f16(np.where((color[..., 3] > 0)[..., None], out, np.asarray(clear_color, F)))


def ulps16(a, b):
    ia = a.astype(np.float16).view(np.int16).astype(np.int32)
    ib = b.astype(np.float16).view(np.int16).astype(np.int32)
    ia = np.where(ia < 0, -32768 - ia, ia); ib = np.where(ib < 0, -32768 - ib, ib)    # sign-magnitude -> ordered
    return np.abs(ia - ib)


@pytest.mark.parametrize("scene,config,size,ratio,frames", [("cornell", "cornell_1080p", (96, 64), 1.0, 4),
                                                            ("cornell", "cornell_1080p", (90, 60), 1.5, 3),
                                                            ("city", "city_4k", (128, 72), 1.0, 3)])
def test_oracle_post_process_equals_independent_numpy_restatement(scene, config, size, ratio, frames):
    b = Bench(scene, size[0], size[1], config=config, upscale_ratio=ratio, clear_color=(0.1, 0.2, 0.3, 1.0))
    orc = b.oracle()
    for f in range(1, frames + 1):
        orc.render_frame(b.moving_inputs(f))
    fr = Frame(orc, frames, ratio)
    assert (fr.RW, fr.RH) == (int(np.ceil(F(1.0) / F(ratio) * F(size[0]))), int(np.ceil(F(1.0) / F(ratio) * F(size[1]))))
    denoised = []
    for signal in range(3):
        img, variance = demodulation(fr, signal)
        for level in range(4):
            img = denoise_level(fr, img, variance, level, firefly=signal != 0)     # denoise_direct has no FIREFLY_FILTERING
        denoised.append(img)
        got = orc.readback(L.OUT_DENOISED_DIRECT + signal).astype(F)
        d = ulps16(got, img)
        assert (d == 0).mean() >= 0.997, (signal, float((d == 0).mean()))
        assert (d > 1).sum() <= 4, (signal, int((d > 1).sum()), int(d.max()))
        assert float(got[..., :3].max()) > 0.01                                    # a real signal went through
    tm = tone_mapping([orc.readback(L.OUT_DENOISED_DIRECT + s).astype(F) for s in range(3)], (0.1, 0.2, 0.3, 1.0))
    got = orc.readback(L.OUT_TONE_MAPPED).astype(F)
    d = ulps16(got, tm)
    assert (d == 0).mean() >= 0.999 and d.max() <= 1, (float((d == 0).mean()), int(d.max()))   # divisions only: exact but for double rounding
    background = fr.position[..., 3] < F32_EPSILON if ratio == 1.0 else None
    if background is not None and background.any():
        assert np.allclose(got[background][:, :3], [0.1, 0.2, 0.3], atol=1e-3)     # clear colour where nothing was hit
