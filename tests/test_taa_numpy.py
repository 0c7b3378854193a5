"""`taa_jasmine` (the last pass of `HikariSettings::default()` before presentation), pinned from the outside:
a SECOND, independent restatement of src/shaders/taa.wgsl:57-170 in whole-image numpy float32 arithmetic written from the WGSL —
the closest-depth velocity pick, reprojection, the five textureGather footprints that decide depth / position / content misses,
the 5-tap Catmull-Rom history fetch, YCoCg variance clipping on a miss, the 0.1 / ratio blend — fed with the images the oracle
holds (tone-mapped current frame, its own previous output, both G-buffer generations) under a translating camera, compared with
the oracle's new `taa_output`.  Only +, *, /, sqrt are involved, so the two agree bit for bit except where a float comparison
sits on the fence.  Measured: 99.6 - 100 % of the texels bit-identical, never more than 1 f16 ulp apart.  CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_direct_lit_numpy import F, ulps16


def RGB_to_YCoCg(c):
    return np.stack([c[..., 0] / F(4) + c[..., 1] / F(2) + c[..., 2] / F(4), c[..., 0] / F(2) - c[..., 2] / F(2),
                     -c[..., 0] / F(4) + c[..., 1] / F(2) - c[..., 2] / F(4)], -1)


def clamp01(a):
    # clamp = min(max(x, 0), 1) with IEEE minNum / maxNum (the rule include/hk_math.h fixes where WGSL leaves NaN open): NaN -> 0.
    # It matters: the variance under the square root can round below zero in one channel, and the clipped colour is then NaN there.
    return np.fmin(np.fmax(a, F(0)), F(1))


def YCoCg_to_RGB(c):
    return clamp01(np.stack([c[..., 0] + c[..., 1] - c[..., 2], c[..., 0] + c[..., 2], c[..., 0] - c[..., 1] - c[..., 2]], -1))


class Tex:
    def __init__(self, a):
        self.a = a; self.h, self.w = a.shape[:2]

    def texel(self, x, y):
        return self.a[np.clip(y, 0, self.h - 1), np.clip(x, 0, self.w - 1)]

    def nearest(self, u, v):
        return self.texel(np.floor(u * F(self.w)).astype(np.int64), np.floor(v * F(self.h)).astype(np.int64))

    def linear(self, u, v):
        fx, fy = u * F(self.w) - F(0.5), v * F(self.h) - F(0.5)
        x0, y0 = np.floor(fx), np.floor(fy)
        ax, ay = (fx - x0)[..., None], (fy - y0)[..., None]
        x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
        top = self.texel(x0, y0) * (F(1) - ax) + self.texel(x0 + 1, y0) * ax
        bot = self.texel(x0, y0 + 1) * (F(1) - ax) + self.texel(x0 + 1, y0 + 1) * ax
        return top * (F(1) - ay) + bot * ay

    def gather_w(self, u, v):
        fx, fy = u * F(self.w) - F(0.5), v * F(self.h) - F(0.5)
        i, j = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
        return np.stack([self.texel(i, j + 1)[..., 3], self.texel(i + 1, j + 1)[..., 3], self.texel(i + 1, j)[..., 3], self.texel(i, j)[..., 3]], -1)


def taa_numpy(render, previous_render, position, previous_position, velocity_uv, previous_velocity_uv, ratio, clear_color):
    H, W = render.a.shape[:2]
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    size = np.array([W, H], F)
    tx, ty = F(1) / F(W), F(1) / F(H)
    u, v = (xs.astype(F) + F(0.5)) / F(W), (ys.astype(F) + F(0.5)) / F(H)
    original = render.nearest(u, v)
    current = original[..., :3]
    # nearest_velocity (:57-77)
    d = np.stack([position.nearest(u + tx, v + ty)[..., 3], position.nearest(u - tx, v + ty)[..., 3],
                  position.nearest(u + tx, v - ty)[..., 3], position.nearest(u - tx, v - ty)[..., 3]], -1)
    dmax = d.max(-1)
    depth = position.nearest(u, v)[..., 3]
    eq = d == dmax[..., None]
    ox = (np.where(eq, np.array([1, -1, 1, -1], F), F(0)) * tx).sum(-1, dtype=F)
    oy = (np.where(eq, np.array([1, 1, -1, -1], F), F(0)) * ty).sum(-1, dtype=F)
    closer = depth < dmax
    vel = velocity_uv.nearest(u + np.where(closer, ox, F(0)), v + np.where(closer, oy, F(0)))[..., :2]
    pu, pv = u - vel[..., 0], v - vel[..., 1]
    boundary_miss = (np.abs(pu - F(0.5)) > F(0.5)) | (np.abs(pv - F(0.5)) > F(0.5))
    cpd = position.nearest(u, v)
    has_content = cpd[..., 3] > 0
    depth_miss = cpd[..., 3] == 0
    position_miss = cpd[..., 3] == 0
    for bx, by in ((0.0, 0.0), (1.5, 1.5), (-1.5, 1.5), (1.5, -1.5), (-1.5, -1.5)):
        su, sv = pu + F(bx) * tx, pv + F(by) * ty
        pd = previous_position.gather_w(su, sv)
        with np.errstate(all="ignore"):
            ratio_d = np.where(pd == 0, F(1), cpd[..., 3:4] / pd)
        has_content = has_content | (pd > 0).any(-1)
        depth_miss = depth_miss | (ratio_d < F(0.95)).any(-1)
        pp = previous_position.nearest(su, sv)[..., :3]
        dd = cpd[..., :3] - pp
        position_miss = position_miss | (np.sqrt(dd[..., 0] * dd[..., 0] + dd[..., 1] * dd[..., 1] + dd[..., 2] * dd[..., 2]) > F(0.5))
    pvel = previous_velocity_uv.nearest(pu, pv)[..., :2]
    dv = vel - pvel
    velocity_miss = np.sqrt(dv[..., 0] * dv[..., 0] + dv[..., 1] * dv[..., 1]) > F(0.00005)
    # Catmull-Rom history (:121-139)
    sp = np.stack([pu, pv], -1) * size
    tp1 = np.floor(sp - F(0.5)) + F(0.5)
    f = sp - tp1
    w0 = f * (F(-0.5) + f * (F(1.0) - F(0.5) * f))
    w1 = F(1.0) + f * f * (F(-2.5) + F(1.5) * f)
    w2 = f * (F(0.5) + f * (F(2.0) - F(1.5) * f))
    w3 = f * f * (F(-0.5) + F(0.5) * f)
    w12 = w1 + w2
    off12 = w2 / (w1 + w2)
    ts = np.array([tx, ty], F)
    p0, p3, p12 = (tp1 - F(1.0)) * ts, (tp1 + F(2.0)) * ts, (tp1 + off12) * ts
    fetch = lambda a, b_: clamp01(previous_render.linear(a, b_)[..., :3])
    prev = np.zeros(current.shape, F)
    prev = prev + fetch(p12[..., 0], p0[..., 1]) * w12[..., 0:1] * w0[..., 1:2]
    prev = prev + fetch(p0[..., 0], p12[..., 1]) * w0[..., 0:1] * w12[..., 1:2]
    prev = prev + fetch(p12[..., 0], p12[..., 1]) * w12[..., 0:1] * w12[..., 1:2]
    prev = prev + fetch(p3[..., 0], p12[..., 1]) * w3[..., 0:1] * w12[..., 1:2]
    prev = prev + fetch(p12[..., 0], p3[..., 1]) * w12[..., 0:1] * w3[..., 1:2]
    clip_it = boundary_miss | (position_miss & velocity_miss & depth_miss)
    srt = lambda a, b_: RGB_to_YCoCg(clamp01(render.nearest(a, b_)[..., :3]))
    s = [srt(u - tx, v + ty), srt(u, v + ty), srt(u + tx, v + ty), srt(u - tx, v), RGB_to_YCoCg(current), srt(u + tx, v),
         srt(u - tx, v - ty), srt(u, v - ty), srt(u + tx, v - ty)]
    m1 = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7] + s[8]
    m2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6] + s[7] * s[7] + s[8] * s[8]
    mean = m1 / F(9.0)
    with np.errstate(all="ignore"):
        sigma = np.sqrt(m2 / F(9.0) - mean * mean)
        pc = RGB_to_YCoCg(prev)
        lo, hi = mean - sigma, mean + sigma
        p_clip, e_clip = F(0.5) * (hi + lo), F(0.5) * (hi - lo)
        v_clip = pc - p_clip
        a_unit = np.abs(v_clip / e_clip)
        ma = np.fmax(a_unit[..., 0], np.fmax(a_unit[..., 1], a_unit[..., 2]))
        clipped = YCoCg_to_RGB(np.where((ma > 1.0)[..., None], p_clip + v_clip / ma[..., None], pc))
    prev = np.where(clip_it[..., None], clipped, prev)
    t = F(0.1) / F(ratio)
    out = prev * (F(1.0) - t) + current * t
    out4 = np.concatenate([out, original[..., 3:4]], -1)
    out4 = np.where(has_content[..., None], out4, np.asarray(clear_color, F))
    return out4.astype(F), clip_it & has_content


@pytest.mark.parametrize("scene,config,size", [("cornell", "cornell_1080p", (96, 72)), ("minimal", None, (96, 64))])
def test_oracle_taa_equals_independent_numpy_restatement(scene, config, size):
    kw = dict(taa=plugin.TAA_JASMINE, upscale_kind=plugin.UPSCALE_FSR1, upscale_ratio=1.0, clear_color=(0.2, 0.3, 0.4, 1.0))
    b = Bench(scene, size[0], size[1], config=config, **kw) if config else Bench(scene, size[0], size[1], **kw)
    orc = b.oracle()
    clipped_total = 0
    prev_position = prev_velocity = None
    for f in range(1, 7):
        inp = b.moving_inputs(f, step=(0.06, 0.02, -0.04)) if f != 4 else b.moving_inputs(f, step=(0.0, 0.0, 0.0))
        inp.temporal_upscalers = 1
        inp.fsr1 = 0
        previous_taa = orc.readback(L.OUT_TAA).copy() if f > 1 else None
        orc.render_frame(inp)
        position, velocity = orc.readback(L.OUT_GBUFFER_POSITION), orc.readback(L.OUT_GBUFFER_VELOCITY_UV)
        if f > 2:
            want, clipped = taa_numpy(Tex(orc.readback(L.OUT_TONE_MAPPED).astype(F)), Tex(previous_taa.astype(F)), Tex(position),
                                      Tex(prev_position), Tex(velocity), Tex(prev_velocity), 1.0, (0.2, 0.3, 0.4, 1.0))
            got = orc.readback(L.OUT_TAA).astype(F)
            d = ulps16(got, want).max(-1)
            assert (d == 0).mean() >= 0.995 and (d <= 1).mean() >= 0.999, (f, float((d == 0).mean()), float((d <= 1).mean()))
            clipped_total += int(clipped.sum())
        prev_position, prev_velocity = position.copy(), velocity.copy()
    assert clipped_total > 50          # the disocclusion branch ran


def test_oracle_taa_after_smaa_tu4x_the_default_pipeline():
    """HikariSettings::default(): Upscale::SMAA_TU_2_0 + Taa::Jasmine — TAA runs on the SMAA-upscaled image (twice the render size,
    post_process.rs:1010-1014, 726-731) and blends with 0.1 / upscale_ratio"""
    size, ratio = (112, 80), 2.0
    b = Bench("cornell", size[0], size[1], config="cornell_1080p", taa=plugin.TAA_JASMINE, upscale_kind=plugin.UPSCALE_SMAA_TU4X,
              upscale_ratio=ratio, clear_color=(0.2, 0.3, 0.4, 1.0))
    orc = b.oracle()
    prev_position = prev_velocity = None
    for f in range(1, 7):
        inp = b.moving_inputs(f, step=(0.06, 0.02, -0.04))
        inp.temporal_upscalers = 1
        previous_taa = orc.readback(L.OUT_TAA).copy() if f > 1 else None
        orc.render_frame(inp)
        position, velocity = orc.readback(L.OUT_GBUFFER_POSITION), orc.readback(L.OUT_GBUFFER_VELOCITY_UV)
        if f > 2:
            upscaled = orc.readback(L.OUT_UPSCALED).astype(F)
            assert upscaled.shape[:2] == (size[1], size[0]) == orc.readback(L.OUT_TAA).shape[:2]       # 2 x ceil(size / 2)
            want, clipped = taa_numpy(Tex(upscaled), Tex(previous_taa.astype(F)), Tex(position), Tex(prev_position), Tex(velocity),
                                      Tex(prev_velocity), ratio, (0.2, 0.3, 0.4, 1.0))
            d = ulps16(orc.readback(L.OUT_TAA).astype(F), want).max(-1)
            assert (d == 0).mean() >= 0.995 and (d <= 1).mean() >= 0.999, (f, float((d == 0).mean()), float((d <= 1).mean()))
        prev_position, prev_velocity = position.copy(), velocity.copy()
