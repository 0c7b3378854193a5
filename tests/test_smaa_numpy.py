"""`smaa_tu4x` + `smaa_tu4x_extrapolate` (`Upscale::SmaaTu4x`, the default upscaler), pinned from the
outside: a SECOND, independent restatement of src/shaders/smaa.wgsl:81-271 in whole-image numpy float32 arithmetic written
from the WGSL — which output pixel of each 2 x 2 quad is the current sample and which the re-projected previous one (frame
parity), the closest-depth velocity, depth / instance / velocity misses from five gather footprints, the bias search and
2 x 2 YCoCg variance clipping, the sub-pixel-velocity remix, then the differential blend that fills the other two pixels of
the quad — fed with the oracle's images (tone-mapped current and previous frame, both G-buffer generations) under a
translating camera, compared with the oracle's `upscale_output`.  Measured (ratios 1 / 1.5 / 2, frames 3-6): 99.91 - 99.98 % of
the output texels bit-identical, 99.9 % within 1 f16 ulp; the handful beyond sit on a float comparison (`dds < min_ds` ties,
nearest look-ups landing exactly on a texel boundary at ratio 1).  CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_direct_lit_numpy import F, TAU, fract, luminance, ulps16
from tests.test_taa_numpy import RGB_to_YCoCg, Tex, YCoCg_to_RGB


def gather(tex, u, v, comp):
    fx, fy = u * F(tex.w) - F(0.5), v * F(tex.h) - F(0.5)
    i, j = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
    return np.stack([tex.texel(i, j + 1)[..., comp], tex.texel(i + 1, j + 1)[..., comp], tex.texel(i + 1, j)[..., comp], tex.texel(i, j)[..., comp]], -1)


def smaa_numpy(render, previous_render, position, previous_position, velocity_uv, previous_velocity_uv, instance_material, number):
    RH, RW = render.a.shape[:2]
    OW, OH = 2 * RW, 2 * RH
    ys, xs = np.meshgrid(np.arange(RH), np.arange(RW), indexing="ij")
    u, v = (xs.astype(F) + F(0.5)) / F(RW), (ys.astype(F) + F(0.5)) / F(RH)
    tx, ty = F(1) / F(OW), F(1) / F(OH)
    cur_j, prev_j = (0, 1) if (number & 1) == 0 else (1, 0)
    current_color = render.nearest(u, v)[..., :3]
    pu, pv = (F(2) * xs.astype(F) + F(prev_j) + F(0.5)) / F(OW), (F(2) * ys.astype(F) + F(prev_j) + F(0.5)) / F(OH)
    # nearest_velocity at previous_output_uv, texel of the position texture (:52-72)
    dx, dy = F(1) / F(position.w), F(1) / F(position.h)
    d = np.stack([position.nearest(pu + dx, pv + dy)[..., 3], position.nearest(pu - dx, pv + dy)[..., 3],
                  position.nearest(pu + dx, pv - dy)[..., 3], position.nearest(pu - dx, pv - dy)[..., 3]], -1)
    dmax = d.max(-1)
    depth = position.nearest(pu, pv)[..., 3]
    eq = d == dmax[..., None]
    ox = (np.where(eq, np.array([1, -1, 1, -1], F), F(0)) * dx).sum(-1, dtype=F)
    oy = (np.where(eq, np.array([1, 1, -1, -1], F), F(0)) * dy).sum(-1, dtype=F)
    closer = depth < dmax
    vel = velocity_uv.nearest(pu + np.where(closer, ox, F(0)), pv + np.where(closer, oy, F(0)))[..., :2]
    ru, rv = pu - vel[..., 0], pv - vel[..., 1]
    previous_color = previous_render.nearest(ru, rv)[..., :3]
    boundary_miss = (np.abs(ru - F(0.5)) > F(0.5)) | (np.abs(rv - F(0.5)) > F(0.5))
    current_instance = instance_material.nearest(pu, pv)[..., 0]
    current_depth = depth
    depth_miss = current_depth == 0
    instance_miss = np.zeros_like(depth_miss)
    biases = ((0.0, 0.0), (2.5, 2.5), (-2.5, 2.5), (2.5, -2.5), (-2.5, -2.5))
    for bx, by in biases:
        su, sv = ru + F(bx) * tx, rv + F(by) * ty
        pd = gather(previous_position, su, sv, 3)
        with np.errstate(all="ignore"):
            ratio = np.where(pd == 0, F(1), current_depth[..., None] / pd)
        low = (ratio < F(0.95)).any(-1)
        depth_miss = depth_miss | low
        previous_instance = instance_material.nearest(su, sv)[..., 0]
        instance_miss = instance_miss | (low & (np.abs(previous_instance - current_instance) > F(1.0)))
    pvel = previous_velocity_uv.nearest(ru, rv)[..., :2]
    dv = vel - pvel
    velocity_miss = np.sqrt(dv[..., 0] * dv[..., 0] + dv[..., 1] * dv[..., 1]) > F(0.0001)
    clip_it = boundary_miss | ((depth_miss | instance_miss) & velocity_miss)
    # bias search (:151-160) and 2 x 2 variance clipping (:162-176)
    bias_u = np.zeros_like(pu); bias_v = np.zeros_like(pv); min_ds = np.full(pu.shape, F(10.0))
    for bx, by in biases:
        ds = gather(position, pu + F(bx) * tx, pv + F(by) * ty, 3)
        diff = current_depth[..., None] - ds
        dds = np.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1] + diff[..., 2] * diff[..., 2] + diff[..., 3] * diff[..., 3])
        better = dds < min_ds
        bias_u = np.where(better, F(bx) * tx, bias_u); bias_v = np.where(better, F(by) * ty, bias_v)
        min_ds = np.fmin(min_ds, dds)
    cr, cg, cb = (gather(render, pu + bias_u, pv + bias_v, k) for k in range(3))
    s = [RGB_to_YCoCg(np.stack([cr[..., k], cg[..., k], cb[..., k]], -1)) for k in range(4)]
    m1 = s[0] + s[1] + s[2] + s[3]
    m2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3]
    mean = m1 / F(4.0)
    with np.errstate(all="ignore"):
        sigma = np.sqrt(m2 / F(4.0) - mean * mean)
        pc = RGB_to_YCoCg(previous_color)
        lo, hi = mean - sigma, mean + sigma
        p_clip, e_clip = F(0.5) * (hi + lo), F(0.5) * (hi - lo)
        v_clip = pc - p_clip
        a_unit = np.abs(v_clip / e_clip)
        ma = np.fmax(a_unit[..., 0], np.fmax(a_unit[..., 1], a_unit[..., 2]))
        clipped = YCoCg_to_RGB(np.where((ma > 1.0)[..., None], p_clip + v_clip / ma[..., None], pc))
    previous_color = np.where(clip_it[..., None], clipped, previous_color)
    # sub-pixel remix (:178-186)
    sub = fract(vel / (F(2.0) * np.array([tx, ty], F)))
    blend = np.fmax(sub[..., 0], sub[..., 1])
    blend = np.clip(-np.cos(blend * TAU), F(0), F(1))[..., None]
    remix = render.linear(pu, pv)[..., :3]
    previous_color = previous_color * (F(1.0) - blend) + remix * blend
    out = np.zeros((OH, OW, 4), F)
    out[2 * ys + cur_j, 2 * xs + cur_j] = np.concatenate([current_color, np.ones_like(current_color[..., :1])], -1)
    out[2 * ys + prev_j, 2 * xs + prev_j] = np.concatenate([previous_color, np.ones_like(current_color[..., :1])], -1)
    out = out.astype(np.float16).astype(F)                       # the Rgba16Float store the second pass reads back
    # smaa_tu4x_extrapolate (:200-271)
    pad = np.zeros((OH + 4, OW + 4, 4), F); pad[2:-2, 2:-2] = out
    at = lambda ddx, ddy: pad[2 * ys + 2 + ddy, 2 * xs + 2 + ddx]                  # textureLoad, zero outside
    t, b_, n, e, s_, w = at(0, 0), at(1, 1), at(1, -1), at(2, 0), at(0, 2), at(-1, 1)
    lum3 = lambda p, q: luminance(np.abs(p[..., :3] - q[..., :3]))
    dh = (lum3(w, b_), lum3(t, e)); dvv = (lum3(t, s_), lum3(n, b_))
    fx_ = np.fmax(dvv[0], F(0.001)) * np.fmax(dvv[1], F(0.001))
    fy_ = np.fmax(dh[0], F(0.001)) * np.fmax(dh[1], F(0.001))
    fz = F(1.0) / (fx_ + fy_)
    def blend4(tt, bb, ll, rr):
        c = np.zeros_like(tt)
        c = c + (ll + rr) * fx_[..., None]
        c = c + (tt + bb) * fy_[..., None]
        return (F(0.5) * fz)[..., None] * c
    out[2 * ys + 1, 2 * xs] = blend4(t, s_, w, b_)
    out[2 * ys, 2 * xs + 1] = blend4(n, b_, t, e)
    return out, clip_it


@pytest.mark.parametrize("scene,config,size,ratio", [("cornell", "cornell_1080p", (80, 60), 1.0), ("cornell", "cornell_1080p", (112, 80), 2.0),
                                                      ("minimal", None, (96, 64), 1.5)])
def test_oracle_smaa_tu4x_equals_independent_numpy_restatement(scene, config, size, ratio):
    kw = dict(taa=plugin.TAA_NONE, upscale_kind=plugin.UPSCALE_SMAA_TU4X, upscale_ratio=ratio)
    b = Bench(scene, size[0], size[1], config=config, **kw) if config else Bench(scene, size[0], size[1], **kw)
    orc = b.oracle()
    prev = None
    clipped_total = 0
    for f in range(1, 7):
        inp = b.moving_inputs(f, step=(0.06, 0.02, -0.04)) if f != 4 else b.moving_inputs(f, step=(0.0, 0.0, 0.0))
        inp.temporal_upscalers = 1
        orc.render_frame(inp)
        cur = dict(tone=orc.readback(L.OUT_TONE_MAPPED).astype(F), position=orc.readback(L.OUT_GBUFFER_POSITION),
                   velocity=orc.readback(L.OUT_GBUFFER_VELOCITY_UV), im=orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL))
        if f > 2:
            want, clipped = smaa_numpy(Tex(cur["tone"]), Tex(prev["tone"]), Tex(cur["position"]), Tex(prev["position"]), Tex(cur["velocity"]),
                                       Tex(prev["velocity"]), Tex(cur["im"]), f)
            got = orc.readback(L.OUT_UPSCALED).astype(F)
            assert got.shape == want.shape
            d = ulps16(got, want).max(-1)
            assert (d == 0).mean() >= 0.99 and (d <= 1).mean() >= 0.998, (f, float((d == 0).mean()), float((d <= 1).mean()))
            clipped_total += int(clipped.sum())
        prev = cur
    assert clipped_total > 50
