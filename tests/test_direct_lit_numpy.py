"""The sun pass of `direct_lit` (SURVEY.md 8(a) row P2 with F6's directional branch, F7 and the first step of F8), pinned
from the outside: a SECOND, independent restatement of src/shaders/light.wgsl:1044-1261 for a pixel with no history — blue-noise
look-up and golden-ratio sequence (:1075-1079), cone sample about the sun through `normal_basis` (:552-559, :611-616,
utils.wgsl), biased shadow ray, `input_radiance`, the reservoir update from empty, `r.w`, the variance estimate and the
Burley + GGX shading of bevy_pbr 0.9 — as whole-image numpy arithmetic written from the WGSL, with
occlusion by brute force over every world triangle in float64 instead of any BVH.  It is fed with the G-buffer the oracle
renders and must reproduce the oracle's `render[0]` plane.  Differences left: polynomial sin/cos and fused dot products in
hk_math.h against libm / unfused numpy (absorbed by the Rgba16Float store but for 1 ulp here and there) and shadow-ray
verdicts on pixels whose ray grazes an edge (counted, <= 0.5 % of the covered pixels).  Measured: 99.94 - 100 % of the
remaining texels bit-identical (minimal.rs, simple.rs; frames 1, 2, 4, 5).  CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_oracle import world_triangles

F = np.float32
TAU = F(6.283185307)
INV_PI = F(1.0) / F(3.141592653589793)
RAY_BIAS = F(0.02)
DISTANCE_MAX = F(65535.0)
GOLDEN_RATIO = F(1.618033989)


def dot(a, b): return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]
def normalize(a): return a / np.sqrt(dot(a, a))[..., None]
def saturate(x): return np.fmin(np.fmax(x, F(0.0)), F(1.0))
def luminance(v): return v[..., 0] * F(0.2126) + v[..., 1] * F(0.7152) + v[..., 2] * F(0.0722)
def fract(x): return x - np.floor(x)


def normal_basis_apply(n, v):                           # utils.wgsl normal_basis(n) * v  (columns t, b, n)
    s = np.fmin(np.sign(n[2]) * F(2.0) + F(1.0), F(1.0))
    u = F(-1.0) / (s + n[2])
    w = n[0] * n[1] * u
    t = np.array([F(1.0) + s * n[0] * n[0] * u, s * w, -s * n[0]], F)
    b = np.array([w, s + n[1] * n[1] * u, -n[1]], F)
    return t * v[..., 0:1] + b * v[..., 1:2] + n * v[..., 2:3]


def occluded_brute_force(tris, origin, direction):
    """any triangle hit at distance > eps along the ray, float64 Moeller-Trumbore in world space, all rays x all triangles"""
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    ab, ac = b - a, c - a
    out = np.zeros(len(origin), bool)
    margin = np.full(len(origin), np.inf)               # how far the nearest miss / hit decision is from flipping
    for k0 in range(0, len(origin), 256):
        o = origin[k0:k0 + 256].astype(np.float64)[:, None, :]
        d = direction[k0:k0 + 256].astype(np.float64)[:, None, :]
        p = np.cross(d, ac[None])
        det = (ab[None] * p).sum(-1)
        ok = np.abs(det) > 1e-12
        inv = 1.0 / np.where(ok, det, 1.0)
        ao = o - a[None]
        u = (ao * p).sum(-1) * inv
        q = np.cross(ao, ab[None])
        v = (q * d).sum(-1) * inv
        t = (q * ac[None]).sum(-1) * inv
        inside = np.minimum(np.minimum(u, v), 1.0 - u - v)          # > 0 inside the triangle
        hit = ok & (inside >= 0) & (t > 1e-7)
        out[k0:k0 + 256] = hit.any(1)
        near_edge = np.where(ok & (t > 1e-7), np.abs(inside), np.inf).min(1)
        margin[k0:k0 + 256] = near_edge
    return out, margin


def F_Schlick(f0, f90, VoH): return f0 + (f90 - f0) * np.power(F(1.0) - VoH, F(5.0))


def shade_lit(V, N, Lv, mat, radiance):                 # light.wgsl shading() -> lit(), input alpha = 1
    base = mat["base_color"][..., :3]
    metallic, reflectance = mat["metallic"][..., None], mat["reflectance"][..., None]
    rough = np.clip(mat["perceptual_roughness"], F(0.089), F(1.0)); rough = rough * rough      # perceptualRoughnessToRoughness
    F0 = F(0.16) * reflectance * reflectance * (F(1.0) - metallic) + base * metallic
    diffuse_color = base * (F(1.0) - metallic)
    H = normalize(Lv + V)
    NoL, NoH, LoH = saturate(dot(N, Lv)), saturate(dot(N, H)), saturate(dot(Lv, H))
    NoV = np.fmax(dot(N, V), F(0.0001))
    f90 = F(0.5) + F(2.0) * rough * LoH * LoH
    diffuse = diffuse_color * (F_Schlick(F(1.0), f90, NoL) * F_Schlick(F(1.0), f90, NoV) * INV_PI)[..., None]
    a = NoH * rough
    k = rough / (F(1.0) - NoH * NoH + a * a)
    D = k * k * INV_PI
    a2 = rough * rough
    lam_v = NoL * np.sqrt((NoV - a2 * NoV) * NoV + a2)
    lam_l = NoV * np.sqrt((NoL - a2 * NoL) * NoL + a2)
    Vis = F(0.5) / (lam_v + lam_l)
    f90s = saturate(F0[..., 0] * F(16.5) + F0[..., 1] * F(16.5) + F0[..., 2] * F(16.5))
    Fr = F0 + (f90s[..., None] - F0) * np.power(F(1.0) - LoH, F(5.0))[..., None]
    specular = (D * Vis)[..., None] * Fr
    return (specular + diffuse) * radiance * NoL[..., None]


def gbuffer_at_render_pixels(b, orc, frame_number):
    """(position+depth, snorm-decoded normal, instance/material, velocity/uv) as the light passes see them: one entry per RENDER
    pixel.  Rendering below the output resolution (light.rs:622-624) runs the passes over ceil(size / ratio) pixels that read the
    full-size G-buffer at jittered_deferred_coords(uv) = i32((uv -+ 0.25 texel * (ratio - 1)) * size)  (light.wgsl:1007-1017);
    at ratio 1 that is the identity."""
    pos = orc.readback(L.OUT_GBUFFER_POSITION)
    normal = np.maximum(orc.readback(L.OUT_GBUFFER_NORMAL).astype(F) / F(127.0), F(-1.0))[..., :3]
    im = orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)
    vu = orc.readback(L.OUT_GBUFFER_VELOCITY_UV)
    ratio = F(b.settings.upscale_ratio)
    if ratio != 1.0:
        DH, DW = pos.shape[:2]
        H, W = int(np.ceil(F(1.0) / ratio * F(DH))), int(np.ceil(F(1.0) / ratio * F(DW)))
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        sel = F(-0.25) if (frame_number & 1) == 0 else F(0.25)
        du = (xs.astype(F) + F(0.5)) / F(W) + sel * (F(1.0) / F(DW)) * (ratio - F(1.0))
        dv = (ys.astype(F) + F(0.5)) / F(H) + sel * (F(1.0) / F(DH)) * (ratio - F(1.0))
        dx, dy = np.trunc(du * F(DW)).astype(np.int64), np.trunc(dv * F(DH)).astype(np.int64)
        pos, normal, im, vu = pos[dy, dx], normal[dy, dx], im[dy, dx], vu[dy, dx]
    return pos, normal, im, vu


def direct_sun_numpy(b, orc, frame_number, noise):
    pos, normal, im, _ = gbuffer_at_render_pixels(b, orc, frame_number)
    H, W = pos.shape[:2]
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    position, depth = pos[..., :3], pos[..., 3]
    # s.random (:1075-1079): texture frame % 16, nearest + repeat, then the golden-ratio shift
    tex = noise.reshape(16, 64, 64, 4)[frame_number % 16].astype(F) / F(255.0)
    nu = (xs.astype(F) + F(frame_number) + F(0.5)) / F(64.0)
    nv = (ys.astype(F) + F(frame_number) + F(0.5)) / F(64.0)
    random = tex[np.floor(nv * F(64.0)).astype(np.int64) % 64, np.floor(nu * F(64.0)).astype(np.int64) % 64]
    random = fract(random + F(frame_number) * GOLDEN_RATIO)
    # select_light_candidate, directional only (:611-616)
    sun = np.array(list(b.lights.direction_to_light), F)
    cos_angle = np.cos(F(b.settings.solar_angle)).astype(F)
    z = F(1.0) - (F(1.0) - cos_angle) * random[..., 2]
    theta = TAU * random[..., 3]
    r = np.sqrt(F(1.0) - z * z)
    cone_dir = np.stack([r * np.cos(theta), r * np.sin(theta), z], -1).astype(F)
    direction = normal_basis_apply(sun, cone_dir).astype(F)
    covered = depth >= F(1.1920929e-7)
    trace = covered & (dot(direction, normal) > 0)
    origin = position + normal * RAY_BIAS
    tris, _ = world_triangles(b.world.buffers())
    idx = np.nonzero(trace.ravel())[0]
    occ, margin = occluded_brute_force(tris, origin.reshape(-1, 3)[idx], direction.reshape(-1, 3)[idx])
    occluded = np.zeros(H * W, bool); occluded[idx] = occ
    grazing = np.zeros(H * W, bool); grazing[idx] = margin < 2e-3
    occluded, grazing = occluded.reshape(H, W), grazing.reshape(H, W)
    # input_radiance (:842-872): an unoccluded ray inside the cone sees the sun colour
    in_cone = dot(direction, sun) >= cos_angle
    lit = trace & ~occluded & in_cone
    sun_color = np.array(list(b.lights.directional_color), F)[:3]
    radiance = np.where(lit[..., None], sun_color, F(0.0))
    w_new = luminance(radiance)                                    # candidate.p = 1
    # update_reservoir on an empty reservoir: the sample is taken iff w_new / w_sum = 1 > rand, i.e. iff w_new > 0
    taken = w_new > 0
    with np.errstate(all="ignore"):
        r_w = np.where(taken, w_new / (F(1.0) * luminance(radiance)), F(0.0))     # w_sum / (count * luminance)
        variance = np.fmin(w_new * w_new / F(1.0) - np.power(w_new / F(1.0), F(2.0)), F(10.0))
        # shading with L = normalize(sample_position - visible_position); sample = position + direction * DISTANCE_MAX (:488-494)
        Lv = normalize((position + direction * DISTANCE_MAX) - position)
        V = normalize(np.array(list(b.view.world_position), F) - position)
        mats = b.world.buffers()["materials"][np.floor(im[..., 1]).astype(np.int64)]
        color = shade_lit(V, normal, Lv, mats, radiance) * r_w[..., None]
    color = np.where(taken[..., None], color, F(0.0))
    # the sun pipeline is specialised with RENDER_EMISSIVE (light.rs:409-412): + compute_emissive_radiance(surface.emissive) (:594-596, :1240-1242)
    color = color + F(255.0) * mats["emissive"][..., 3:4] * mats["emissive"][..., :3]
    color = np.where(covered[..., None], color, F(0.0))
    # only the NO_TEXTURE form of retreive_surface (:730-742) is restated here: pixels of textured materials are left out
    textured = (mats["base_color_texture"] != 0xFFFFFFFF) | (mats["emissive_texture"] != 0xFFFFFFFF) | \
               (mats["metallic_roughness_texture"] != 0xFFFFFFFF) | (mats["occlusion_texture"] != 0xFFFFFFFF)
    return color.astype(F), np.where(covered, variance, F(0.0)), grazing | (textured & covered), covered


def ulps16(a, b):
    ia = a.astype(np.float16).view(np.int16).astype(np.int32)
    ib = b.astype(np.float16).view(np.int16).astype(np.int32)
    ia = np.where(ia < 0, -32768 - ia, ia); ib = np.where(ib < 0, -32768 - ib, ib)
    return np.abs(ia - ib)


@pytest.mark.parametrize("scene,size,frames,ratio", [("minimal", (96, 64), (1, 2, 4), 1.0), ("simple", (112, 64), (1, 5), 1.0),
                                                     ("minimal", (120, 80), (1, 2), 1.5), ("minimal", (120, 80), (2,), 2.0)])
def test_oracle_direct_sun_equals_independent_numpy_restatement(scene, size, frames, ratio):
    # temporal_reuse = 0: the reservoir is never stored (:1226-1228), so every frame starts from an empty history
    b = Bench(scene, size[0], size[1], taa=plugin.TAA_NONE, upscale_ratio=ratio, temporal_reuse=0, denoise=0, indirect_bounces=1)
    orc = b.oracle()
    noise = plugin.load_noise()
    for f in range(1, max(frames) + 1):
        inp = b.inputs(f)
        assert f % inp.frame.direct_validate_interval != 0 or f not in frames     # non-validation frames only
        orc.render_frame(inp)
        if f not in frames:
            continue
        want, variance, grazing, covered = direct_sun_numpy(b, orc, f, noise)
        got = orc.readback(L.OUT_RENDER_DIRECT).astype(F)
        d = ulps16(got[..., :3], want).max(-1)
        clean = covered & ~grazing
        assert covered.mean() > 0.3 and (got[..., :3].sum(-1) > 0.05).mean() > 0.15          # sun-lit pixels exist
        assert (got[..., :3].sum(-1)[covered] == 0).mean() > 0.02                              # and shadowed / back-facing ones
        assert (d[clean] <= 1).mean() >= 0.995, (f, float((d[clean] <= 1).mean()))
        assert (d[clean] > 2).sum() <= max(2, clean.sum() // 1000), (f, int((d[clean] > 2).sum()))
        assert (d[clean] == 0).mean() >= 0.995, (f, float((d[clean] == 0).mean()))
        assert grazing.sum() <= 0.05 * covered.sum()          # grazing rays + (simple.rs) the two earth-textured spheres
        assert (got[..., 3][covered] == 1).all() and not got[~covered].any()
        assert np.array_equal(orc.readback(L.OUT_VARIANCE_DIRECT), variance)                  # 0 everywhere: one sample, no spread
