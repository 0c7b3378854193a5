"""Temporal ReSTIR with history, pinned from the outside on the emissive `direct_lit` pass —
the pass that lights cornell.  A SECOND, independent restatement, in numpy from the WGSL, of what happens to a pixel that
HAS a reservoir from the previous frame:
  unpack_reservoir (:77-106: f16 pairs, unorm16 random numbers, snorm8 normals with lifetime / sample_position.w in the
  fourth byte, instance id as a float) -> check_previous_reservoir (:917-935) -> the new candidate of this frame ->
  update_reservoir with rand = fract(sum(random)) and the keep-or-replace rule (:146-171) -> the M clamp of temporal_restir
  (:937-952) -> r.w, lifetime + 1, the variance estimate (:1216-1224) -> pack_reservoir (:108-136) and the shaded output.
The previous frame's reservoir buffer is read back from the oracle and handed to the restatement; its result is compared
with the oracle's NEW reservoir buffer field by field (bits of the packed record) and with render[1].  Static camera, so a
pixel's history is its own record (previous_uv = uv).  Pixels whose candidate ray grazes an edge are left out (counted).
Measured (cornell, frames 2-4, M clamp 3): every packed field but sample_position bit-identical on >= 99.5 % of the records
(sample_position: kept samples bit-identical, replaced ones within 5e-6), count exact, render[1] 99.9 % bit-identical.
It also confirmed a consequence of the WGSL that is easy to miss: `visible_instance` travels with the SAMPLE, so a record whose
sample was never replaced stores instance 0 and is rejected by the instance test one frame later (a third of cornell's
pixels at frame 2) — the oracle does the same.  CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_direct_lit_numpy import F, GOLDEN_RATIO, RAY_BIAS, dot, fract, luminance, normalize, ulps16
from tests.test_indirect_numpy import DONT_SAMPLE, Scene, shading

U = np.uint32


def unpack_f16x2(w):
    w = np.asarray(w, U)
    return (w & U(0xFFFF)).astype(np.uint16).view(np.float16).astype(F), (w >> U(16)).astype(np.uint16).view(np.float16).astype(F)


def pack_f16x2(a, b):
    with np.errstate(over="ignore"):                                       # w2_sum beyond 65504 stores as +inf, as pack2x16float does
        return _pack_f16x2(a, b)


def _pack_f16x2(a, b):
    return np.asarray(a, F).astype(np.float16).view(np.uint16).astype(U) | (np.asarray(b, F).astype(np.float16).view(np.uint16).astype(U) << U(16))


def unpack_unorm16x2(w):
    w = np.asarray(w, U)
    return (w & U(0xFFFF)).astype(F) / F(65535.0), (w >> U(16)).astype(F) / F(65535.0)


def pack_unorm16x2(a, b):
    q = lambda x: np.floor(F(0.5) + F(65535.0) * np.clip(x, F(0.0), F(1.0))).astype(U)
    return q(a) | (q(b) << U(16))


def unpack_snorm8x4(w):
    w = np.asarray(w, U)
    b = np.stack([(w >> U(8 * k)) & U(0xFF) for k in range(4)], -1).astype(np.uint8).view(np.int8).astype(F)
    return np.maximum(b / F(127.0), F(-1.0))


def pack_snorm8x4(v):
    # clamp as min(max(v, -1), 1) with IEEE minNum / maxNum: the normalised zero normal of a never-replaced sample is NaN and
    # packs as -127 under the rule hk_math.h fixes for both implementations (WGSL leaves it open)
    q = np.floor(F(0.5) + F(127.0) * np.fmin(np.fmax(v, F(-1.0)), F(1.0))).astype(np.int32) & 0xFF
    return (q[..., 0] | (q[..., 1] << 8) | (q[..., 2] << 16) | (q[..., 3] << 24)).astype(U)


def unpack_reservoir(p):
    r = {}
    r["count"], r["w"] = unpack_f16x2(p["reservoir"][..., 0])
    r["w_sum"], r["w2_sum"] = unpack_f16x2(p["reservoir"][..., 1])
    a, b_ = unpack_f16x2(p["radiance"][..., 0]); c, d = unpack_f16x2(p["radiance"][..., 1])
    r["radiance"] = np.stack([a, b_, c, d], -1)
    a, b_ = unpack_unorm16x2(p["random"][..., 0]); c, d = unpack_unorm16x2(p["random"][..., 1])
    r["random"] = np.stack([a, b_, c, d], -1)
    t2 = unpack_snorm8x4(p["visible_normal"])
    r["visible_position"] = p["visible_position"].astype(F)
    with np.errstate(all="ignore"):
        r["visible_normal"] = normalize(t2[..., :3])
    r["lifetime"] = F(127.0) * (F(1.0) + t2[..., 3])
    t2 = unpack_snorm8x4(p["sample_normal"])
    r["sample_position"] = np.concatenate([p["sample_position"][..., :3], t2[..., 3:4]], -1).astype(F)
    with np.errstate(all="ignore"):
        r["sample_normal"] = normalize(t2[..., :3])
    r["visible_instance"] = p["sample_position"][..., 3].astype(np.int64)
    return r


def pack_records(r, visible_position, visible_normal):
    """pack_reservoir (:108-136) of a reservoir dict with the given visible point"""
    n = len(r["count"])
    packed = np.zeros(n, L.PACKED_RESERVOIR)
    packed["reservoir"][:, 0] = pack_f16x2(r["count"], r["w"]); packed["reservoir"][:, 1] = pack_f16x2(r["w_sum"], r["w2_sum"])
    packed["radiance"][:, 0] = pack_f16x2(r["radiance"][:, 0], r["radiance"][:, 1])
    packed["radiance"][:, 1] = pack_f16x2(r["radiance"][:, 2], r["radiance"][:, 3])
    packed["random"][:, 0] = pack_unorm16x2(r["random"][:, 0], r["random"][:, 1])
    packed["random"][:, 1] = pack_unorm16x2(r["random"][:, 2], r["random"][:, 3])
    packed["visible_position"] = visible_position
    packed["sample_position"] = np.concatenate([r["sample_position"][:, :3], r["visible_instance"].astype(F)[:, None]], 1)
    packed["visible_normal"] = pack_snorm8x4(np.concatenate([visible_normal, (r["lifetime"] / F(127.0) - F(1.0))[:, None]], 1))
    packed["sample_normal"] = pack_snorm8x4(np.concatenate([r["sample_normal"], r["sample_position"][:, 3:4]], 1))
    return packed


def temporal_emissive_numpy(b, orc, frame_number, noise, previous, inp=None, emissive=True):
    """`inp` = the frame inputs when the camera moves: history is then fetched at previous_uv = uv - velocity (:1089-1090).
    emissive = False restates the sun pipeline of the same entry point (sample_directional, DONT_SAMPLE_EMISSIVE, RENDER_EMISSIVE)."""
    sc = Scene(b)
    pos = orc.readback(L.OUT_GBUFFER_POSITION)
    g_normal = np.maximum(orc.readback(L.OUT_GBUFFER_NORMAL).astype(F) / F(127.0), F(-1.0))[..., :3]
    im = orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)
    H, W = pos.shape[:2]
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    tex = noise.reshape(16, 64, 64, 4)[frame_number % 16].astype(F) / F(255.0)
    nu = (xs.astype(F) + F(frame_number) + F(0.5)) / F(64.0)
    nv = (ys.astype(F) + F(frame_number) + F(0.5)) / F(64.0)
    random = tex[np.floor(nv * F(64.0)).astype(np.int64) % 64, np.floor(nu * F(64.0)).astype(np.int64) % 64]
    random = fract(random + F(frame_number) * GOLDEN_RATIO).reshape(-1, 4)
    covered = (pos[..., 3] >= F(1.1920929e-7)).reshape(-1)
    idx = np.nonzero(covered)[0]
    P = pos[..., :3].reshape(-1, 3)[idx]; depth = pos[..., 3].reshape(-1)[idx]
    N = g_normal.reshape(-1, 3)[idx]                                       # raw, not normalised (:1071)
    rnd = random[idx]
    instance = np.floor(im[..., 0]).astype(np.int64).reshape(-1)[idx]
    material = np.floor(im[..., 1]).astype(np.int64).reshape(-1)[idx]
    n = len(idx)
    # --- the candidate of this frame (:1104-1151, EMISSIVE_LIT)
    c_dir, c_p, c_tmax, c_em, c_mat, graze = sc.select_light_candidate(rnd, P, N, instance, sample_emissive=emissive)
    info_pos, info_nrm = sc.info_position.copy(), sc.info_normal.copy()
    trace = (dot(c_dir, N) > 0) & (c_p > 0)
    if emissive:
        trace &= c_em != DONT_SAMPLE
    origin = (P + N * RAY_BIAS).astype(F)
    occ, og = sc.occluded(origin, c_dir, np.where(np.isfinite(c_tmax), c_tmax, 3.4e38), c_em if emissive else np.full(n, -1, np.int64))
    graze |= og & trace
    lit = trace & ~occ
    s_radiance = np.zeros((n, 4), F)
    if emissive:
        em = sc.bufs["materials"][c_mat]["emissive"]
        s_radiance[:, :3] = np.where(lit[:, None], F(255.0) * em[:, 3:4] * em[:, :3], F(0.0))
        s_radiance[:, 3] = np.where(trace, F(1.0), F(0.0))                 # input_radiance alpha = 1 whenever it was called
    else:                                                                  # input_radiance(ray, info, true, DONT_SAMPLE_EMISSIVE, false), :842-872
        in_cone = dot(c_dir, np.tile(sc.sun, (n, 1))) >= sc.cos_solar
        s_radiance[:, :3] = np.where((lit & in_cone)[:, None], sc.sun_color, F(0.0))
        s_radiance[:, 3] = np.where(trace & (occ | in_cone), F(1.0), F(0.0))   # an unoccluded ray outside the cone is "ambient": alpha 0
    # an occluded shadow ray rewrites info with the occluder (:526-533): position is traversal-order dependent, but such a
    # candidate has weight 0 and can never enter the reservoir, so it is not needed
    with np.errstate(all="ignore"):
        w_new = np.where(c_p > 0, luminance(s_radiance[:, :3]) / c_p, F(0.0))
    # --- history (:1089-1095, :181-190): the record at previous_uv = uv - velocity (the pixel's own under a static camera)
    pu = (xs.reshape(-1)[idx].astype(F) + F(0.5)) / F(W); pv = (ys.reshape(-1)[idx].astype(F) + F(0.5)) / F(H)
    if inp is not None:
        velocity = orc.readback(L.OUT_GBUFFER_VELOCITY_UV).reshape(-1, 4)[idx, :2]
        pu, pv = pu - velocity[:, 0], pv - velocity[:, 1]
    inside = (np.abs(pu - F(0.5)) < F(0.5)) & (np.abs(pv - F(0.5)) < F(0.5))
    pcx = np.clip(np.trunc(pu * F(W)).astype(np.int64), 0, W - 1); pcy = np.clip(np.trunc(pv * F(H)).astype(np.int64), 0, H - 1)
    fetched = previous.reshape(-1)[pcy * W + pcx].copy()
    fetched[~inside] = np.zeros((), L.PACKED_RESERVOIR)                     # var r: Reservoir — all zero, NOT unpack(zero record)
    prev = unpack_reservoir(fetched)
    for k_, v_ in prev.items():                                            # unpack of an absent record would yield lifetime 127 and NaN normals
        prev[k_] = np.where(inside.reshape((-1,) + (1,) * (v_.ndim - 1)), v_, 0).astype(v_.dtype)
    with np.errstate(all="ignore"):
        ratio = prev["visible_position"][:, 3] / depth
        ratio = np.where(ratio < 1.0, F(1.0) / ratio, ratio)
        depth_miss = ratio > F(1.05) * (F(1.0) + F(0.5) * rnd[:, 0])
        miss = depth_miss | (dot(N, prev["visible_normal"]) < F(0.9)) | (prev["visible_instance"] != instance)
    r = {k: (np.where(miss.reshape((-1,) + (1,) * (v.ndim - 1)), 0, v)).astype(v.dtype) for k, v in prev.items()}
    # --- block A (:1104-1153): a new candidate unless this is a validation frame with an established reservoir
    interval = int(b.settings.emissive_validate_interval if emissive else b.settings.direct_validate_interval)
    validation = frame_number % interval == 0
    do_new = np.full(n, not validation) | (r["count"] < F(4.0))
    s_now = dict(radiance=np.where(do_new[:, None], s_radiance, F(0.0)), random=rnd,
                 sample_position=np.where(do_new[:, None], info_pos, F(0.0)), sample_normal=np.where(do_new[:, None], info_nrm, F(0.0)))
    rand = fract(rnd[:, 0] + rnd[:, 1] + rnd[:, 2] + rnd[:, 3])
    with np.errstate(all="ignore"):
        w_sum = r["w_sum"] + w_new
        take = do_new & (rand < w_new / w_sum)
    r["w_sum"] = np.where(do_new, w_sum, r["w_sum"])
    r["w2_sum"] = np.where(do_new, r["w2_sum"] + w_new * w_new, r["w2_sum"])
    r["count"] = np.where(do_new, r["count"] + F(1.0), r["count"])
    # (*r).s = s replaces the WHOLE sample, visible_instance included; visible_position / visible_normal are refreshed for
    # every pixel afterwards (:1213-1214) but visible_instance is not: a record whose sample was never replaced keeps
    # instance 0 and fails the instance test of the next frame unless the pixel shows instance 0
    s_visible_position = np.concatenate([P, depth[:, None]], 1)
    whole = dict(radiance=s_radiance, random=rnd, sample_position=info_pos, sample_normal=info_nrm, visible_position=s_visible_position,
                 visible_normal=N)
    for k, v in whole.items():
        r[k] = np.where(take[:, None], v, r[k]).astype(F)
    r["visible_instance"] = np.where(take, instance, r["visible_instance"])
    m = F(b.settings.max_temporal_reuse_count)
    over = do_new & (r["count"] > m)
    with np.errstate(all="ignore"):
        r["w_sum"] = np.where(over, r["w_sum"] * (m / r["count"]), r["w_sum"])
        r["w2_sum"] = np.where(over, r["w2_sum"] * (m / r["count"]), r["w2_sum"])
        r["count"] = np.where(over, m, r["count"])
    reset = np.zeros(n, bool)
    if validation:
        # --- block B (:1155-1208): re-derive the reservoir sample's light point from ITS random numbers and visible point, shoot the
        # ray from the CURRENT visible point towards the stored sample position, compare what arrives with what was stored
        with np.errstate(all="ignore"):
            v_dir, v_p, v_tmax, v_em, v_mat, v_graze = sc.select_light_candidate(r["random"], r["visible_position"][:, :3].copy(), r["visible_normal"].copy(), instance,
                                                                                 sample_emissive=emissive)
            v_info_pos, v_info_nrm = sc.info_position.copy(), sc.info_normal.copy()
            ray_dir = normalize(r["sample_position"][:, :3] - P).astype(F)
            v_trace = (dot(v_dir, r["visible_normal"]) > 0) & (v_p > 0)
            if emissive:
                v_trace &= v_em != DONT_SAMPLE
            ray_dir = np.where(np.isfinite(ray_dir), ray_dir, F(0.0))
        v_occ, v_og = sc.occluded(origin, ray_dir, np.where(np.isfinite(v_tmax), v_tmax, 3.4e38), v_em if emissive else np.full(n, -1, np.int64))
        established = r["count"] >= F(4.0)
        # an occluded validation ray of an established reservoir stores the occluder's position, which depends on the
        # traversal order (any-hit): not reproducible by brute force, left out
        graze |= (v_graze & v_trace) | (v_og & v_trace) | (v_occ & v_trace & established)
        validate_radiance = np.zeros((n, 4), F)
        if emissive:
            v_em_mat = sc.bufs["materials"][v_mat]["emissive"]
            validate_radiance[:, :3] = np.where((v_trace & ~v_occ)[:, None], F(255.0) * v_em_mat[:, 3:4] * v_em_mat[:, :3], F(0.0))
            validate_radiance[:, 3] = np.where(v_trace, F(1.0), F(0.0))
        else:
            v_cone = dot(ray_dir, np.tile(sc.sun, (n, 1))) >= sc.cos_solar
            validate_radiance[:, :3] = np.where((v_trace & ~v_occ & v_cone)[:, None], sc.sun_color, F(0.0))
            validate_radiance[:, 3] = np.where(v_trace & (v_occ | v_cone), F(1.0), F(0.0))
        s_val = dict(radiance=np.where(established[:, None], validate_radiance, s_now["radiance"]),
                     random=np.where(established[:, None], r["random"], rnd),
                     sample_position=np.where(established[:, None], v_info_pos, s_now["sample_position"]),
                     sample_normal=np.where(established[:, None], v_info_nrm, s_now["sample_normal"]))
        with np.errstate(all="ignore"):
            lum_ratio = luminance(validate_radiance[:, :3]) / np.fmax(luminance(r["radiance"][:, :3]), F(0.0001))
            reset = (lum_ratio > F(1.25)) | (lum_ratio < F(0.8))
            w_val = np.where(v_p > 0, luminance(s_val["radiance"][:, :3]) / v_p, F(0.0))
        # the reservoir as it stands BEFORE the reset is what store_previous_spatial_reservoir receives (:1199-1202): its own
        # visible point (the history's, or this frame's where block A replaced the sample), w not yet recomputed
        temporal_emissive_numpy.before_reset = pack_records(r, r["visible_position"], r["visible_normal"])
        # a reset installs `s` whatever its weight: where s is this frame's OCCLUDED candidate, its position is the occluder's
        # (traversal-order dependent, :526-533) — left out
        graze |= reset & ~established & occ & trace
        for k, v in s_val.items():
            r[k] = np.where(reset[:, None], v, r[k]).astype(F)
        r["visible_instance"] = np.where(reset, instance, r["visible_instance"])
        r["count"] = np.where(reset, F(1.0), r["count"]); r["lifetime"] = np.where(reset, F(0.0), r["lifetime"])
        r["w_sum"] = np.where(reset, w_val, r["w_sum"]); r["w2_sum"] = np.where(reset, w_val * w_val, r["w2_sum"])
    temporal_emissive_numpy.reset = reset
    with np.errstate(all="ignore"):
        total = r["count"] * luminance(r["radiance"][:, :3])
        r["w"] = np.where(total > 0, r["w_sum"] / total, F(0.0))
        r["lifetime"] = r["lifetime"] + F(1.0)
        variance = r["w2_sum"] / r["count"] - np.power(r["w_sum"] / r["count"], F(2.0))
        variance = np.fmin(np.where(r["count"] < 1.0, variance, variance / r["count"]), F(10.0))
        variance_scale = r["w2_sum"] / r["count"] / np.fmax(r["count"], F(1.0))       # the minuend: the estimate is a difference of two such terms
    # --- pack (:108-136) with this frame's visible point
    packed = np.zeros(n, L.PACKED_RESERVOIR)
    packed["reservoir"][:, 0] = pack_f16x2(r["count"], r["w"]); packed["reservoir"][:, 1] = pack_f16x2(r["w_sum"], r["w2_sum"])
    packed["radiance"][:, 0] = pack_f16x2(r["radiance"][:, 0], r["radiance"][:, 1])
    packed["radiance"][:, 1] = pack_f16x2(r["radiance"][:, 2], r["radiance"][:, 3])
    packed["random"][:, 0] = pack_unorm16x2(r["random"][:, 0], r["random"][:, 1])
    packed["random"][:, 1] = pack_unorm16x2(r["random"][:, 2], r["random"][:, 3])
    packed["visible_position"] = np.concatenate([P, depth[:, None]], 1)
    packed["sample_position"] = np.concatenate([r["sample_position"][:, :3], r["visible_instance"].astype(F)[:, None]], 1)
    packed["visible_normal"] = pack_snorm8x4(np.concatenate([N, (r["lifetime"] / F(127.0) - F(1.0))[:, None]], 1))
    packed["sample_normal"] = pack_snorm8x4(np.concatenate([r["sample_normal"], r["sample_position"][:, 3:4]], 1))
    # --- output (:1230-1259, no RENDER_EMISSIVE on this pipeline)
    eye = np.array(list((inp.view if inp is not None else b.view).world_position), F)
    view = normalize(eye - P)
    with np.errstate(all="ignore"):
        surface = sc.bufs["materials"][material]
        out = shading(view, N, normalize(r["sample_position"][:, :3] - P), surface, r["radiance"], sc.ambient)
        out = out * r["w"][:, None]
        if not emissive:                                                   # RENDER_EMISSIVE on the sun pipeline (light.rs:409-412)
            out = out + F(255.0) * surface["emissive"][:, 3:4] * surface["emissive"][:, :3]
    # the invalidation scatter (:1092-1095): a rejected history zeroes the previous-SPATIAL record it was fetched from
    in_frame = (np.abs(pu - F(0.5)) <= F(0.5)) & (np.abs(pv - F(0.5)) <= F(0.5))
    scatter_targets = (pcy * W + pcx)[miss & in_frame]
    temporal_emissive_numpy.scatter_targets = scatter_targets
    temporal_emissive_numpy.scatter_writers = idx[miss & in_frame]
    return idx, packed, out.astype(F), (variance.astype(F), variance_scale.astype(F)), graze, take, miss


@pytest.mark.parametrize("scene,size", [("cornell", (72, 72)), ("soup5", (80, 56))])
def test_oracle_temporal_reuse_equals_independent_numpy_restatement(scene, size):
    if scene.startswith("soup"):
        from bevy_hikari_b200 import scenes
        scenes.SCENE_BUILDERS[scene] = lambda: scenes.soup(int(scene[4:]))
    b = Bench(scene, size[0], size[1], taa=plugin.TAA_NONE, upscale_ratio=1.0, temporal_reuse=1, denoise=0, indirect_bounces=1,
              emissive_spatial_reuse=0, indirect_spatial_reuse=0, max_temporal_reuse_count=3)     # M clamp reached at frame 4
    orc = b.oracle()
    noise = plugin.load_noise()
    H, W = size[1], size[0]
    replaced = kept = clamped = rejected = 0
    for f in range(1, 5):                                  # emissive_validate_interval = 5: frames 1..4 take new candidates only
        inp = b.inputs(f)
        assert f % inp.frame.emissive_validate_interval != 0
        read_buffer = L.OUT_RESERVOIR_0 + 2 + (f % 2)      # light.rs:518-546: binding 0 = buf[base + head], binding 1 = buf[base + 1 - head]
        previous = orc.readback(read_buffer).copy()
        orc.render_frame(inp)
        if f == 1:
            assert not previous.view(np.uint8).any()       # zeroed history
            continue
        idx, packed, out, variance, graze, take, miss = temporal_emissive_numpy(b, orc, f, noise, previous)
        written = orc.readback(L.OUT_RESERVOIR_0 + 2 + 1 - (f % 2)).reshape(-1)[idx]
        clean = ~graze
        assert clean.mean() > 0.9
        for field in ("radiance", "random", "visible_position", "visible_normal", "sample_normal"):
            same = (written[field] == packed[field]) if written[field].ndim == 1 else (written[field] == packed[field]).all(-1)
            assert same[clean].mean() >= 0.995, (f, field, float(same[clean].mean()))
        # sample_position is stored in fp32: a kept sample is copied bit for bit, a replaced one carries the hit point, whose
        # distance the oracle computes in fp32 object space and this checker in float64 world space
        sp_same = (written["sample_position"] == packed["sample_position"]).all(-1)
        assert sp_same[clean & ~take].mean() >= 0.999
        err = np.abs(written["sample_position"][clean] - packed["sample_position"][clean]).max(-1)
        assert (err <= 5e-6).mean() >= 0.98 and err.max() <= 2e-4, (f, float((err <= 5e-6).mean()), float(err.max()))
        # (count, w, w_sum, w2_sum) as f16: count exactly, the sums within an ulp (fp32 order of the candidate's weight)
        gc, gw = unpack_f16x2(written["reservoir"][:, 0]); wc, ww = unpack_f16x2(packed["reservoir"][:, 0])
        assert np.array_equal(gc[clean], wc[clean])
        for got, want in ((gw, ww), unpack_f16x2(written["reservoir"][:, 1])[:1] + unpack_f16x2(packed["reservoir"][:, 1])[:1]):
            assert (ulps16(got[clean], want[clean]) <= 1).mean() >= 0.995
        render = orc.readback(L.OUT_RENDER_EMISSIVE).astype(F).reshape(-1, 4)[idx]
        d = ulps16(render[:, :3], out).max(-1)
        assert (d[clean] <= 1).mean() >= 0.99 and (d[clean] == 0).mean() >= 0.97, (f, float((d[clean] <= 1).mean()), float((d[clean] == 0).mean()))
        got_var = orc.readback(L.OUT_VARIANCE_EMISSIVE).reshape(-1)[idx]
        variance, scale = variance
        assert (np.abs(got_var[clean] - variance[clean]) <= 1e-4 * scale[clean] + 1e-7).all()     # cancellation: relative to the terms, not to the difference
        replaced += int((take & clean).sum()); kept += int((~take & clean & ~miss).sum()); clamped += int((wc[clean] == 3).sum() if f == 4 else 0)
        rejected += int((miss & clean).sum())
    # both outcomes of the update, the clamp and the rejection of a stale record (instance test) were exercised
    assert replaced > 50 and kept > 50 and clamped > 50 and rejected > 0, (replaced, kept, clamped, rejected)


def test_oracle_temporal_reuse_and_invalidation_scatter_under_camera_motion():
    """translating camera: history is fetched at the re-projected pixel, rejected where depth / normal / instance disagree, and
    every rejection zeroes the previous-spatial record at the re-projected pixel (light.wgsl:1092-1095, the racy scatter the
    kernels resolve deterministically) — the buffer the oracle leaves behind must be the buffer before the pass with exactly
    those records replaced by a packed empty reservoir"""
    W, H = 80, 64
    b = Bench("cornell", W, H, taa=plugin.TAA_NONE, upscale_ratio=1.0, temporal_reuse=1, denoise=0, indirect_bounces=1,
              emissive_spatial_reuse=0, indirect_spatial_reuse=0)
    orc = b.oracle()
    noise = plugin.load_noise()
    empty = np.zeros((), L.PACKED_RESERVOIR)
    empty["visible_normal"] = pack_snorm8x4(np.array([[0.0, 0.0, 0.0, -1.0]], F))[0]        # lifetime 0 -> 0 / 127 - 1
    scattered = rejected = 0
    for f in range(1, 5):
        inp = b.moving_inputs(f, step=(0.08, 0.03, -0.05))
        head = f % 2
        orc.prepass(inp)
        orc.run_pass(inp, 0); orc.run_pass(inp, 1)                              # albedo, sun pass (scatters into the same buffer)
        previous = orc.readback(L.OUT_RESERVOIR_0 + 2 + head).copy()
        spatial_before = orc.readback(L.OUT_RESERVOIR_0 + 4 + head).reshape(-1).copy()
        orc.run_pass(inp, 2)                                                    # the emissive pass under test
        spatial_after = orc.readback(L.OUT_RESERVOIR_0 + 4 + head).reshape(-1).copy()
        written = orc.readback(L.OUT_RESERVOIR_0 + 2 + 1 - head).reshape(-1)
        render = orc.readback(L.OUT_RENDER_EMISSIVE).astype(F).reshape(-1, 4)
        orc.run_pass(inp, 4)                                                    # the rest of the frame's light passes
        if f == 1:
            continue
        idx, packed, out, variance, graze, take, miss = temporal_emissive_numpy(b, orc, f, noise, previous, inp)
        clean = ~graze
        w = written[idx]
        for field in ("radiance", "random", "visible_position", "visible_normal", "sample_normal"):
            same = (w[field] == packed[field]) if w[field].ndim == 1 else (w[field] == packed[field]).all(-1)
            assert same[clean].mean() >= 0.99, (f, field, float(same[clean].mean()))
        gc, _ = unpack_f16x2(w["reservoir"][:, 0]); wc, _ = unpack_f16x2(packed["reservoir"][:, 0])
        assert (gc[clean] == wc[clean]).mean() >= 0.995
        d = ulps16(render[idx, :3], out).max(-1)
        assert (d[clean] <= 1).mean() >= 0.985, (f, float((d[clean] <= 1).mean()))
        # scatter: covered pixels that rejected their history + background pixels (:1059-1063 write their own record)
        expected = spatial_before.copy()
        background = np.setdiff1d(np.arange(W * H), idx)
        bg = np.zeros((), L.PACKED_RESERVOIR)
        bg["visible_normal"] = empty["visible_normal"]; bg["reservoir"][0] = pack_f16x2(F(1.0), F(0.0))     # set_reservoir(&r, s, 0.0): count 1
        # the oracle's rule for the race: writers in raster order, the last one wins (deviation 2)
        writes = [(int(wr), int(tg), empty) for wr, tg in zip(temporal_emissive_numpy.scatter_writers, temporal_emissive_numpy.scatter_targets)]
        writes += [(int(px), int(px), bg) for px in background]
        for _, tg, value in sorted(writes, key=lambda t: t[0]):
            expected[tg] = value
        same = (expected.view(np.uint8).reshape(-1, 64) == spatial_after.view(np.uint8).reshape(-1, 64)).all(1)
        assert same.mean() >= 0.998, (f, float(same.mean()))
        scattered += len(temporal_emissive_numpy.scatter_targets); rejected += int(miss.sum())
    assert scattered > 200 and rejected >= scattered


def test_oracle_validation_frames_equal_independent_numpy_restatement():
    """the validation branch of direct_lit (light.wgsl:1155-1208) with a light whose radiance changes between frames (a
    StandardMaterial edited at run time): on every second frame the reservoir's sample is re-derived from its own random
    numbers, re-traced from the current visible point and compared with the stored radiance; a change beyond -20 % / +25 %
    resets the reservoir to the validated sample.  Frames 2 and 4 run both blocks (count < 4), frames 6 and 8 validation only."""
    W, H = 72, 72
    b = Bench("cornell", W, H, taa=plugin.TAA_NONE, upscale_ratio=1.0, temporal_reuse=1, denoise=0, indirect_bounces=1,
              emissive_spatial_reuse=0, indirect_spatial_reuse=0, emissive_validate_interval=2, direct_validate_interval=2)
    orc = b.oracle()
    noise = plugin.load_noise()
    w = b.world
    light_material = int(w.buffers()["instances"][int(w.buffers()["emissives"][0]["instance"])]["material"])
    base = b.scene.materials[light_material].copy()
    alpha = {1: 1.0, 2: 1.0, 3: 1.0, 4: 2.0, 5: 2.0, 6: 0.7, 7: 0.7, 8: 0.72}      # x2 at frame 4 (reset), x0.35 at 6 (reset), +3 % at 8 (kept)
    resets, kept_validations, validation_only = {}, 0, 0
    for f in range(1, 9):
        m = base.copy()
        m["emissive"] = (base["emissive"][0], base["emissive"][1], base["emissive"][2], base["emissive"][3] * alpha[f])
        w.set_material(light_material, m)
        w.prepare_materials(); w.previous_transform_system(); w.prepare_instances()
        orc.update_instances_desc(w.scene_desc())
        inp = b.inputs(f)
        previous = orc.readback(L.OUT_RESERVOIR_0 + 2 + (f % 2)).copy()
        orc.render_frame(inp)
        if f == 1:
            continue
        idx, packed, out, variance, graze, take, miss = temporal_emissive_numpy(b, orc, f, noise, previous)
        written = orc.readback(L.OUT_RESERVOIR_0 + 2 + 1 - (f % 2)).reshape(-1)[idx]
        clean = ~graze
        assert clean.mean() > 0.85, (f, float(clean.mean()))
        for field in ("radiance", "random", "visible_position", "visible_normal", "sample_normal"):
            same = (written[field] == packed[field]) if written[field].ndim == 1 else (written[field] == packed[field]).all(-1)
            assert same[clean].mean() >= 0.99, (f, field, float(same[clean].mean()))
        gc, _ = unpack_f16x2(written["reservoir"][:, 0]); wc, _ = unpack_f16x2(packed["reservoir"][:, 0])
        assert (gc[clean] == wc[clean]).mean() >= 0.995, (f, float((gc[clean] == wc[clean]).mean()))
        render = orc.readback(L.OUT_RENDER_EMISSIVE).astype(F).reshape(-1, 4)[idx]
        d = ulps16(render[:, :3], out).max(-1)
        assert (d[clean] <= 1).mean() >= 0.985, (f, float((d[clean] <= 1).mean()))
        if f % 2 == 0:
            reset = temporal_emissive_numpy.reset
            resets[f] = int((reset & clean).sum())
            kept_validations += int((~reset & clean).sum())
            validation_only += int((wc[clean] >= 4).sum()) if f >= 6 else 0
    # the light doubled before frame 4 and dropped to 35 % before frame 6: every lit reservoir resets there.  Without a change
    # (frame 2) or with +3 % (frame 8) only reservoirs whose stored radiance is 0 while the validation ray arrives (or the
    # reverse) reset
    assert resets[4] > 2 * resets[2] and resets[6] > 2 * resets[8] and kept_validations > 1000 and validation_only > 200, (resets, kept_validations, validation_only)


def test_oracle_validation_scatter_equals_independent_numpy_restatement():
    """the third kind of write into the previous-spatial buffer (light.wgsl:1199-1202): a validation frame that finds the light
    changed stores the reservoir AS IT WAS into the record it was fetched from, before resetting it — compared record by record
    with the buffer the oracle's emissive pass leaves behind"""
    W, H = 64, 64
    b = Bench("cornell", W, H, taa=plugin.TAA_NONE, upscale_ratio=1.0, temporal_reuse=1, denoise=0, indirect_bounces=1,
              emissive_spatial_reuse=0, indirect_spatial_reuse=0, emissive_validate_interval=2, direct_validate_interval=5)
    orc = b.oracle()
    noise = plugin.load_noise()
    w = b.world
    light_material = int(w.buffers()["instances"][int(w.buffers()["emissives"][0]["instance"])]["material"])
    base = b.scene.materials[light_material].copy()
    empty = np.zeros((), L.PACKED_RESERVOIR)
    empty["visible_normal"] = pack_snorm8x4(np.array([[0.0, 0.0, 0.0, -1.0]], F))[0]
    stored = 0
    for f in range(1, 5):
        m = base.copy()
        m["emissive"] = (base["emissive"][0], base["emissive"][1], base["emissive"][2], base["emissive"][3] * (2.0 if f >= 4 else 1.0))
        w.set_material(light_material, m)
        w.prepare_materials(); w.previous_transform_system(); w.prepare_instances()
        orc.update_instances_desc(w.scene_desc())
        inp = b.inputs(f)
        head = f % 2
        orc.prepass(inp)
        orc.run_pass(inp, 0); orc.run_pass(inp, 1)
        previous = orc.readback(L.OUT_RESERVOIR_0 + 2 + head).copy()
        before = orc.readback(L.OUT_RESERVOIR_0 + 4 + head).reshape(-1).copy()
        orc.run_pass(inp, 2)
        after = orc.readback(L.OUT_RESERVOIR_0 + 4 + head).reshape(-1).copy()
        orc.run_pass(inp, 4)
        if f != 4:
            continue
        idx, packed, out, variance, graze, take, miss = temporal_emissive_numpy(b, orc, f, noise, previous)
        reset = temporal_emissive_numpy.reset
        records = temporal_emissive_numpy.before_reset
        expected = before.copy()
        background = np.setdiff1d(np.arange(W * H), idx)
        bg = np.zeros((), L.PACKED_RESERVOIR)
        bg["visible_normal"] = empty["visible_normal"]; bg["reservoir"][0] = pack_f16x2(F(1.0), F(0.0))
        writes = [(int(px), int(px), empty) for px in idx[miss]]                    # rejected history (static camera: own pixel)
        writes += [(int(px), int(px), bg) for px in background]
        order = {int(px): k for k, px in enumerate(idx)}
        for px in idx[reset]:                                                       # the validation store comes after the miss store of the same pixel
            writes.append((int(px) + 0.5, int(px), records[order[int(px)]]))
        for _, tg, value in sorted(writes, key=lambda t: t[0]):
            expected[tg] = value
        clean_px = idx[~graze]
        same = (expected.view(np.uint8).reshape(-1, 64) == after.view(np.uint8).reshape(-1, 64)).all(1)
        assert same[clean_px].mean() >= 0.99, float(same[clean_px].mean())
        assert same[background].all()
        stored = int(reset[~graze].sum())
    assert stored > 300                                                              # the doubled light reset every lit reservoir


def test_oracle_sun_pass_with_history_and_validation_equals_independent_numpy_restatement():
    """the same entry point specialised as the sun pipeline (sample_directional, no emissive walk, RENDER_EMISSIVE) on minimal.rs over
    frames 2-7: history, M clamp, and the validation branch on frames 3 and 6 (direct_validate_interval = 3; frame 6 validates
    established reservoirs)"""
    W, H = 88, 60
    b = Bench("minimal", W, H, taa=plugin.TAA_NONE, upscale_ratio=1.0, temporal_reuse=1, denoise=0, indirect_bounces=1,
              emissive_spatial_reuse=0, indirect_spatial_reuse=0, max_temporal_reuse_count=6)
    orc = b.oracle()
    noise = plugin.load_noise()
    validated = 0
    for f in range(1, 8):
        inp = b.inputs(f)
        previous = orc.readback(L.OUT_RESERVOIR_0 + 0 + (f % 2)).copy()          # buffers [0, 1]: the sun pass's temporal pair (light.rs:518-546)
        orc.render_frame(inp)
        if f == 1:
            continue
        idx, packed, out, variance, graze, take, miss = temporal_emissive_numpy(b, orc, f, noise, previous, emissive=False)
        written = orc.readback(L.OUT_RESERVOIR_0 + 0 + 1 - (f % 2)).reshape(-1)[idx]
        clean = ~graze
        assert clean.mean() > 0.9
        for field in ("radiance", "random", "visible_position", "visible_normal", "sample_normal"):
            same = (written[field] == packed[field]) if written[field].ndim == 1 else (written[field] == packed[field]).all(-1)
            assert same[clean].mean() >= 0.99, (f, field, float(same[clean].mean()))
        gc, _ = unpack_f16x2(written["reservoir"][:, 0]); wc, _ = unpack_f16x2(packed["reservoir"][:, 0])
        assert (gc[clean] == wc[clean]).mean() >= 0.995, (f, float((gc[clean] == wc[clean]).mean()))
        render = orc.readback(L.OUT_RENDER_DIRECT).astype(F).reshape(-1, 4)[idx]
        d = ulps16(render[:, :3], out).max(-1)
        assert (d[clean] <= 1).mean() >= 0.99, (f, float((d[clean] <= 1).mean()))
        if f % 3 == 0:
            validated += int((wc[clean] >= 4).sum()) if f == 6 else 0
    assert validated > 500
