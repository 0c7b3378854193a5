"""`spatial_reuse` (a third of the frame time), both pipelines, pinned from the outside.  The pass works on
reservoirs and the G-buffer only (no rays), so a SECOND, independent restatement of src/shaders/light.wgsl:1500-1684 in
whole-image numpy arithmetic, written from the WGSL, can be held against the oracle record by record:
  own temporal reservoir -> start from the previous SPATIAL reservoir while the own lifetime is within bounds (:1545-1547) ->
  merge self -> for i in 1..N: Fibonacci-spiral offset rotated by the own sample's random numbers and hash(frame number)
  (:1568-1572, utils.wgsl hash), truncation to a pixel, bounds / depth-ratio / normal / count / back-facing rejections
  (:1577-1606), the screen-space depth march with its tap count rule (:1608-1628), the Jacobian clamp (:985-1004), the
  merge with the neighbour's sample deciding by the NEIGHBOUR's random numbers (:146-179) -> M clamp -> r.w -> pack, render,
  and the variance only where the own count was <= 4 (:1664-1669).
Inputs are read back from the oracle between its passes (temporal buffer of this frame, previous-spatial buffer as the
temporal passes left it); outputs compared: the written spatial reservoir buffer (all 64 bytes of every record), render[signal]
and the variance plane where the pass writes it.  Measured (cornell emissive + indirect pipelines, minimal.rs; frames 3-5 with the M
clamp reached): 99.98 - 100 % of the records bit-identical, render 99.9 - 100 % bit-identical and never more than 1 f16 ulp
apart.  CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_direct_lit_numpy import F, GOLDEN_RATIO, TAU, dot, fract, luminance, normalize, ulps16
from tests.test_indirect_numpy import shading
from tests.test_temporal_numpy import pack_f16x2, pack_snorm8x4, pack_unorm16x2, unpack_f16x2, unpack_reservoir

U = np.uint32
F32_MAX = F(3.402823466e38)
SAMPLE_FIELDS = ("radiance", "random", "visible_position", "visible_normal", "visible_instance", "sample_position", "sample_normal")


def hash_u32(v):                                         # utils.wgsl hash
    with np.errstate(over="ignore"):
        s = U(v) ^ U(2747636419)
        s = U(s * U(2654435769)); s = s ^ (s >> U(16))
        s = U(s * U(2654435769)); s = s ^ (s >> U(16))
        s = U(s * U(2654435769))
    return s


def merge(r, q, p, mask):
    """merge_reservoir(&r, q, p) for the pixels in `mask` (:173-179 with update_reservoir :146-171)"""
    with np.errstate(all="ignore"):
        w_new = p * q["w"] * q["count"]
        count0 = r["count"].copy()
        w_sum = r["w_sum"] + w_new
        take = mask & (fract(q["random"][:, 0] + q["random"][:, 1] + q["random"][:, 2] + q["random"][:, 3]) < w_new / w_sum)
    r["w_sum"] = np.where(mask, w_sum, r["w_sum"])
    r["w2_sum"] = np.where(mask, r["w2_sum"] + w_new * w_new, r["w2_sum"])
    r["count"] = np.where(mask, count0 + q["count"], r["count"])
    for k in SAMPLE_FIELDS:
        t = take.reshape((-1,) + (1,) * (r[k].ndim - 1))
        r[k] = np.where(t, q[k], r[k])
    return take


def compute_jacobian(q, s_visible_position):             # :985-1004, r = the own sample
    with np.errstate(all="ignore"):
        normal = q["sample_normal"]
        a = s_visible_position[:, :3] - q["sample_position"][:, :3]
        bq = q["visible_position"][:, :3] - q["sample_position"][:, :3]
        cos_phi_1 = np.abs(dot(normalize(a), normal))
        cos_phi_2 = np.abs(dot(normalize(bq), normal))
        term_1 = cos_phi_1 / np.fmax(F(0.0001), cos_phi_2)
        num = np.sqrt(dot(bq, bq)); num = num * num
        den = np.sqrt(dot(a, a)); den = den * den
        term_2 = num / np.fmax(den, F(0.0001))
        return np.fmin(np.fmax(term_1 * term_2, F(1.0)), F(50.0))


def spatial_numpy(b, orc, frame_number, emissive, temporal, previous_spatial):
    pos_full = orc.readback(L.OUT_GBUFFER_POSITION)
    im_full = orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL)
    DH, DW = pos_full.shape[:2]
    H, W = temporal.shape[:2]                              # render size = ceil(size / ratio) (light.rs:622-624)
    ratio = F(b.settings.upscale_ratio)
    sel = F(-0.25) if (frame_number & 1) == 0 else F(0.25)

    def deferred(u_, v_):
        """jittered_deferred_coords (:1007-1017): i32((uv -+ 0.25 texel * (ratio - 1)) * size); identity at ratio 1"""
        du = u_ + sel * (F(1.0) / F(DW)) * (ratio - F(1.0)); dv = v_ + sel * (F(1.0) / F(DH)) * (ratio - F(1.0))
        return np.trunc(du * F(DW)).astype(np.int64), np.trunc(dv * F(DH)).astype(np.int64)
    n = H * W
    ys, xs = [a.reshape(-1) for a in np.meshgrid(np.arange(H), np.arange(W), indexing="ij")]
    u = (xs.astype(F) + F(0.5)) / F(W); v = (ys.astype(F) + F(0.5)) / F(H)
    gx, gy = deferred(u, v)
    depth_img = pos_full[..., 3]
    depth = depth_img[gy, gx]
    position = pos_full[gy, gx, :3]
    mats = b.world.buffers()["materials"][np.floor(im_full[gy, gx, 1]).astype(np.int64)]
    ambient = np.array(list(b.lights.ambient_color), F)[:3]
    own = unpack_reservoir(temporal.reshape(-1))
    covered = depth >= F(1.1920929e-7)
    s = {k: own[k].copy() for k in SAMPLE_FIELDS}
    use_variance = own["count"] <= F(4.0)
    limit = F32_MAX if b.settings.max_reservoir_lifetime <= 1.0 else F(b.settings.max_reservoir_lifetime)
    prev = unpack_reservoir(previous_spatial.reshape(-1))                 # static camera: previous_uv = uv, inside (0, 1)
    from_previous = own["lifetime"] <= limit
    r = {k: np.where(from_previous.reshape((-1,) + (1,) * (own[k].ndim - 1)), prev[k], own[k]) for k in own}
    view = normalize(np.array(list(b.view.world_position), F) - position)
    with np.errstate(all="ignore"):
        own_dir = normalize(s["sample_position"][:, :3] - s["visible_position"][:, :3])
        target = luminance(s["radiance"][:, :3]) if emissive else luminance(shading(view, s["visible_normal"], own_dir, mats, s["radiance"], ambient))
    merge(r, own, target, covered)
    r["visible_position"] = np.where(covered[:, None], s["visible_position"], r["visible_position"])
    r["visible_normal"] = np.where(covered[:, None], s["visible_normal"], r["visible_normal"])
    count_n, reach = (8, F(10.0)) if emissive else (16, F(20.0))
    rotation = s["random"][:, 0] + s["random"][:, 1] + s["random"][:, 2] + s["random"][:, 3]
    frame_random = F(hash_u32(frame_number)) / F(4294967295.0)
    stats = dict(merged=0, depth=0, normal=0, back=0, occluded=0, outside=0)
    for i in range(1, count_n + 1):
        angle = TAU * fract(F(i) * GOLDEN_RATIO + rotation + frame_random)
        radius = np.sqrt(F(i) / F(count_n)) * reach
        ox, oy = radius * np.cos(angle), radius * np.sin(angle)
        sx = np.trunc(ox + xs.astype(F)).astype(np.int64); sy = np.trunc(oy + ys.astype(F)).astype(np.int64)
        su, sv = (sx.astype(F) + F(0.5)) / F(W), (sy.astype(F) + F(0.5)) / F(H)
        ok = covered & ~((su < 0) | (sv < 0) | (su > 1) | (sv > 1))
        stats["outside"] += int((covered & ~ok).sum())
        cx, cy = np.clip(sx, 0, W - 1), np.clip(sy, 0, H - 1)
        sidx = cy * W + cx
        sgx, sgy = deferred(su, sv)                                             # sample_deferred_coords (:1576)
        sample_depth = depth_img[np.clip(sgy, 0, DH - 1), np.clip(sgx, 0, DW - 1)]
        q = {k: own[k][sidx] for k in own}
        with np.errstate(all="ignore"):
            depth_ratio = depth / sample_depth
            passed = ok & ~((depth_ratio < F(0.9)) | (depth_ratio > F(1.1)))
            stats["depth"] += int((ok & ~passed).sum()); ok = passed
            passed = ok & ~((q["count"] < F(1.1920929e-7)) | (dot(s["visible_normal"], q["visible_normal"]) < F(0.866)))
            stats["normal"] += int((ok & ~passed).sum()); ok = passed
            direction = normalize(q["sample_position"][:, :3] - s["visible_position"][:, :3])
            passed = ok & ~(dot(direction, s["visible_normal"]) < 0)
            stats["back"] += int((ok & ~passed).sum()); ok = passed
            # depth march (:1608-1628)
            interval = np.fmax(F(1.0), radius / F(5.0))
            taps = int(radius / interval)
            length = np.sqrt(ox * ox + oy * oy)
            ux, uy = ox / length, oy / length
            occluded = np.zeros(n, bool)
            for j in range(1, taps + 1):
                dist = F(j) * interval
                tu, tv = u + (dist * ux) / F(W), v + (dist * uy) / F(H)
                tx, ty = deferred(tu, tv)                                       # tap_deferred_coords (:1617); textureLoad outside = 0
                inside = (tx >= 0) & (tx < DW) & (ty >= 0) & (ty < DH)
                tap_depth = np.where(inside, depth_img[np.clip(ty, 0, DH - 1), np.clip(tx, 0, DW - 1)], F(0.0))
                t = F(j) / F(taps + 1)
                ref = depth * (F(1.0) - t) + sample_depth * t
                occluded |= ~occluded & (tap_depth > ref + F(0.00001))
            passed = ok & ~occluded
            stats["occluded"] += int((ok & ~passed).sum()); ok = passed
            jac = np.where(q["sample_position"][:, 3] > F(0.5), compute_jacobian(q, s["visible_position"]), F(1.0))
            if emissive:
                tgt = luminance(q["radiance"][:, :3]) / jac
            else:
                tgt = luminance(shading(view, s["visible_normal"], direction, mats, q["radiance"], ambient)) / jac
        merge(r, q, tgt, ok)
        stats["merged"] += int(ok.sum())
    m = F(b.settings.max_spatial_reuse_count)
    with np.errstate(all="ignore"):
        over = covered & (r["count"] > m)
        r["w_sum"] = np.where(over, r["w_sum"] * (m / r["count"]), r["w_sum"])
        r["w2_sum"] = np.where(over, r["w2_sum"] * (m / r["count"]), r["w2_sum"])
        r["count"] = np.where(over, m, r["count"])
        out = shading(view, s["visible_normal"], normalize(r["sample_position"][:, :3] - s["visible_position"][:, :3]), mats, r["radiance"], ambient)
        total = r["count"] * (luminance(r["radiance"][:, :3]) if emissive else luminance(out))
        w = np.where(total > 0, r["w_sum"] / total, F(0.0))
        r["w"] = np.where(covered, w, r["w"])
        r["lifetime"] = np.where(covered, r["lifetime"] + F(1.0), r["lifetime"])
        variance = r["w2_sum"] / r["count"] - np.power(r["w_sum"] / r["count"], F(2.0))
        variance = np.fmin(np.where(r["count"] < 1.0, variance, variance / r["count"]), F(10.0))
        render = out * r["w"][:, None]
    # background pixels store pack(unpack(own)) and render 0 (:1526-1530)
    for k in r:
        c = covered.reshape((-1,) + (1,) * (r[k].ndim - 1))
        r[k] = np.where(c, r[k], own[k])
    packed = np.zeros(n, L.PACKED_RESERVOIR)
    packed["reservoir"][:, 0] = pack_f16x2(r["count"], r["w"]); packed["reservoir"][:, 1] = pack_f16x2(r["w_sum"], r["w2_sum"])
    packed["radiance"][:, 0] = pack_f16x2(r["radiance"][:, 0], r["radiance"][:, 1])
    packed["radiance"][:, 1] = pack_f16x2(r["radiance"][:, 2], r["radiance"][:, 3])
    packed["random"][:, 0] = pack_unorm16x2(r["random"][:, 0], r["random"][:, 1])
    packed["random"][:, 1] = pack_unorm16x2(r["random"][:, 2], r["random"][:, 3])
    packed["visible_position"] = r["visible_position"]
    packed["sample_position"] = np.concatenate([r["sample_position"][:, :3], r["visible_instance"].astype(F)[:, None]], 1)
    packed["visible_normal"] = pack_snorm8x4(np.concatenate([r["visible_normal"], (r["lifetime"] / F(127.0) - F(1.0))[:, None]], 1))
    packed["sample_normal"] = pack_snorm8x4(np.concatenate([r["sample_normal"], r["sample_position"][:, 3:4]], 1))
    render = np.where(covered[:, None], render, F(0.0))
    return packed, render.astype(F), variance.astype(F), use_variance & covered, covered, stats


@pytest.mark.parametrize("scene,size,emissive,ratio", [("cornell", (72, 72), False, 1.0), ("cornell", (72, 72), True, 1.0), ("minimal", (80, 56), False, 1.0),
                                                       ("cornell", (120, 100), False, 1.5), ("cornell", (128, 96), True, 2.0)])
def test_oracle_spatial_reuse_equals_independent_numpy_restatement(scene, size, emissive, ratio):
    b = Bench(scene, size[0], size[1], taa=plugin.TAA_NONE, upscale_ratio=ratio, temporal_reuse=1, denoise=0, indirect_bounces=2,
              emissive_spatial_reuse=1, indirect_spatial_reuse=1, max_spatial_reuse_count=40)
    orc = b.oracle()
    t_base, s_base, signal, spatial_pass = (2, 4, 1, 3) if emissive else (6, 8, 2, 5)
    totals = dict(merged=0, depth=0, normal=0, back=0, occluded=0, outside=0)
    for f in range(1, 6):
        inp = b.inputs(f)
        head = f % 2
        orc.prepass(inp)
        for p in (0, 1, 2) + (() if emissive else (3, 4)):                 # light_node order up to the pass under test
            orc.run_pass(inp, p)
        temporal = orc.readback(L.OUT_RESERVOIR_0 + t_base + 1 - head).copy()          # reservoir_buffer of this frame
        previous_spatial = orc.readback(L.OUT_RESERVOIR_0 + s_base + head).copy()      # as the temporal passes left it
        orc.run_pass(inp, spatial_pass)
        written = orc.readback(L.OUT_RESERVOIR_0 + s_base + 1 - head).reshape(-1)
        render = orc.readback(L.OUT_RENDER_DIRECT + signal).astype(F).reshape(-1, 4)
        got_var = orc.readback(L.OUT_VARIANCE_DIRECT + signal).reshape(-1)
        if emissive:
            orc.run_pass(inp, 4); orc.run_pass(inp, 5)                                 # finish the frame's light passes
        if f < 3:
            continue
        packed, out, variance, use_variance, covered, stats = spatial_numpy(b, orc, f, emissive, temporal, previous_spatial)
        for k in totals:
            totals[k] += stats[k]
        identical = (written.view(np.uint8).reshape(-1, 64) == packed.view(np.uint8).reshape(-1, 64)).all(1)
        assert identical.mean() >= 0.995, (f, float(identical.mean()))                 # all 64 bytes of the record
        gc, _ = unpack_f16x2(written["reservoir"][:, 0]); wc, _ = unpack_f16x2(packed["reservoir"][:, 0])
        assert (gc == wc).mean() >= 0.999 and wc.max() <= 40
        d = ulps16(render[:, :3], out).max(-1)
        assert (d[covered] == 0).mean() >= 0.99 and (d[covered] <= 1).mean() >= 0.999, (f, float((d[covered] == 0).mean()))
        assert not render[~covered].any()
        if use_variance.any():
            assert np.abs(got_var[use_variance] - variance[use_variance]).max() <= 1e-4
    # every branch of the neighbour loop was taken many times
    print(scene, emissive, totals)
    if scene == "cornell":
        assert totals["merged"] > 3000 and all(totals[k] > 20 for k in ("depth", "normal", "occluded", "outside")), totals
    else:       # a receding ground plane: most neighbours fail the depth-ratio test
        assert totals["merged"] > 500 and totals["depth"] > 5000, totals
