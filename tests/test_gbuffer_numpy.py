"""The G-buffer and `full_screen_albedo` (P1), pinned from the outside.  The reference rasterises
(prepass.wgsl:40-100); this back end and its oracle cast one primary ray per pixel (deviation 1).  What every plane
must CONTAIN is the reference's: world position + NDC depth, the interpolated per-vertex-normalised world normal as Rgba8Snorm,
dpdx / dpdy of NDC depth, (instance + 0.5, material + 0.5), the motion vector clip_to_uv(view_proj * p) -
clip_to_uv(previous_view_proj * p_previous) and the interpolated uv.  A SECOND, independent computation in numpy float64:
pixel-centre rays from `inverse_view_proj`, closest hit by brute force over every world triangle, barycentric interpolation,
and the depth gradient from the affine fit of NDC depth through the hit triangle's three projected vertices (no finite
differences, no second ray) — compared plane by plane with the oracle under a translating camera; then `env_brdf`
(light.wgsl:891-908, bevy_pbr EnvBRDFApprox) on those planes against the albedo plane.  CPU only."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench
from tests.test_direct_lit_numpy import F, dot, normalize, ulps16
from tests.test_indirect_numpy import Scene, env_brdf_approx


def mat(a):                                              # 16 floats, column-major -> (4, 4) with M @ column-vector semantics
    return np.array(list(a), np.float64).reshape(4, 4).T


def project(M, p):
    hp = np.concatenate([p, np.ones(p.shape[:-1] + (1,))], -1) @ M.T
    return hp


def clip_to_uv(clip):                                    # utils.wgsl clip_to_uv
    uv = clip[..., :2] / clip[..., 3:4]
    uv = (uv + 1.0) * 0.5
    uv[..., 1] = 1.0 - uv[..., 1]
    return uv


@pytest.mark.parametrize("scene,size", [("cornell", (80, 64)), ("simple", (96, 54)), ("soup5", (80, 56))])
def test_oracle_gbuffer_equals_independent_numpy_computation(scene, size):
    if scene.startswith("soup"):
        from bevy_hikari_b200 import scenes
        scenes.SCENE_BUILDERS[scene] = lambda: scenes.soup(int(scene[4:]))
    W, H = size
    b = Bench(scene, W, H, taa=plugin.TAA_NONE, upscale_ratio=1.0, denoise=0, indirect_bounces=1)
    orc = b.oracle()
    sc = Scene(b)
    for f in (1, 2):
        inp = b.moving_inputs(f, step=(0.05, 0.02, -0.03))
        orc.prepass(inp)
        orc.run_pass(inp, 0)                             # full_screen_albedo
    VP, IVP = mat(inp.view.view_proj), mat(inp.view.inverse_view_proj)
    PVP = mat(inp.previous_view.view_proj)
    eye = np.array(list(inp.view.world_position), np.float64)
    ys, xs = [a.reshape(-1) for a in np.meshgrid(np.arange(H), np.arange(W), indexing="ij")]
    ndc = np.stack([2.0 * (xs + 0.5) / W - 1.0, 1.0 - 2.0 * (ys + 0.5) / H, np.full(len(xs), 0.5), np.ones(len(xs))], -1)
    far = ndc @ IVP.T
    far = far[:, :3] / far[:, 3:4]
    direction = far - eye
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    t, inst, tri, u, v, graze = sc.closest(np.tile(eye, (len(xs), 1)), direction)
    hit = np.isfinite(t)
    pos = orc.readback(L.OUT_GBUFFER_POSITION).reshape(-1, 4)
    nrm = orc.readback(L.OUT_GBUFFER_NORMAL).reshape(-1, 4)
    grad = orc.readback(L.OUT_GBUFFER_DEPTH_GRADIENT).reshape(-1, 2)
    im = orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL).reshape(-1, 2)
    vu = orc.readback(L.OUT_GBUFFER_VELOCITY_UV).reshape(-1, 4)
    clean = hit & ~graze
    # coverage and ids
    assert ((pos[:, 3] > 0) == hit)[~graze].all()
    assert not pos[~hit & ~graze].any() and not im[~hit & ~graze].any() and not nrm[~hit & ~graze].any()
    assert (np.floor(im[clean, 0]) == inst[clean]).mean() >= 0.999
    mat_of = np.array([int(i["material"]) for i in sc.inst])
    same_inst = clean & (np.floor(im[:, 0]) == inst)
    assert (im[same_inst, 1] == mat_of[inst[same_inst]] + 0.5).all() and (im[same_inst, 0] == inst[same_inst] + 0.5).all()
    k = np.nonzero(same_inst)[0]
    # position and NDC depth (prepass.wgsl:79: world_position.xyz, clip_position.z)
    world = eye + direction[k] * t[k, None]
    # fp32 hit distance in object space vs float64 in world space: the error scales with the distance along the ray
    assert (np.abs(pos[k, :3] - world).max(1) <= 1e-4 * t[k] + 1e-5).all()
    clip = project(VP, world)
    depth = clip[:, 2] / clip[:, 3]
    assert np.allclose(pos[k, 3], depth, rtol=2e-4)
    # normal (:60, :80): per-vertex normalised world normals, interpolated, not renormalised, stored as snorm8
    want_n = np.zeros((len(k), 3))
    want_uv = np.zeros((len(k), 2))
    dgrad = np.zeros((len(k), 2))
    for i in np.unique(inst[k]):
        m = inst[k] == i
        one = sc.inst[i]
        verts = sc.bufs["vertices"][int(one["mesh"]["vertex"]) + sc.vidx[i][tri[k][m]].astype(np.int64)]
        itm = one["inverse_transpose_model"].reshape(4, 4)[:3, :3].astype(np.float64)
        wn = verts["normal"].astype(np.float64) @ itm                        # n.x * col0 + n.y * col1 + n.z * col2, per vertex
        wn /= np.linalg.norm(wn, axis=2, keepdims=True)
        uu, vv = u[k][m, None].astype(np.float64), v[k][m, None].astype(np.float64)
        want_n[m] = wn[:, 0] + uu * (wn[:, 1] - wn[:, 0]) + vv * (wn[:, 2] - wn[:, 0])
        tuv = np.stack([verts["u"], verts["v"]], -1).astype(np.float64)
        want_uv[m] = tuv[:, 0] + uu * (tuv[:, 1] - tuv[:, 0]) + vv * (tuv[:, 2] - tuv[:, 0])
        # depth gradient: NDC depth is affine in NDC x, y on a planar triangle -> fit through its three projected vertices
        P = sc.tris[i][tri[k][m]]                                            # (n, 3 vertices, xyz) world
        c3 = project(VP, P)
        n3 = c3[..., :3] / c3[..., 3:4]
        A = np.concatenate([n3[..., :2], np.ones(n3.shape[:2] + (1,))], -1)  # rows [x y 1]
        with np.errstate(all="ignore"):
            coef = np.linalg.solve(A, n3[..., 2:3])[..., 0]                  # z = a x + b y + c
        dgrad[m] = np.stack([coef[:, 0] * 2.0 / W, -coef[:, 1] * 2.0 / H], -1)
    got_n = np.maximum(nrm[k].astype(np.float64) / 127.0, -1.0)[:, :3]
    assert (np.abs(got_n * 127.0 - np.clip(want_n, -1, 1) * 127.0) <= 0.51).mean() >= 0.999        # the same snorm8 code, +- rounding at .5
    assert (nrm[k, 3] == 127).all()
    assert np.allclose(vu[k, 2:], want_uv, atol=3e-4) and (np.abs(vu[k, 2:] - want_uv).max(1) <= 2e-5).mean() >= 0.75      # fp32 barycentrics of far, large triangles
    well = np.abs(dgrad).max(1) < 1e3                                        # edge-on triangles: the fit is ill-conditioned
    scale = np.abs(dgrad[well]).max(1, keepdims=True) + 1e-7
    assert (np.abs(grad[k][well] - dgrad[well]) <= 2e-3 * scale + 1e-7).mean() >= 0.995
    # motion vector (:94-95): static geometry, so the previous world position is the same point under the previous view
    velocity = clip_to_uv(clip) - clip_to_uv(project(PVP, world))
    assert np.abs(velocity).max() > 1e-3 and np.allclose(vu[k, :2], velocity, atol=3e-6)
    # full_screen_albedo (light.wgsl:1019-1042) from the oracle's own planes: env_brdf with the raw G-buffer normal
    albedo = orc.readback(L.OUT_ALBEDO).astype(F).reshape(-1, 4)
    mats = sc.bufs["materials"][np.floor(im[:, 1]).astype(np.int64)]
    untextured = (mats["base_color_texture"] == 0xFFFFFFFF) & (mats["metallic_roughness_texture"] == 0xFFFFFFFF)
    sel = np.nonzero((pos[:, 3] >= F(1.1920929e-7)) & untextured)[0]
    m = mats[sel]
    Vd = normalize(np.array(list(inp.view.world_position), F) - pos[sel, :3])
    Nn = np.maximum(nrm[sel].astype(F) / F(127.0), F(-1.0))[:, :3]
    base = m["base_color"][:, :3]
    metallic, reflectance = m["metallic"][:, None], m["reflectance"][:, None]
    rough = np.clip(m["perceptual_roughness"], F(0.089), F(1.0)); rough = rough * rough
    F0 = F(0.16) * reflectance * reflectance * (F(1.0) - metallic) + base * metallic
    NoV = np.fmax(dot(Nn, Vd), F(0.0001))
    want_albedo = env_brdf_approx(base * (F(1.0) - metallic), np.ones_like(rough), NoV) + env_brdf_approx(F0, rough, NoV)
    d = ulps16(albedo[sel, :3], want_albedo).max(-1)
    assert len(sel) > 500 and (d == 0).mean() >= 0.99 and d.max() <= 1, (len(sel), float((d == 0).mean()), int(d.max()))
    assert (albedo[sel, 3] == 1).all() and not albedo[pos[:, 3] < F(1.1920929e-7)].any()


# ------------------------------------------------------------------------------------------- textured surfaces
def decode_texture(t):
    """RGBA8 -> float32 texels: sRGB transfer on rgb when the image is an sRGB one (base colour / emissive slots), alpha linear"""
    a = t["rgba"].astype(np.float64) / 255.0
    if t["srgb"]:
        rgb = a[..., :3]
        a[..., :3] = np.where(rgb <= 0.04045, rgb / 12.92, ((rgb + 0.055) / 1.055) ** 2.4)
    return a.astype(F)


def wrap(i, n, mode):                                    # address modes: 0 repeat, 1 clamp-to-edge, 2 mirror-repeat
    if mode == 0:
        return np.mod(i, n)
    if mode == 1:
        return np.clip(i, 0, n - 1)
    j = np.mod(i, 2 * n)
    return np.where(j < n, j, 2 * n - 1 - j)


def sample(t, texels, u, v):
    """textureSampleLevel(textures[id], samplers[id], uv, 0.0) with fp32 bilinear weights (deviation 4)"""
    h, w = texels.shape[:2]
    mu, mv = t["address_mode_u"], t["address_mode_v"]
    if not t["filter_linear"]:
        return texels[wrap(np.floor(v * F(h)).astype(np.int64), h, mv), wrap(np.floor(u * F(w)).astype(np.int64), w, mu)]
    fx, fy = u * F(w) - F(0.5), v * F(h) - F(0.5)
    x0, y0 = np.floor(fx), np.floor(fy)
    ax, ay = (fx - x0)[:, None], (fy - y0)[:, None]
    x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
    xa, xb, ya, yb = wrap(x0, w, mu), wrap(x0 + 1, w, mu), wrap(y0, h, mv), wrap(y0 + 1, h, mv)
    top = texels[ya, xa] * (F(1) - ax) + texels[ya, xb] * ax
    bot = texels[yb, xa] * (F(1) - ax) + texels[yb, xb] * ax
    return top * (F(1) - ay) + bot * ay


@pytest.mark.parametrize("scene,size", [("samplers", (128, 80)), ("city", (128, 72))])
def test_oracle_textured_surfaces_equal_independent_numpy_sampler(scene, size):
    """retreive_surface with textures (light.wgsl:749-781) through `full_screen_albedo`: every sampler state the texture path
    distinguishes (repeat / clamp / mirror on either axis, nearest and linear, sRGB and linear data, uvs from -1 to 2) in
    scenes.samplers, and the glTF textures of the city houses"""
    W, H = size
    b = Bench(scene, W, H, taa=plugin.TAA_NONE, upscale_ratio=1.0, denoise=0, indirect_bounces=1)
    orc = b.oracle()
    inp = b.inputs(1)
    orc.prepass(inp)
    orc.run_pass(inp, 0)
    pos = orc.readback(L.OUT_GBUFFER_POSITION).reshape(-1, 4)
    nrm = orc.readback(L.OUT_GBUFFER_NORMAL).reshape(-1, 4)
    im = orc.readback(L.OUT_GBUFFER_INSTANCE_MATERIAL).reshape(-1, 2)
    vu = orc.readback(L.OUT_GBUFFER_VELOCITY_UV).reshape(-1, 4)
    albedo = orc.readback(L.OUT_ALBEDO).astype(F).reshape(-1, 4)
    sel = np.nonzero(pos[:, 3] >= F(1.1920929e-7))[0]
    mats = b.world.buffers()["materials"][np.floor(im[sel, 1]).astype(np.int64)]
    textures = [dict(t, texels=decode_texture(t)) for t in b.scene.textures]
    u, v = vu[sel, 2], vu[sel, 3]
    base = mats["base_color"].copy()
    metallic = mats["metallic"].copy()
    occlusion = np.ones(len(sel), F)
    used = set()
    for slot, apply in (("base_color_texture", "base"), ("metallic_roughness_texture", "metallic"), ("occlusion_texture", "occlusion")):
        ids = mats[slot]
        for tid in np.unique(ids[ids != 0xFFFFFFFF]):
            m = ids == tid
            t = textures[int(tid)]
            s = sample(t, t["texels"], u[m], v[m])
            used.add((t["address_mode_u"], t["address_mode_v"], t["filter_linear"], t["srgb"]))
            if apply == "base":
                base[m] = base[m] * s
            elif apply == "metallic":
                metallic[m] = metallic[m] * s[:, 0]
            else:
                occlusion[m] = s[:, 0]
    Vd = normalize(np.array(list(inp.view.world_position), F) - pos[sel, :3])
    Nn = np.maximum(nrm[sel].astype(F) / F(127.0), F(-1.0))[:, :3]
    reflectance = mats["reflectance"][:, None]
    rough = np.clip(mats["perceptual_roughness"], F(0.089), F(1.0)); rough = rough * rough
    F0 = F(0.16) * reflectance * reflectance * (F(1.0) - metallic[:, None]) + base[:, :3] * metallic[:, None]
    NoV = np.fmax(dot(Nn, Vd), F(0.0001))
    want = (env_brdf_approx(base[:, :3] * (F(1.0) - metallic[:, None]), np.ones_like(rough), NoV) + env_brdf_approx(F0, rough, NoV)) * occlusion[:, None]
    d = ulps16(albedo[sel, :3], want).max(-1)
    textured = mats["base_color_texture"] != 0xFFFFFFFF
    assert textured.sum() > 500 and (d[textured] == 0).mean() >= 0.99 and d.max() <= 1, (int(textured.sum()), float((d[textured] == 0).mean()), int(d.max()))
    if scene == "samplers":     # all three address modes, both filters, both encodings were sampled
        assert {m for mu, mv, _, _ in used for m in (mu, mv)} == {0, 1, 2} and {f for _, _, f, _ in used} == {0, 1} and {s for *_, s in used} == {0, 1}, used
