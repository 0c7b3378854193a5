#!/usr/bin/env python
"""TEST INFRASTRUCTURE — the reference's raster prepass (src/prepass.rs + src/shaders/prepass.wgsl) as executable text.

The G-buffer is the one input of the compute path that the reference does not compute with a compute shader: it rasterises the meshes
(prepass.rs:200-260: triangle list, no culling, depth test GreaterEqual on a reverse-Z depth buffer, five colour targets cleared to 0)
through prepass.wgsl's `vertex` and `fragment`.  Product and oracle cast primary rays instead (deviation 1), so the two can
agree only up to the arithmetic of interpolation, and differ in coverage on the pixels a triangle edge crosses.  This module measures
exactly that: prepass.wgsl is translated like the compute shaders (oracle/wgsl/wgsl2cpp.py; `vertex` and `fragment` run as written,
pipeline specialisation TEMPORAL_ANTI_ALIASING / SMAA_TU4X as prepass.rs:190-199 selects it) and driven by a small software rasteriser that
does what the fixed-function stages of a GPU do, with the choices a GPU makes stated:

  * vertices -> clip space by `vertex`; triangles are clipped against the near plane (z <= w: reverse Z) with every varying interpolated
    linearly in clip space, as a GPU's clipper does; the far plane of the infinite projection is never reached, the side planes are
    left to the viewport's scissor (guard band);
  * viewport transform, vertex positions snapped to 1/256 pixel (the sub-pixel precision of every desktop GPU), coverage at pixel centres
    by exact integer edge functions with the top-left rule;
  * varyings interpolated perspective-correctly from the barycentrics of the snapped triangle, in double precision, rounded once to f32;
    `clip_position.z` (the fragment's depth) interpolated linearly in screen space; depth test GreaterEqual, draw order = instance order;
  * `dpdx` / `dpdy`: fine derivatives inside the pixel's 2 x 2 quad (the fragment shader is evaluated at the quad's other pixels with
    extrapolated varyings, as helper invocations are);
  * colour targets in the reference's formats (Rgba32Float, Rgba8Snorm, Rg32Float, Rg32Float, Rgba32Float).

bevy_pbr's `Mesh` uniform and the two mesh functions prepass.wgsl imports come from oracle/wgsl/prelude/ (restated, like the other bevy_pbr
imports).  Only in the build container (needs /root/reference and g++)."""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference as R  # noqa: E402
import wgsl2cpp  # noqa: E402

HARNESS = r"""
// ---------------------------------------------------------------------------------------------------------------- raster harness
// (oracle/wgsl/raster_prepass.py) — appended to the translation of prepass.wgsl; `vertex` / `fragment` above are the reference's text.
namespace wgsl {
struct RasterTargets { f32* position; u32* normal; f32* depth_gradient; f32* instance_material; f32* velocity_uv; f32* depth; int w, h; };
static RasterTargets rt;
static long long raster_skipped = 0, raster_fragments = 0;

static VertexOutput interpolate(const VertexOutput v[3], const double lam[3], double sx, double sy) {
    // perspective-correct: b_i = (lam_i / w_i) / sum_j (lam_j / w_j); clip_position = (pixel centre, linear NDC depth, 1 / w)
    double iw[3], s = 0.0;
    for (int i = 0; i < 3; ++i) { iw[i] = lam[i] / (double)v[i].clip_position.w; s += iw[i]; }
    double b[3] = {iw[0] / s, iw[1] / s, iw[2] / s};
    VertexOutput o{};
    auto mix4 = [&](const vec4<f32>& a0, const vec4<f32>& a1, const vec4<f32>& a2) {
        vec4<f32> r; for (int k = 0; k < 4; ++k) r.v[k] = (f32)(b[0] * a0.v[k] + b[1] * a1.v[k] + b[2] * a2.v[k]); return r; };
    o.world_position = mix4(v[0].world_position, v[1].world_position, v[2].world_position);
    o.previous_world_position = mix4(v[0].previous_world_position, v[1].previous_world_position, v[2].previous_world_position);
    for (int k = 0; k < 3; ++k) o.world_normal.v[k] = (f32)(b[0] * v[0].world_normal.v[k] + b[1] * v[1].world_normal.v[k] + b[2] * v[2].world_normal.v[k]);
    for (int k = 0; k < 2; ++k) o.uv.v[k] = (f32)(b[0] * v[0].uv.v[k] + b[1] * v[1].uv.v[k] + b[2] * v[2].uv.v[k]);
    double z = 0.0;
    for (int i = 0; i < 3; ++i) z += lam[i] * ((double)v[i].clip_position.z / (double)v[i].clip_position.w);
    o.clip_position = vec4<f32>((f32)sx, (f32)sy, (f32)z, (f32)s);
    return o;
}

extern "C" {
void raster_bind_targets(f32* position, u32* normal, f32* depth_gradient, f32* instance_material, f32* velocity_uv, f32* depth, int w, int h) {
    rt = RasterTargets{position, normal, depth_gradient, instance_material, velocity_uv, depth, w, h};
}
long long raster_skipped_triangles() { return raster_skipped; }
long long raster_fragment_count() { return raster_fragments; }
void raster_reset_counters() { raster_skipped = raster_fragments = 0; }

static void raster_triangle(VertexOutput v[3]);
static VertexOutput clip_lerp(const VertexOutput& a, const VertexOutput& b, double t) {      // linear in clip space
    VertexOutput o{};
    auto L = [&](f32 x, f32 y) { return (f32)((double)x + ((double)y - (double)x) * t); };
    for (int k = 0; k < 4; ++k) { o.clip_position.v[k] = L(a.clip_position.v[k], b.clip_position.v[k]); o.world_position.v[k] = L(a.world_position.v[k], b.world_position.v[k]);
                                  o.previous_world_position.v[k] = L(a.previous_world_position.v[k], b.previous_world_position.v[k]); }
    for (int k = 0; k < 3; ++k) o.world_normal.v[k] = L(a.world_normal.v[k], b.world_normal.v[k]);
    for (int k = 0; k < 2; ++k) o.uv.v[k] = L(a.uv.v[k], b.uv.v[k]);
    return o;
}
// one draw: a triangle list of the bound mesh (uniforms `mesh`, `previous_mesh`, `instance_index` bound by the caller)
void raster_draw(const f32* positions, const f32* normals, const f32* uvs, const u32* indices, size_t triangle_count) {
    for (size_t t = 0; t < triangle_count; ++t) {
        VertexOutput v[3];
        for (int k = 0; k < 3; ++k) {
            const u32 i = indices[3 * t + k];
            Vertex in{};
            in.position = vec3<f32>(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
            in.normal = vec3<f32>(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
            in.uv = vec2<f32>(uvs[2 * i], uvs[2 * i + 1]);
            v[k] = vertex(in);
        }
        // near plane (reverse Z: 0 <= z <= w): Sutherland-Hodgman against d = w - z >= 0
        VertexOutput poly[4]; int n = 0;
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = (double)v[k].clip_position.w - (double)v[k].clip_position.z;
        for (int k = 0; k < 3; ++k) {
            const int m = (k + 1) % 3;
            if (d[k] >= 0.0) poly[n++] = v[k];
            if ((d[k] >= 0.0) != (d[m] >= 0.0)) poly[n++] = clip_lerp(v[k], v[m], d[k] / (d[k] - d[m]));
        }
        if (n < 3) { raster_skipped += 1; continue; }
        for (int k = 1; k + 1 < n; ++k) { VertexOutput tri[3] = {poly[0], poly[k], poly[k + 1]}; raster_triangle(tri); }
    }
}
}  // extern "C"
static void raster_triangle(VertexOutput v[3]) {
    const int W = rt.w, H = rt.h;
    {
        for (int k = 0; k < 3; ++k) if (!(v[k].clip_position.w > 0.0f)) { raster_skipped += 1; return; }
        // viewport transform (y down), snapped to 1/256 pixel
        long long X[3], Y[3];
        for (int k = 0; k < 3; ++k) {
            const double nx = (double)v[k].clip_position.x / (double)v[k].clip_position.w, ny = (double)v[k].clip_position.y / (double)v[k].clip_position.w;
            X[k] = llround((nx * 0.5 + 0.5) * W * 256.0);
            Y[k] = llround((0.5 - ny * 0.5) * H * 256.0);
        }
        long long area = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
        if (area == 0) return;
        if (area < 0) { std::swap(X[1], X[2]); std::swap(Y[1], Y[2]); std::swap(v[1], v[2]); area = -area; }   // cull_mode: None — both windings are drawn
        long long minx = std::min(X[0], std::min(X[1], X[2])), maxx = std::max(X[0], std::max(X[1], X[2]));
        long long miny = std::min(Y[0], std::min(Y[1], Y[2])), maxy = std::max(Y[0], std::max(Y[1], Y[2]));
        int px0 = (int)std::max(0LL, (minx - 128) / 256), px1 = (int)std::min((long long)W - 1, (maxx + 128) / 256);
        int py0 = (int)std::max(0LL, (miny - 128) / 256), py1 = (int)std::min((long long)H - 1, (maxy + 128) / 256);
        // edge i is opposite vertex i; with the triangle wound positively in this y-down frame the interior has e_i >= 0
        auto edge = [&](int a, int b, long long x, long long y) { return (X[b] - X[a]) * (y - Y[a]) - (Y[b] - Y[a]) * (x - X[a]); };
        auto top_left = [&](int a, int b) {      // D3D / Vulkan fill rule for the orientation used here (clockwise on screen, y down)
            const long long dx = X[b] - X[a], dy = Y[b] - Y[a];
            return (dy == 0 && dx > 0) || dy < 0;
        };
        // orientation: make the winding clockwise on screen (y down) so that top-left is as defined above
        // (area > 0 with the formula above in a y-down frame IS clockwise on screen)
        auto lambda_at = [&](int px, int py, double lam[3]) {
            const long long x = px * 256LL + 128, y = py * 256LL + 128;
            const long long e0 = edge(1, 2, x, y), e1 = edge(2, 0, x, y), e2 = edge(0, 1, x, y);
            lam[0] = (double)e0 / (double)area; lam[1] = (double)e1 / (double)area; lam[2] = (double)e2 / (double)area;
        };
        for (int py = py0; py <= py1; ++py)
            for (int px = px0; px <= px1; ++px) {
                const long long x = px * 256LL + 128, y = py * 256LL + 128;
                const long long e[3] = {edge(1, 2, x, y), edge(2, 0, x, y), edge(0, 1, x, y)};
                const int ea[3] = {1, 2, 0}, eb[3] = {2, 0, 1};
                bool inside = true;
                for (int k = 0; k < 3; ++k) inside = inside && (e[k] > 0 || (e[k] == 0 && top_left(ea[k], eb[k])));
                if (!inside) continue;
                double lam[3];
                lambda_at(px, py, lam);
                const VertexOutput here = interpolate(v, lam, px + 0.5, py + 0.5);
                const size_t idx = (size_t)py * W + px;
                if (!(here.clip_position.z >= rt.depth[idx])) continue;                 // CompareFunction::GreaterEqual, reverse Z
                if (!(here.clip_position.z >= 0.0f && here.clip_position.z <= 1.0f)) continue;   // depth clipping of the viewport
                // fine derivatives: the fragment shader at the other pixels of the 2 x 2 quad (helper invocations, extrapolated varyings)
                const int qx = px & ~1, qy = py & ~1;
                VertexOutput quad[2][2];
                for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) { double l2[3]; lambda_at(qx + i, qy + j, l2); quad[j][i] = interpolate(v, l2, qx + i + 0.5, qy + j + 0.5); }
                wgsl_derivatives& D = wgsl_derivative_state();
                D.mode = 1; D.slot = 0; D.n = 0; fragment(quad[py & 1][0]);      // left pixel of this pixel's quad row
                D.mode = 1; D.slot = 1; D.n = 0; fragment(quad[py & 1][1]);      // right pixel
                D.mode = 1; D.slot = 2; D.n = 0; fragment(quad[0][px & 1]);      // top pixel of this pixel's quad column
                D.mode = 1; D.slot = 3; D.n = 0; fragment(quad[1][px & 1]);      // bottom pixel
                D.mode = 2; D.nx = D.ny = 0;
                const FragmentOutput out = fragment(here);
                D.mode = 0;
                rt.depth[idx] = here.clip_position.z;
                for (int k = 0; k < 4; ++k) rt.position[4 * idx + k] = out.position.v[k];
                rt.normal[idx] = pack4x8snorm(out.normal);
                rt.depth_gradient[2 * idx] = out.depth_gradient.x; rt.depth_gradient[2 * idx + 1] = out.depth_gradient.y;
                rt.instance_material[2 * idx] = out.instance_material.x; rt.instance_material[2 * idx + 1] = out.instance_material.y;
                for (int k = 0; k < 4; ++k) rt.velocity_uv[4 * idx + k] = out.velocity_uv.v[k];
                raster_fragments += 1;
            }
    }
}
}  // namespace wgsl
"""


def build(defs):
    os.makedirs(R.OUT, exist_ok=True)
    src = os.path.join(wgsl2cpp.REF_SHADERS, "prepass.wgsl")
    cpp_text = wgsl2cpp.translate(src, defs) + HARNESS
    rt = open(os.path.join(HERE, "wgsl_rt.h")).read()
    tag = hashlib.sha1((cpp_text + rt + " ".join(R.CXXFLAGS)).encode()).hexdigest()[:12]
    name = "prepass" + ("_" + "_".join(sorted(defs)).lower() if defs else "")
    so = os.path.join(R.OUT, f"{name}_{tag}.so")
    if not os.path.exists(so):
        cpp = os.path.join(R.OUT, name + ".cpp")
        open(cpp, "w").write(cpp_text)
        r = subprocess.run(["g++"] + R.CXXFLAGS + [cpp, "-o", so], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on the translation of prepass.wgsl {defs}:\n{r.stderr[-3000:]}")
    return C.CDLL(so, mode=C.RTLD_LOCAL)


class RasterPrepass:
    """PrepassNode::run (prepass.rs:760-852) for one camera: all instances drawn into cleared targets"""

    def __init__(self, taa=False, smaa=False):
        defs = (["TEMPORAL_ANTI_ALIASING"] if taa else []) + (["SMAA_TU4X"] if smaa else [])       # prepass.rs:190-199
        self.lib = build(defs)
        self.lib.raster_skipped_triangles.restype = self.lib.raster_fragment_count.restype = C.c_longlong

    def _bind(self, name, raw):
        a = np.ascontiguousarray(raw)
        fn = getattr(self.lib, "bind_" + name)
        fn.argtypes = [C.c_void_p, C.c_size_t]
        fn(a.ctypes.data, a.nbytes)

    def render(self, inputs, width, height, meshes, inst_mesh, instances, previous_models=None):
        """meshes: list of (positions, normals, uvs, indices); inst_mesh: mesh id per instance; instances: the instance records
        (model, inverse_transpose_model, material); previous_models: (n, 16) or None (static).  Returns the five planes + counters."""
        self._bind("frame", np.frombuffer(bytes(inputs.frame), np.uint8))
        self._bind("view", np.frombuffer(bytes(inputs.view), np.uint8))
        self._bind("previous_view", np.frombuffer(bytes(inputs.previous_view), np.uint8))
        n = width * height
        position = np.zeros((height, width, 4), np.float32)
        normal = np.zeros((height, width), np.uint32)
        depth_gradient = np.zeros((height, width, 2), np.float32)
        instance_material = np.zeros((height, width, 2), np.float32)
        velocity_uv = np.zeros((height, width, 4), np.float32)
        depth = np.zeros((height, width), np.float32)                    # Camera3d::depth_load_op default: Clear(0.0), reverse Z
        self.lib.raster_bind_targets.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int]
        self.lib.raster_bind_targets(position.ctypes.data, normal.ctypes.data, depth_gradient.ctypes.data, instance_material.ctypes.data,
                                     velocity_uv.ctypes.data, depth.ctypes.data, width, height)
        self.lib.raster_reset_counters()
        self.lib.raster_draw.argtypes = [C.c_void_p] * 4 + [C.c_size_t]
        for i, rec in enumerate(instances):
            model = np.ascontiguousarray(rec["model"], np.float32).reshape(16)
            mesh = np.zeros(36, np.float32)                                          # Mesh: model, inverse_transpose_model, flags (+ padding)
            mesh[:16] = model
            mesh[16:32] = np.ascontiguousarray(rec["inverse_transpose_model"], np.float32).reshape(16)
            self._bind("mesh", mesh.view(np.uint8))
            prev = np.zeros(32, np.float32)
            prev[:16] = model if previous_models is None else np.asarray(previous_models[i], np.float32).reshape(16)
            self._bind("previous_mesh", prev.view(np.uint8))
            self._bind("instance_index", np.array([i, int(rec["material"])], np.uint32).view(np.uint8))
            pos, nrm, uv, idx = meshes[inst_mesh[i]]
            pos = np.ascontiguousarray(pos, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
            uv = np.ascontiguousarray(uv, np.float32); idx = np.ascontiguousarray(idx, np.uint32).reshape(-1)
            self.lib.raster_draw(pos.ctypes.data, nrm.ctypes.data, uv.ctypes.data, idx.ctypes.data, len(idx) // 3)
        return {"position": position, "normal": normal, "depth_gradient": depth_gradient, "instance_material": instance_material,
                "velocity_uv": velocity_uv, "depth": depth, "skipped_triangles": int(self.lib.raster_skipped_triangles()),
                "fragments": int(self.lib.raster_fragment_count())}
