// kernels_upscale.cu — the temporal upscalers that follow tone mapping in the default HikariSettings pipeline
// (SURVEY.md 8(f) rank 1): smaa_tu4x + smaa_tu4x_extrapolate (src/shaders/smaa.wgsl:81-271) and taa_jasmine
// (src/shaders/taa.wgsl:79-170), dispatched by PostProcessNode::run (src/post_process.rs:1236-1277), and FSR 1.0 EASU + RCAS
// (Upscale::Fsr1, post_process.rs:1279-1308; rank 4).  Texture addressing (both samplers clamp-to-edge, post_process.rs:697-708):
//   nearest: texel floor(uv * size);  linear: bilinear around uv * size - 0.5 with fp32 weights;
//   textureGather: the 2x2 bilinear footprint as (x: (i0,j1), y: (i1,j1), z: (i1,j0), w: (i0,j0)).
#include "hk_device.cuh"
#include "hk_kernels.h"

namespace hkd {

// An image of w x h texels of which this context stores the window that starts at texel (x0, y0) with row pitch `pitch`
// (the whole image for a full-frame context; the tile's allocation — owned rectangle + ghost ring — for a tile).
// Addressing clamps to the IMAGE edge, as the sampler does; the caller guarantees that what it samples lies in the window.
struct Image16 {   // Rgba16Float
    const uint2* t; int w, h, x0, y0, ww, wh;      // image size; window origin and size (row pitch = ww)
    __device__ __forceinline__ vec4 fetch(int x, int y) const {
        // memory safety on tiles: a sample that leaves the window (motion beyond the margin) reads the window's edge
        x = min(max(x, x0), x0 + ww - 1); y = min(max(y, y0), y0 + wh - 1);
        uint2 u = t[(size_t)(y - y0) * (size_t)ww + (size_t)(x - x0)];
        uvec2 q; q.x = u.x; q.y = u.y;
        return unpack_rgba16f(q);
    }
    __device__ __forceinline__ vec4 texel(int x, int y) const {
        x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1);
        return fetch(x, y);
    }
    __device__ __forceinline__ vec4 load(int x, int y) const {   // textureLoad: zero outside
        if (x < 0 || y < 0 || x >= w || y >= h) return v4(0.0f);
        return fetch(x, y);
    }
    __device__ __forceinline__ vec4 nearest(vec2 uv) const { return texel((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    __device__ __forceinline__ vec4 linear(vec2 uv) const {
        float fx = uv.x * (float)w - 0.5f, fy = uv.y * (float)h - 0.5f;
        float x0 = floorf(fx), y0 = floorf(fy), ax = fx - x0, ay = fy - y0;
        vec4 t00 = texel((int)x0, (int)y0), t10 = texel((int)x0 + 1, (int)y0), t01 = texel((int)x0, (int)y0 + 1), t11 = texel((int)x0 + 1, (int)y0 + 1);
        vec4 top = t00 * (1.0f - ax) + t10 * ax, bot = t01 * (1.0f - ax) + t11 * ax;
        return top * (1.0f - ay) + bot * ay;
    }
    __device__ __forceinline__ void gather(vec2 uv, vec4 out[4]) const {
        float fx = uv.x * (float)w - 0.5f, fy = uv.y * (float)h - 0.5f;
        int i0 = (int)floorf(fx), j0 = (int)floorf(fy);
        out[0] = texel(i0, j0 + 1); out[1] = texel(i0 + 1, j0 + 1); out[2] = texel(i0 + 1, j0); out[3] = texel(i0, j0);
    }
};
struct Image32 {   // a float4 plane of the G-buffer (same windowing)
    const float4* t; int w, h, x0, y0, ww, wh;
    __device__ __forceinline__ vec4 texel(int x, int y) const {
        x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1);
        x = min(max(x, x0), x0 + ww - 1); y = min(max(y, y0), y0 + wh - 1);
        return f4v(t[(size_t)(y - y0) * (size_t)ww + (size_t)(x - x0)]);
    }
    __device__ __forceinline__ vec4 nearest(vec2 uv) const { return texel((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    __device__ __forceinline__ vec4 gather_w(vec2 uv) const {
        float fx = uv.x * (float)w - 0.5f, fy = uv.y * (float)h - 0.5f;
        int i0 = (int)floorf(fx), j0 = (int)floorf(fy);
        return v4(texel(i0, j0 + 1).w, texel(i0 + 1, j0 + 1).w, texel(i0 + 1, j0).w, texel(i0, j0).w);
    }
};

__device__ __forceinline__ vec3 RGB_to_YCoCg(vec3 rgb) {  // smaa.wgsl:23-28
    float y = (rgb.x / 4.0f) + (rgb.y / 2.0f) + (rgb.z / 4.0f);
    float co = (rgb.x / 2.0f) - (rgb.z / 2.0f);
    float cg = (-rgb.x / 4.0f) + (rgb.y / 2.0f) - (rgb.z / 4.0f);
    return v3(y, co, cg);
}
__device__ __forceinline__ vec3 clamp01(vec3 v) { return v3(clampf(v.x, 0.0f, 1.0f), clampf(v.y, 0.0f, 1.0f), clampf(v.z, 0.0f, 1.0f)); }
__device__ __forceinline__ vec3 YCoCg_to_RGB(vec3 c) { return clamp01(v3(c.x + c.y - c.z, c.x + c.z, c.x - c.y - c.z)); }  // :30-35
__device__ __forceinline__ vec3 vabs(vec3 a) { return v3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
__device__ __forceinline__ vec3 vsqrt(vec3 a) { return v3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
__device__ __forceinline__ vec3 clip_towards_aabb_center(vec3 previous_color, vec3 aabb_min, vec3 aabb_max) {  // smaa.wgsl:37-45
    vec3 p_clip = 0.5f * (aabb_max + aabb_min);
    vec3 e_clip = 0.5f * (aabb_max - aabb_min);
    vec3 v_clip = previous_color - p_clip;
    vec3 v_unit = v_clip / e_clip;
    vec3 a_unit = vabs(v_unit);
    float ma_unit = fmax_(a_unit.x, fmax_(a_unit.y, a_unit.z));
    return (ma_unit > 1.0f) ? p_clip + v_clip / ma_unit : previous_color;
}
// smaa.wgsl:52-72 / taa.wgsl:57-77
__device__ __forceinline__ vec2 nearest_velocity(const Image32& position, const Image32& velocity_uv, vec2 uv, vec2 texel_size) {
    vec4 depths;
    depths.x = position.nearest(uv + v2(texel_size.x, texel_size.y)).w;
    depths.y = position.nearest(uv + v2(-texel_size.x, texel_size.y)).w;
    depths.z = position.nearest(uv + v2(texel_size.x, -texel_size.y)).w;
    depths.w = position.nearest(uv + v2(-texel_size.x, -texel_size.y)).w;
    float max_depth = fmax_(fmax_(depths.x, depths.y), fmax_(depths.z, depths.w));
    float depth = position.nearest(uv).w;
    vec2 offset = v2(0.0f, 0.0f);
    if (depth < max_depth) {
        vec4 sx = v4(depths.x == max_depth ? 1.0f : 0.0f, depths.y == max_depth ? -1.0f : 0.0f, depths.z == max_depth ? 1.0f : 0.0f,
                     depths.w == max_depth ? -1.0f : 0.0f);
        vec4 sy = v4(depths.x == max_depth ? 1.0f : 0.0f, depths.y == max_depth ? 1.0f : 0.0f, depths.z == max_depth ? -1.0f : 0.0f,
                     depths.w == max_depth ? -1.0f : 0.0f);
        offset = v2(dot(v4(texel_size.x), sx), dot(v4(texel_size.y), sy));
    }
    vec4 v = velocity_uv.nearest(uv + offset);
    return v2(v.x, v.y);
}
__device__ __forceinline__ float distance4(vec4 a, vec4 b) { vec4 d = a - b; return sqrtf(dot(d, d)); }
__device__ __forceinline__ float distance2(vec2 a, vec2 b) { vec2 d = a - b; return sqrtf(dot(d, d)); }
__device__ __forceinline__ bool any_lt(vec4 a, float t) { return a.x < t || a.y < t || a.z < t || a.w < t; }
__device__ __forceinline__ vec4 depth_ratio_of(float current, vec4 previous) {   // select(current / previous, 1.0, previous == 0.0)
    return v4(previous.x == 0.0f ? 1.0f : current / previous.x, previous.y == 0.0f ? 1.0f : current / previous.y,
              previous.z == 0.0f ? 1.0f : current / previous.z, previous.w == 0.0f ? 1.0f : current / previous.w);
}
__device__ __forceinline__ void store16(uint2* plane, size_t i, vec4 v) {
    uvec2 w = pack_rgba16f(v);
    plane[i] = make_uint2(w.x, w.y);
}
// Windows.  Full-frame context: render-size images are tight RW x RH (x scale for the upscaled ones).  Tile (ratio 1):
// every image is stored over the tile's allocation, scaled by `scale` for the upscaled ones.
__device__ __forceinline__ Image16 render_image(const KParams& P, const uint2* plane, int scale) {
    const Band& b = P.band;
    if (!P.tile_images) return Image16{plane, scale * b.RW, scale * b.RH, 0, 0, scale * b.RW, scale * b.RH};
    return Image16{plane, scale * b.W, scale * b.H, scale * b.ax0, scale * b.a0, scale * b.AW, scale * (b.a1 - b.a0)};
}
__device__ __forceinline__ size_t render_image_index(const KParams& P, int scale, int x, int y) {
    const Band& b = P.band;
    if (!P.tile_images) return (size_t)y * (size_t)(scale * b.RW) + (size_t)x;
    return (size_t)(y - scale * b.a0) * (size_t)(scale * b.AW) + (size_t)(x - scale * b.ax0);
}
// The SMAA TU4x output and the TAA images after it: OW x OH texels (Band::OW / OH), stored tightly by a full-frame context and over
// 2 x the allocation by a tile (ratio 1: OW = 2 W exactly).  A store outside the texture is dropped, as the storage texture drops it.
__device__ __forceinline__ Image16 upscaled_image(const KParams& P, const uint2* plane) {
    const Band& b = P.band;
    if (!P.tile_images) return Image16{plane, b.OW, b.OH, 0, 0, b.OW, b.OH};
    return Image16{plane, 2 * b.W, 2 * b.H, 2 * b.ax0, 2 * b.a0, 2 * b.AW, 2 * (b.a1 - b.a0)};
}
__device__ __forceinline__ void upscaled_store(const KParams& P, uint2* plane, int x, int y, vec4 v) {
    const Band& b = P.band;
    if (!P.tile_images) {
        if (x >= b.OW || y >= b.OH) return;
        store16(plane, (size_t)y * (size_t)b.OW + (size_t)x, v);
    } else {
        store16(plane, (size_t)(y - 2 * b.a0) * (size_t)(2 * b.AW) + (size_t)(x - 2 * b.ax0), v);
    }
}
__device__ __forceinline__ Image32 gbuffer_image(const KParams& P, const float4* plane) {
    const Band& b = P.band;
    return Image32{plane, b.W, b.H, b.ax0, b.a0, b.AW, b.a1 - b.a0};
}
__device__ __forceinline__ const uint2* tone_plane(const KParams& P, uint32_t which) {
    return P.tile_images ? P.planes.tone_ring_db[which] : P.planes.tone_mapped_db[which];
}

// --------------------------------------------------------------------------------------------- smaa_tu4x
__global__ void __launch_bounds__(CTA_THREADS) k_smaa_tu4x(const __grid_constant__ KParams P) {  // smaa.wgsl:81-199
    int x, y;
    tile_pixel(x, y, P);
    if (!tile_active(P, x, y)) return;
    const int OW = P.tile_images ? 2 * P.band.W : P.band.OW, OH = P.tile_images ? 2 * P.band.H : P.band.OH, W = P.band.W, H = P.band.H;   // textureDimensions(output_texture)
    const uint32_t cur = P.in.frame.number % 2u, prv = 1u - cur;
    const Image16 render = render_image(P, tone_plane(P, cur), 1), previous_render = render_image(P, tone_plane(P, prv), 1);
    const Image32 position = gbuffer_image(P, P.planes.pos_depth_db[P.gbuffer_current]);
    const Image32 previous_position = gbuffer_image(P, P.planes.pos_depth_db[P.gbuffer_current ^ 1]);
    const Image32 velocity_uv = gbuffer_image(P, P.planes.velocity_uv_db[P.gbuffer_current]);
    const Image32 previous_velocity_uv = gbuffer_image(P, P.planes.velocity_uv_db[P.gbuffer_current ^ 1]);
    const int current_jitter = ((P.in.frame.number & 1u) == 0u) ? 0 : 1;
    const int previous_jitter = 1 - current_jitter;
    const vec2 uv = render_uv(P, x, y);
    const vec2 texel_size = v2(1.0f, 1.0f) / v2((float)OW, (float)OH);
    const vec2 uv_biases[5] = {v2(0.0f, 0.0f), v2(2.5f, 2.5f) * texel_size, v2(-2.5f, 2.5f) * texel_size, v2(2.5f, -2.5f) * texel_size,
                               v2(-2.5f, -2.5f) * texel_size};
    const int cox = 2 * x + current_jitter, coy = 2 * y + current_jitter;
    const vec3 current_color = xyz(render.nearest(uv));
    const int pox = 2 * x + previous_jitter, poy = 2 * y + previous_jitter;
    const vec2 previous_output_uv = (v2((float)pox, (float)poy) + 0.5f) / v2((float)OW, (float)OH);
    const vec2 deferred_texel = v2(1.0f, 1.0f) / v2((float)W, (float)H);
    const vec2 velocity = nearest_velocity(position, velocity_uv, previous_output_uv, deferred_texel);
    const vec2 previous_reprojected_uv = previous_output_uv - velocity;
    vec3 previous_color = xyz(previous_render.nearest(previous_reprojected_uv));
    const bool boundary_miss = fabsf(previous_reprojected_uv.x - 0.5f) > 0.5f || fabsf(previous_reprojected_uv.y - 0.5f) > 0.5f;
    auto instance_at = [&](vec2 u) {
        int px = min(max((int)floorf(u.x * (float)W), 0), W - 1), py = min(max((int)floorf(u.y * (float)H), 0), H - 1);
        px = min(max(px, P.band.ax0), P.band.ax1 - 1); py = min(max(py, P.band.a0), P.band.a1 - 1);     // window safety, see Image16::fetch
        return P.planes.instance_material[band_index(P.band, px, py)].x;
    };
    const float current_instance = instance_at(previous_output_uv);
    bool instance_miss = false;
    const float current_depth = position.nearest(previous_output_uv).w;
    bool depth_miss = current_depth == 0.0f;
    for (uint32_t i = 0u; i < 5u; i += 1u) {
        vec4 depth_ratio = depth_ratio_of(current_depth, previous_position.gather_w(previous_reprojected_uv + uv_biases[i]));
        depth_miss = depth_miss || any_lt(depth_ratio, 0.95f);
        float previous_instance = instance_at(previous_reprojected_uv + uv_biases[i]);
        instance_miss = instance_miss || (any_lt(depth_ratio, 0.95f) && fabsf(previous_instance - current_instance) > 1.0f);
    }
    vec4 pv = previous_velocity_uv.nearest(previous_reprojected_uv);
    const bool velocity_miss = distance2(velocity, v2(pv.x, pv.y)) > 0.0001f;
    if (boundary_miss || ((depth_miss || instance_miss) && velocity_miss)) {
        vec2 uv_bias = v2(0.0f, 0.0f);
        float min_ds = 10.0f;
        for (uint32_t i = 0u; i < 5u; i += 1u) {
            float dds = distance4(v4(current_depth), position.gather_w(previous_output_uv + uv_biases[i]));
            if (dds < min_ds) uv_bias = uv_biases[i];
            min_ds = fmin_(min_ds, dds);
        }
        vec4 g[4];
        render.gather(previous_output_uv + uv_bias, g);
        vec3 s1 = RGB_to_YCoCg(xyz(g[0])), s2 = RGB_to_YCoCg(xyz(g[1])), s3 = RGB_to_YCoCg(xyz(g[2])), s4 = RGB_to_YCoCg(xyz(g[3]));
        vec3 moment_1 = s1 + s2 + s3 + s4;
        vec3 moment_2 = s1 * s1 + s2 * s2 + s3 * s3 + s4 * s4;
        vec3 mean = moment_1 / 4.0f;
        vec3 variance = vsqrt((moment_2 / 4.0f) - (mean * mean));
        previous_color = RGB_to_YCoCg(previous_color);
        previous_color = clip_towards_aabb_center(previous_color, mean - variance, mean + variance);
        previous_color = YCoCg_to_RGB(previous_color);
    }
    vec2 sv = velocity / (2.0f * texel_size);
    float blend_factor = fmax_(fract(sv.x), fract(sv.y));
    float sn, cs;
    sincos_(blend_factor * TAU, &sn, &cs);
    blend_factor = clampf(-cs, 0.0f, 1.0f);
    vec3 remix_color = xyz(render.linear(previous_output_uv));
    previous_color = mix(previous_color, remix_color, blend_factor);
    upscaled_store(P, P.planes.upscale_output, cox, coy, v4(current_color, 1.0f));
    upscaled_store(P, P.planes.upscale_output, pox, poy, v4(previous_color, 1.0f));
}

__global__ void __launch_bounds__(CTA_THREADS) k_smaa_tu4x_extrapolate(const __grid_constant__ KParams P) {  // smaa.wgsl:201-271
    int x, y;
    tile_pixel(x, y, P);
    if (!tile_active(P, x, y)) return;
    const Image16 out = upscaled_image(P, P.planes.upscale_output);
    vec4 t = out.load(2 * x, 2 * y), b = out.load(2 * x + 1, 2 * y + 1), n = out.load(2 * x + 1, 2 * y - 1), e = out.load(2 * x + 2, 2 * y);
    vec4 s_ = out.load(2 * x, 2 * y + 2), w = out.load(2 * x - 1, 2 * y + 1);
    auto lum3 = [](vec4 a, vec4 c) { return luminance(vabs(xyz(a) - xyz(c))); };
    vec2 dh = v2(lum3(w, b), lum3(t, e));
    vec2 dv = v2(lum3(t, s_), lum3(n, b));
    vec2 factor_xy = v2(fmax_(dv.x, 0.001f) * fmax_(dv.y, 0.001f), fmax_(dh.x, 0.001f) * fmax_(dh.y, 0.001f));
    float factor_z = 1.0f / (factor_xy.x + factor_xy.y);
    auto blend = [&](vec4 tt, vec4 bb, vec4 ll, vec4 rr) {
        vec4 color = v4(0.0f);
        color = color + (ll + rr) * factor_xy.x;
        color = color + (tt + bb) * factor_xy.y;
        return color * (0.5f * factor_z);
    };
    upscaled_store(P, P.planes.upscale_output, 2 * x, 2 * y + 1, blend(t, s_, w, b));
    upscaled_store(P, P.planes.upscale_output, 2 * x + 1, 2 * y, blend(n, b, t, e));
}

// ------------------------------------------------------------------------------------------- taa_jasmine
// launched over 2 RW x 2 RH after SMAA TU4x (else RW x RH) as the reference dispatches it: col_hi / row_hi carry it
__global__ void __launch_bounds__(CTA_THREADS) k_taa_jasmine(const __grid_constant__ KParams P, int smaa) {  // taa.wgsl:79-170
    int x, y;
    tile_pixel(x, y, P);
    if (!tile_active(P, x, y)) return;
    // textureDimensions(output_texture): taa_output is created at the scale of upscale_output after SMAA TU4x (post_process.rs:717,726-729)
    const int OW = smaa ? (P.tile_images ? 2 * P.band.W : P.band.OW) : P.band.RW, OH = smaa ? (P.tile_images ? 2 * P.band.H : P.band.OH) : P.band.RH;
    const int W = P.band.W, H = P.band.H;
    if (x >= OW || y >= OH) return;        // the dispatch covers 2 RW x 2 RH (scaled_size *= 2, :1258): invocations outside the texture store nothing
    const uint32_t cur = P.in.frame.number % 2u, prv = 1u - cur;
    const Image16 render = smaa ? upscaled_image(P, P.planes.upscale_output) : render_image(P, tone_plane(P, cur), 1);
    const Image16 previous_render = smaa ? upscaled_image(P, P.planes.taa_output[prv]) : render_image(P, P.planes.taa_output[prv], 1);
    const Image32 position = gbuffer_image(P, P.planes.pos_depth_db[P.gbuffer_current]);
    const Image32 previous_position = gbuffer_image(P, P.planes.pos_depth_db[P.gbuffer_current ^ 1]);
    const Image32 velocity_uv = gbuffer_image(P, P.planes.velocity_uv_db[P.gbuffer_current]);
    const Image32 previous_velocity_uv = gbuffer_image(P, P.planes.velocity_uv_db[P.gbuffer_current ^ 1]);
    auto sample_previous = [&](vec2 u) { return clamp01(xyz(previous_render.linear(u))); };
    auto sample_render = [&](vec2 u) { return RGB_to_YCoCg(clamp01(xyz(render.nearest(u)))); };
    const vec2 size = v2((float)OW, (float)OH);
    const vec2 texel_size = v2(1.0f, 1.0f) / size;
    const vec2 uv = (v2((float)x, (float)y) + 0.5f) / size;
    const vec4 original_color = render.nearest(uv);
    const vec3 current_color = xyz(original_color);
    const vec2 velocity = nearest_velocity(position, velocity_uv, uv, texel_size);
    const vec2 previous_uv = uv - velocity;
    const bool boundary_miss = fabsf(previous_uv.x - 0.5f) > 0.5f || fabsf(previous_uv.y - 0.5f) > 0.5f;
    const vec2 uv_biases[5] = {v2(0.0f, 0.0f), v2(1.5f, 1.5f) * texel_size, v2(-1.5f, 1.5f) * texel_size, v2(1.5f, -1.5f) * texel_size,
                               v2(-1.5f, -1.5f) * texel_size};
    const vec4 current_position_depth = position.nearest(uv);
    bool has_content = current_position_depth.w > 0.0f;
    bool depth_miss = current_position_depth.w == 0.0f;
    bool position_miss = current_position_depth.w == 0.0f;
    for (uint32_t i = 0u; i < 5u; i += 1u) {
        vec4 pd = previous_position.gather_w(previous_uv + uv_biases[i]);
        vec4 depth_ratio = depth_ratio_of(current_position_depth.w, pd);
        has_content = has_content || pd.x > 0.0f || pd.y > 0.0f || pd.z > 0.0f || pd.w > 0.0f;
        depth_miss = depth_miss || any_lt(depth_ratio, 0.95f);
        vec3 previous_pos = xyz(previous_position.nearest(previous_uv + uv_biases[i]));
        position_miss = position_miss || length(xyz(current_position_depth) - previous_pos) > 0.5f;
    }
    auto store_output = [&](vec4 v) {
        if (smaa) upscaled_store(P, P.planes.taa_output[cur], x, y, v);
        else store16(P.planes.taa_output[cur], render_image_index(P, 1, x, y), v);
    };
    if (!has_content) {
        store_output(v4(P.in.frame.clear_color[0], P.in.frame.clear_color[1], P.in.frame.clear_color[2], P.in.frame.clear_color[3]));
        return;
    }
    vec4 pv = previous_velocity_uv.nearest(previous_uv);
    const bool velocity_miss = distance2(velocity, v2(pv.x, pv.y)) > 0.00005f;
    // 5-tap Catmull-Rom, taa.wgsl:121-139
    vec2 sample_position = (uv - velocity) * size;
    vec2 tp1 = v2(floorf(sample_position.x - 0.5f), floorf(sample_position.y - 0.5f)) + 0.5f;
    vec2 f = sample_position - tp1;
    auto W0 = [](float f_) { return f_ * (-0.5f + f_ * (1.0f - 0.5f * f_)); };
    auto W1 = [](float f_) { return 1.0f + f_ * f_ * (-2.5f + 1.5f * f_); };
    auto W2 = [](float f_) { return f_ * (0.5f + f_ * (2.0f - 1.5f * f_)); };
    auto W3 = [](float f_) { return f_ * f_ * (-0.5f + 0.5f * f_); };
    vec2 w0 = v2(W0(f.x), W0(f.y)), w1 = v2(W1(f.x), W1(f.y)), w2 = v2(W2(f.x), W2(f.y)), w3 = v2(W3(f.x), W3(f.y));
    vec2 w12 = w1 + w2;
    vec2 offset12 = w2 / (w1 + w2);
    vec2 tp0 = (tp1 - 1.0f) * texel_size;
    vec2 tp3 = (tp1 + 2.0f) * texel_size;
    vec2 tp12 = (tp1 + offset12) * texel_size;
    vec3 previous_color = v3(0.0f);
    previous_color = previous_color + sample_previous(v2(tp12.x, tp0.y)) * w12.x * w0.y;
    previous_color = previous_color + sample_previous(v2(tp0.x, tp12.y)) * w0.x * w12.y;
    previous_color = previous_color + sample_previous(v2(tp12.x, tp12.y)) * w12.x * w12.y;
    previous_color = previous_color + sample_previous(v2(tp3.x, tp12.y)) * w3.x * w12.y;
    previous_color = previous_color + sample_previous(v2(tp12.x, tp3.y)) * w12.x * w3.y;
    if (boundary_miss || (position_miss && velocity_miss && depth_miss)) {
        vec3 s_tl = sample_render(uv + v2(-texel_size.x, texel_size.y));
        vec3 s_tm = sample_render(uv + v2(0.0f, texel_size.y));
        vec3 s_tr = sample_render(uv + texel_size);
        vec3 s_ml = sample_render(uv - v2(texel_size.x, 0.0f));
        vec3 s_mm = RGB_to_YCoCg(current_color);
        vec3 s_mr = sample_render(uv + v2(texel_size.x, 0.0f));
        vec3 s_bl = sample_render(uv - texel_size);
        vec3 s_bm = sample_render(uv - v2(0.0f, texel_size.y));
        vec3 s_br = sample_render(uv + v2(texel_size.x, -texel_size.y));
        vec3 moment_1 = s_tl + s_tm + s_tr + s_ml + s_mm + s_mr + s_bl + s_bm + s_br;
        vec3 moment_2 = (s_tl * s_tl) + (s_tm * s_tm) + (s_tr * s_tr) + (s_ml * s_ml) + (s_mm * s_mm) + (s_mr * s_mr) + (s_bl * s_bl) +
                        (s_bm * s_bm) + (s_br * s_br);
        vec3 mean = moment_1 / 9.0f;
        vec3 variance = vsqrt((moment_2 / 9.0f) - (mean * mean));
        previous_color = RGB_to_YCoCg(previous_color);
        previous_color = clip_towards_aabb_center(previous_color, mean - variance, mean + variance);
        previous_color = YCoCg_to_RGB(previous_color);
    }
    vec3 output = mix(previous_color, current_color, 0.1f / P.in.frame.upscale_ratio);
    store_output(v4(output, original_color.w));
}

// --------------------------------------------------------------------------------------------- FSR 1.0 (Upscale::Fsr1)
// EASU + RCAS as the reference's SPIR-V blobs compute them (src/shaders/fsr/source.zip: FSR_Pass.glsl with SAMPLE_SLOW_FALLBACK
// -> FsrEasuF ffx_fsr1.h:315-437, FsrRcasF :684-772; dispatched by post_process.rs:1279-1308).  One thread per OUTPUT pixel of
// the camera target; the 12 (EASU) / 5 (RCAS) taps of neighbouring threads overlap almost completely and are served by L1.
// Algorithmic bytes per launch: EASU 8 B x render pixels read + 8 B x target pixels written; RCAS 16 B x target pixels.
// accumulation order of FsrEasuF (:423-434): b c i j f e k l h g o n — the order is part of the result (fp32 sums)
enum { TB = 0, TC = 1, TI = 2, TJ = 3, TF = 4, TE = 5, TK = 6, TL = 7, TH = 8, TG = 9, TO = 10, TN = 11 };

__device__ __forceinline__ void easu_edge(float& dir_x, float& dir_y, float& len, float w, float a, float b, float c, float d, float e) {
    // '+' of lumas: a above, b left, c centre, d right, e below (FsrEasuSetF :275-313)
    float gx = d - b, gy = e - a;
    float sx = clampf(fabsf(gx) * fsr_rcp_lo(fmax_(fabsf(d - c), fabsf(c - b))), 0.0f, 1.0f);
    float sy = clampf(fabsf(gy) * fsr_rcp_lo(fmax_(fabsf(e - c), fabsf(c - a))), 0.0f, 1.0f);
    dir_x += gx * w;
    len += (sx * sx) * w;
    dir_y += gy * w;
    len += (sy * sy) * w;
}

__global__ void __launch_bounds__(CTA_THREADS) k_fsr_easu(const __grid_constant__ KParams P, const FsrEasuConstants con) {
    int x, y;
    tile_pixel(x, y, P);
    if (!tile_active(P, x, y)) return;
    const uint32_t cur = P.in.frame.number % 2u;
    const Band& bd = P.band;
    const Image16 input{P.in.taa_jitter ? P.planes.taa_output[cur] : P.planes.tone_mapped_db[cur], bd.RW, bd.RH, 0, 0, bd.RW, bd.RH};
    float px = (float)x * con.scale_x + con.offset_x, py = (float)y * con.scale_y + con.offset_y;
    const float fxf = floorf(px), fyf = floorf(py);
    px -= fxf; py -= fyf;
    const int fx = (int)fxf, fy = (int)fyf;
    // texel offsets of the taps from f, in accumulation order; both loops are fully unrolled, the tables fold into immediates
    constexpr int TAP_DX[12] = {0, 1, -1, 0, 0, -1, 1, 2, 2, 1, 1, 0}, TAP_DY[12] = {-1, -1, 1, 1, 0, 0, 1, 1, 0, 0, 2, 2};
    vec3 t[12];
    float l[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {   // four clamped texel fetches per emulated gather (texture_gather.glsl), 12 distinct texels
        t[k] = xyz(input.texel(fx + TAP_DX[k], fy + TAP_DY[k]));
        l[k] = t[k].z * 0.5f + (t[k].x * 0.5f + t[k].y);
    }
    float dir_x = 0.0f, dir_y = 0.0f, len = 0.0f;
    easu_edge(dir_x, dir_y, len, (1.0f - px) * (1.0f - py), l[TB], l[TE], l[TF], l[TG], l[TJ]);
    easu_edge(dir_x, dir_y, len, px * (1.0f - py), l[TC], l[TF], l[TG], l[TH], l[TK]);
    easu_edge(dir_x, dir_y, len, (1.0f - px) * py, l[TF], l[TI], l[TJ], l[TK], l[TN]);
    easu_edge(dir_x, dir_y, len, px * py, l[TG], l[TJ], l[TK], l[TL], l[TO]);
    float r = dir_x * dir_x + dir_y * dir_y;
    const bool flat = r < (1.0f / 32768.0f);
    r = flat ? 1.0f : fsr_rsq_lo(r);
    dir_x = (flat ? 1.0f : dir_x) * r;
    dir_y = dir_y * r;
    len = len * 0.5f;
    len *= len;
    const float stretch = (dir_x * dir_x + dir_y * dir_y) * fsr_rcp_lo(fmax_(fabsf(dir_x), fabsf(dir_y)));
    const float len_x = 1.0f + (stretch - 1.0f) * len, len_y = 1.0f + -0.5f * len;
    const float lob = 0.5f + ((1.0f / 4.0f - 0.04f) - 0.5f) * len;
    const float clp = fsr_rcp_lo(lob);
    vec3 acc = v3(0.0f, 0.0f, 0.0f);
    float wsum = 0.0f;
#pragma unroll
    for (int k = 0; k < 12; ++k) {   // FsrEasuTapF :239-273
        const float ox = (float)TAP_DX[k] - px, oy = (float)TAP_DY[k] - py;
        const float vx = ((ox * dir_x) + (oy * dir_y)) * len_x, vy = ((ox * (-dir_y)) + (oy * dir_x)) * len_y;
        const float d2 = fmin_(vx * vx + vy * vy, clp);
        float wb = (2.0f / 5.0f) * d2 + -1.0f, wa = lob * d2 + -1.0f;
        wb *= wb; wa *= wa;
        wb = (25.0f / 16.0f) * wb + (-(25.0f / 16.0f - 1.0f));
        const float w = wb * wa;
        acc = acc + t[k] * w;
        wsum += w;
    }
    // de-ring against the 2x2 centre f g j k
    auto lo = [](float a, float b, float c, float d) { return fmin_(fmin_(a, fmin_(b, c)), d); };
    auto hi = [](float a, float b, float c, float d) { return fmax_(fmax_(a, fmax_(b, c)), d); };
    const float inv = 1.0f / wsum;
    vec3 pix;
    pix.x = fmin_(hi(t[TF].x, t[TG].x, t[TJ].x, t[TK].x), fmax_(lo(t[TF].x, t[TG].x, t[TJ].x, t[TK].x), acc.x * inv));
    pix.y = fmin_(hi(t[TF].y, t[TG].y, t[TJ].y, t[TK].y), fmax_(lo(t[TF].y, t[TG].y, t[TJ].y, t[TK].y), acc.y * inv));
    pix.z = fmin_(hi(t[TF].z, t[TG].z, t[TJ].z, t[TK].z), fmax_(lo(t[TF].z, t[TG].z, t[TJ].z, t[TK].z), acc.z * inv));
    store16(P.planes.upscale_output, (size_t)y * (size_t)bd.W + (size_t)x, v4(pix, 1.0f));
}

__global__ void __launch_bounds__(CTA_THREADS) k_fsr_rcas(const __grid_constant__ KParams P, const float sharp) {
    int x, y;
    tile_pixel(x, y, P);
    if (!tile_active(P, x, y)) return;
    const Band& bd = P.band;
    const Image16 input{P.planes.upscale_output, bd.W, bd.H, 0, 0, bd.W, bd.H};
    // all five loads first (texelFetch, zero outside the image), then arithmetic
    const vec3 b = xyz(input.load(x, y - 1)), d = xyz(input.load(x - 1, y)), e = xyz(input.load(x, y));
    const vec3 f = xyz(input.load(x + 1, y)), h = xyz(input.load(x, y + 1));
    auto channel_lobe = [](float bb, float dd, float ee, float ff, float hh) {
        const float mn = fmin_(fmin_(bb, fmin_(dd, ff)), hh), mx = fmax_(fmax_(bb, fmax_(dd, ff)), hh);
        const float hit_min = fmin_(mn, ee) * (1.0f / (4.0f * mx));
        const float hit_max = (1.0f - fmax_(mx, ee)) * (1.0f / (4.0f * mn + -4.0f));
        return fmax_(-hit_min, hit_max);
    };
    const float lr = channel_lobe(b.x, d.x, e.x, f.x, h.x), lg = channel_lobe(b.y, d.y, e.y, f.y, h.y), lb = channel_lobe(b.z, d.z, e.z, f.z, h.z);
    const float lobe = fmax_(-FSR_RCAS_LIMIT, fmin_(fmax_(lr, fmax_(lg, lb)), 0.0f)) * sharp;
    const float rcp = fsr_rcp_med(4.0f * lobe + 1.0f);
    const vec3 pix = v3((lobe * b.x + lobe * d.x + lobe * h.x + lobe * f.x + e.x) * rcp,
                        (lobe * b.y + lobe * d.y + lobe * h.y + lobe * f.y + e.y) * rcp,
                        (lobe * b.z + lobe * d.z + lobe * h.z + lobe * f.z + e.z) * rcp);
    store16(P.planes.upscale_sharpen_output, (size_t)y * (size_t)bd.W + (size_t)x, v4(pix, 1.0f));
}

static dim3 grid_for(const KParams& P) {
    int rows = P.row_hi - P.row_lo, cols = P.col_hi - P.col_lo;
    return dim3((unsigned)((cols + TILE_W - 1) / TILE_W), (unsigned)((rows + TILE_H - 1) / TILE_H), 1u);
}

}  // namespace hkd

using namespace hkd;

void hk_launch_smaa_tu4x(const KParams& P, cudaStream_t st) {
    if (P.row_hi <= P.row_lo || P.col_hi <= P.col_lo) return;
    k_smaa_tu4x<<<grid_for(P), CTA_THREADS, 0, st>>>(P);
}
void hk_launch_smaa_tu4x_extrapolate(const KParams& P, cudaStream_t st) {
    if (P.row_hi <= P.row_lo || P.col_hi <= P.col_lo) return;
    k_smaa_tu4x_extrapolate<<<grid_for(P), CTA_THREADS, 0, st>>>(P);
}
void hk_launch_taa_jasmine(const KParams& P, bool smaa, cudaStream_t st) {
    if (P.row_hi <= P.row_lo || P.col_hi <= P.col_lo) return;
    k_taa_jasmine<<<grid_for(P), CTA_THREADS, 0, st>>>(P, smaa ? 1 : 0);
}
void hk_launch_fsr_easu(const KParams& P, cudaStream_t st) {
    if (P.row_hi <= P.row_lo || P.col_hi <= P.col_lo) return;
    // FsrConstantsUniform (post_process.rs:518-534): input viewport = input size = scaled_size, output = the camera target
    const hk::FsrEasuConstants con = hk::fsr_easu_constants((float)P.band.RW, (float)P.band.RH, (float)P.band.W, (float)P.band.H);
    k_fsr_easu<<<grid_for(P), CTA_THREADS, 0, st>>>(P, con);
}
void hk_launch_fsr_rcas(const KParams& P, cudaStream_t st) {
    if (P.row_hi <= P.row_lo || P.col_hi <= P.col_lo) return;
    k_fsr_rcas<<<grid_for(P), CTA_THREADS, 0, st>>>(P, hk::fsr_rcas_constant(P.in.fsr_sharpness));
}
