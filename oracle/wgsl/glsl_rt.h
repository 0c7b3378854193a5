// glsl_rt.h — TEST INFRASTRUCTURE: the run-time the C++ translation of the reference's FSR 1.0 sources (oracle/wgsl/glsl2cpp.py) compiles
// against.  The reference ships EASU and RCAS as SPIR-V blobs next to the GLSL they were compiled from (src/shaders/fsr/source.zip:
// FSR_Pass.glsl, AMD's ffx_a.h / ffx_fsr1.h, texture_gather.glsl, compile.bat); the GLSL is read from the zip where it lies, run through
// the C preprocessor with the defines of fsr_pass_easu.glsl / fsr_pass_rcas.glsl, and compiled as C++ against the GLSL types and built-ins
// below (vectors, their operators and the shared interpretation of implementation-defined arithmetic come from wgsl_rt.h).
#pragma once
#include "wgsl_rt.h"

namespace wgsl {
namespace glsl {      // nested: unqualified names resolve to wgsl's built-ins before the C library's, GLSL's type names hide wgsl's templates

typedef vec<f32, 2> vec2; typedef vec<f32, 3> vec3; typedef vec<f32, 4> vec4;
typedef vec<u32, 2> uvec2; typedef vec<u32, 3> uvec3; typedef vec<u32, 4> uvec4;
typedef vec<i32, 2> ivec2; typedef vec<i32, 3> ivec3; typedef vec<i32, 4> ivec4;
typedef unsigned int uint;

// operators WGSL's shaders never needed
#define GLSL_SHIFT(OP)                                                                                                            \
    template <class T, int N> vec<T, N> operator OP(const vec<T, N>& a, const vec<T, N>& b) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] OP b.v[i]; return r; } \
    template <class T, int N> vec<T, N> operator OP(const vec<T, N>& a, typename wgsl_id<T>::type b) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] OP b; return r; }
GLSL_SHIFT(<<) GLSL_SHIFT(>>)
template <class T, int N> vec<T, N> operator~(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = ~a.v[i]; return r; }

using wgsl::abs;
template <int N> vec<i32, N> abs(const vec<i32, N>& a) { vec<i32, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] < 0 ? -a.v[i] : a.v[i]; return r; }
inline f32 inversesqrt(f32 a) { return 1.0f / sqrtf(a); }
template <int N> vec<f32, N> inversesqrt(const vec<f32, N>& a) { vec<f32, N> r; for (int i = 0; i < N; ++i) r.v[i] = 1.0f / sqrtf(a.v[i]); return r; }
inline uint floatBitsToUint(f32 a) { return bitcast<u32>(a); }
inline f32 uintBitsToFloat(uint a) { return bitcast<f32>(a); }
template <int N> vec<u32, N> floatBitsToUint(const vec<f32, N>& a) { vec<u32, N> r; for (int i = 0; i < N; ++i) r.v[i] = bitcast<u32>(a.v[i]); return r; }
template <int N> vec<f32, N> uintBitsToFloat(const vec<u32, N>& a) { vec<f32, N> r; for (int i = 0; i < N; ++i) r.v[i] = bitcast<f32>(a.v[i]); return r; }
inline uint packHalf2x16(const vec2& v) { return pack2x16float(v); }
inline vec2 unpackHalf2x16(uint p) { return unpack2x16float(p); }
inline uint packUnorm2x16(const vec2& v) { return pack2x16unorm(v); }
inline vec2 unpackUnorm2x16(uint p) { return unpack2x16unorm(p); }
inline uint packUnorm4x8(const vec4& v) { return pack4x8unorm(v); }
inline vec4 unpackUnorm4x8(uint p) { return unpack4x8unorm(p); }
inline uint bitfieldExtract(uint v, int off, int bits) { return bits == 0 ? 0u : (v >> off) & (bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u)); }
inline uint bitfieldInsert(uint base, uint ins, int off, int bits) {
    if (bits == 0) return base;
    const uint mask = (bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u)) << off;
    return (base & ~mask) | ((ins << off) & mask);
}
inline int findMSB(uint v) { return v ? 31 - __builtin_clz(v) : -1; }
inline f32 log2_(f32 a) { return log2f(a); }
inline f32 mod(f32 a, f32 b) { return a - b * floorf(a / b); }

// ------------------------------------------------------------------------------------------------ resources
struct texture2D : wgsl_texture {};
struct image2D : wgsl_texture {};
typedef wgsl::sampler sampler_t;                      // `sampler` is also a GLSL keyword-type: the translation renames the GLSL type to sampler_t
struct sampler2D {
    const wgsl_texture* t; const sampler_t* s;
    sampler2D(const wgsl_texture& tex, const sampler_t& smp) : t(&tex), s(&smp) {}
};
inline ivec2 textureSize(const sampler2D& ts, int) { return ivec2(ts.t->w, ts.t->h); }
// texture(): bilinear filtering with the sub-texel precision of a texture unit — coordinates snapped to 1 / 256 of a texel
// (VkPhysicalDeviceLimits::subTexelPrecisionBits = 8 on the desktop GPUs the reference runs on), then fp32 blending.  This matters
// for exactly one thing: texture_gather.glsl's `fakeTextureGather` samples half a texel around the gather point, i.e. AT the four
// texel centres up to the rounding of the coordinate arithmetic (~1e-4 texel); a texture unit returns those texels exactly, a blend
// with unquantised fp32 weights would mix in 1e-4 of the neighbours (seen as 1 f16 ulp on one pixel in ~6000).
inline vec4 texture(const sampler2D& ts, const vec2& p) {
    const wgsl_texture& t = *ts.t; const sampler_t& s = *ts.s;
    if (!t.p || t.w <= 0 || t.h <= 0) return vec4(0.0f);
    if (!s.linear) return textureSampleLevel(t, s, p, 0.0f);
    const f32 px = floorf((p.x * (f32)t.w - 0.5f) * 256.0f + 0.5f) / 256.0f, py = floorf((p.y * (f32)t.h - 0.5f) * 256.0f + 0.5f) / 256.0f;
    const f32 x0 = floorf(px), y0 = floorf(py), ax = px - x0, ay = py - y0;
    const int ix0 = wgsl_wrap((int)x0, t.w, s.mode_u), ix1 = wgsl_wrap((int)x0 + 1, t.w, s.mode_u);
    const int iy0 = wgsl_wrap((int)y0, t.h, s.mode_v), iy1 = wgsl_wrap((int)y0 + 1, t.h, s.mode_v);
    const vec4 top = t.load(ix0, iy0) * (1.0f - ax) + t.load(ix1, iy0) * ax, bottom = t.load(ix0, iy1) * (1.0f - ax) + t.load(ix1, iy1) * ax;
    return top * (1.0f - ay) + bottom * ay;
}
inline vec4 textureLod(const sampler2D& ts, const vec2& p, f32) { return texture(ts, p); }
inline vec4 texelFetch(const sampler2D& ts, const ivec2& p, int) { return ts.t->load(p.x, p.y); }
inline void imageStore(const image2D& img, const ivec2& p, const vec4& v) { img.store(p.x, p.y, v); }

inline uvec3& glsl_local_id() { static thread_local uvec3 v; return v; }
inline uvec3& glsl_group_id() { static thread_local uvec3 v; return v; }
#define gl_LocalInvocationID (glsl_local_id())
#define gl_WorkGroupID (glsl_group_id())

}  // namespace glsl
}  // namespace wgsl
