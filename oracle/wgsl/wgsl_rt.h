// wgsl_rt.h — TEST INFRASTRUCTURE: the run-time the C++ translation of the reference's WGSL (oracle/wgsl/wgsl2cpp.py) compiles
// against.  Types and built-in functions of WGSL as the shaders use them, with IEEE fp32 arithmetic evaluated in source order (no
// contraction: the libraries are built with -ffp-contract=off) — one valid execution of the shader text.  Where WGSL leaves the
// result to the implementation — FMA use inside dot / cross / matrix products, the accuracy of exp / sin / cos / pow, NaN handling of
// min / max, normalize as a reciprocal square root — this run-time takes the choices include/hk_math.h documents (GPU-style FMA
// chains, its polynomial exp2 / sincos, pow by squarings for the exponents the shaders use), i.e. the same ones the oracle and the
// CUDA kernels are built on: what is being pinned is the shaders' logic, wiring and formula structure, which those choices do not touch.
// Memory layouts follow WGSL's host-shareable rules (vec3: 12 bytes; matCxR: C columns padded to the column's alignment), so
// buffers are bound as the bytes the reference's host code writes.  Textures read zero outside their extent and drop stores there,
// as wgpu's robust access does; Rgba16Float / Rgba8Snorm stores round like the GPU formats do (RTNE for f16).
#pragma once
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <type_traits>
#include <thread>
#include <vector>

#include "hk_math.h"      // the ONE interpretation of WGSL's implementation-defined operations shared by oracle and kernels (see its header)

namespace wgsl {

typedef float f32;
typedef uint32_t u32;
typedef int32_t i32;

template <class T> struct wgsl_id { typedef T type; };

// ------------------------------------------------------------------------------------------------ vectors
template <class T, int N> struct vec_data;
template <class T> struct vec_data<T, 2> { union { struct { T x, y; }; T v[2]; }; };
template <class T> struct vec_data<T, 3> { union { struct { T x, y, z; }; T v[3]; }; };
template <class T> struct vec_data<T, 4> { union { struct { T x, y, z, w; }; T v[4]; }; };

template <class T, int N> struct vec;
template <class T> struct wgsl_count { static const int n = 1; };
template <class T, int N> struct wgsl_count<vec<T, N> > { static const int n = N; };

template <class T, int N> struct vec : vec_data<T, N> {
    using vec_data<T, N>::v;
    vec() { for (int i = 0; i < N; ++i) v[i] = T(); }
    template <class A0, class... A> vec(const A0& a0, const A&... a) {
        T* p = v;
        if (sizeof...(A) == 0 && wgsl_count<A0>::n == 1) { T s; T* q = &s; put(q, a0); for (int i = 0; i < N; ++i) v[i] = s; return; }   // splat
        put(p, a0);
        int dummy[] = {0, (put(p, a), 0)...};
        (void)dummy;
    }
    template <class U> static void put(T*& p, const U& s) { *p++ = (T)s; }
    template <class U, int M> static void put(T*& p, const vec<U, M>& s) { for (int i = 0; i < M; ++i) *p++ = (T)s.v[i]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& operator[](u32 i) { return v[i]; }
    const T& operator[](u32 i) const { return v[i]; }
    template <int... I> vec<T, sizeof...(I)> swz() const { return vec<T, sizeof...(I)>(v[I]...); }
    // GLSL only (glsl_rt.h): `a != b` on vectors is ONE bool there (any component differs); WGSL has no conversion and never asks for it
    template <class U = T, class = typename std::enable_if<std::is_same<U, bool>::value>::type>
    explicit operator bool() const { for (int i = 0; i < N; ++i) if (v[i]) return true; return false; }
};
template <class T> using vec2 = vec<T, 2>;
template <class T> using vec3 = vec<T, 3>;
template <class T> using vec4 = vec<T, 4>;
static_assert(sizeof(vec3<f32>) == 12 && sizeof(vec2<u32>) == 8 && sizeof(vec4<f32>) == 16 && alignof(vec4<f32>) == 4, "vector layout");

#define WGSL_BINOP(OP)                                                                                                            \
    template <class T, int N> vec<T, N> operator OP(const vec<T, N>& a, const vec<T, N>& b) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] OP b.v[i]; return r; } \
    template <class T, int N> vec<T, N> operator OP(const vec<T, N>& a, typename wgsl_id<T>::type b) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] OP b; return r; } \
    template <class T, int N> vec<T, N> operator OP(typename wgsl_id<T>::type a, const vec<T, N>& b) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = a OP b.v[i]; return r; } \
    template <class T, int N> vec<T, N>& operator OP##=(vec<T, N>& a, const vec<T, N>& b) { for (int i = 0; i < N; ++i) a.v[i] = a.v[i] OP b.v[i]; return a; } \
    template <class T, int N> vec<T, N>& operator OP##=(vec<T, N>& a, typename wgsl_id<T>::type b) { for (int i = 0; i < N; ++i) a.v[i] = a.v[i] OP b; return a; }
WGSL_BINOP(+) WGSL_BINOP(-) WGSL_BINOP(*) WGSL_BINOP(/) WGSL_BINOP(&) WGSL_BINOP(|) WGSL_BINOP(^)
// integer division / remainder: WGSL leaves x / 0 and x % 0 to the implementation; here 0 divisors give 0 (no trap)
template <int N> vec<i32, N> operator%(const vec<i32, N>& a, const vec<i32, N>& b) { vec<i32, N> r; for (int i = 0; i < N; ++i) r.v[i] = b.v[i] ? a.v[i] % b.v[i] : 0; return r; }
template <int N> vec<i32, N> operator%(const vec<i32, N>& a, i32 b) { vec<i32, N> r; for (int i = 0; i < N; ++i) r.v[i] = b ? a.v[i] % b : 0; return r; }
template <int N> vec<u32, N> operator%(const vec<u32, N>& a, u32 b) { vec<u32, N> r; for (int i = 0; i < N; ++i) r.v[i] = b ? a.v[i] % b : 0; return r; }
template <class T, int N> vec<T, N> operator-(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = -a.v[i]; return r; }
#define WGSL_CMP(OP)                                                                                                              \
    template <class T, int N> vec<bool, N> operator OP(const vec<T, N>& a, const vec<T, N>& b) { vec<bool, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] OP b.v[i]; return r; } \
    template <class T, int N> vec<bool, N> operator OP(const vec<T, N>& a, typename wgsl_id<T>::type b) { vec<bool, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] OP b; return r; }
WGSL_CMP(<) WGSL_CMP(>) WGSL_CMP(<=) WGSL_CMP(>=) WGSL_CMP(==) WGSL_CMP(!=)
template <int N> vec<bool, N> operator!(const vec<bool, N>& a) { vec<bool, N> r; for (int i = 0; i < N; ++i) r.v[i] = !a.v[i]; return r; }
template <int N> bool all(const vec<bool, N>& a) { for (int i = 0; i < N; ++i) if (!a.v[i]) return false; return true; }
template <int N> bool any(const vec<bool, N>& a) { for (int i = 0; i < N; ++i) if (a.v[i]) return true; return false; }
inline bool all(bool a) { return a; }
inline bool any(bool a) { return a; }

// ------------------------------------------------------------------------------------------------ arrays
template <class T, unsigned N = 0u> struct array {
    T e[N];
    array() { for (unsigned i = 0; i < N; ++i) e[i] = T(); }
    template <class... A> array(const T& a0, const A&... a) : e{a0, (T)a...} {}
    T& operator[](i32 i) { return e[(u32)i < N ? (u32)i : N - 1]; }                  // out of bounds: clamped, as robust access does
    const T& operator[](i32 i) const { return e[(u32)i < N ? (u32)i : N - 1]; }
    T& operator[](u32 i) { return e[i < N ? i : N - 1]; }
    const T& operator[](u32 i) const { return e[i < N ? i : N - 1]; }
};
template <class T> struct array<T, 0u> {      // runtime-sized: a view of bound memory
    T* ptr = nullptr;
    size_t len = 0;
    T& at(size_t i) const { static thread_local T sink; if (i < len) return ptr[i]; sink = T(); return sink; }    // robust access: reads 0, stores dropped
    T& operator[](i32 i) const { return at((size_t)(u32)i); }
    T& operator[](u32 i) const { return at(i); }
};
template <class T> u32 arrayLength(const array<T, 0u>* a) { return (u32)a->len; }

// ------------------------------------------------------------------------------------------------ matrices (columns padded as WGSL lays them out)
template <class T, int C, int R> struct mat {
    struct column { vec<T, R> v; T pad[(R == 3) ? 1 : 0]; };
    column c[C];
    mat() {}
    template <class... A> mat(const vec<T, R>& c0, const A&... rest) { const vec<T, R> cols[] = {c0, rest...}; for (int i = 0; i < C; ++i) c[i].v = cols[i]; }
    vec<T, R>& operator[](i32 i) { return c[i].v; }
    const vec<T, R>& operator[](i32 i) const { return c[i].v; }
    vec<T, R>& operator[](u32 i) { return c[i].v; }
    const vec<T, R>& operator[](u32 i) const { return c[i].v; }
};
template <class T> using mat3x3 = mat<T, 3, 3>;
template <class T> using mat4x4 = mat<T, 4, 4>;
template <class T> using mat2x2 = mat<T, 2, 2>;
static_assert(sizeof(mat3x3<f32>) == 48 && sizeof(mat4x4<f32>) == 64, "matrix layout");
// M * v = sum over columns of column * component, accumulated left to right
template <class T, int C, int R> vec<T, R> operator*(const mat<T, C, R>& m, const vec<T, C>& v) {
    vec<T, R> r = m[0] * v.v[0];
    for (int j = 1; j < C; ++j) for (int i = 0; i < R; ++i) r.v[i] = fmaf(m[j].v[i], v.v[j], r.v[i]);      // FMA chain, as GPU compilers emit it
    return r;
}
template <class T, int C, int R> vec<T, C> operator*(const vec<T, R>& v, const mat<T, C, R>& m) {      // row vector times matrix
    vec<T, C> r;
    for (int j = 0; j < C; ++j) { T s = v.v[0] * m[j].v[0]; for (int i = 1; i < R; ++i) s = s + v.v[i] * m[j].v[i]; r.v[j] = s; }
    return r;
}
template <class T, int C, int R, int K> mat<T, K, R> operator*(const mat<T, C, R>& a, const mat<T, K, C>& b) {
    mat<T, K, R> r;
    for (int k = 0; k < K; ++k) r[k] = a * b[k];
    return r;
}
template <class T, int C, int R> mat<T, C, R> operator*(const mat<T, C, R>& m, typename wgsl_id<T>::type s) { mat<T, C, R> r; for (int j = 0; j < C; ++j) r[j] = m[j] * s; return r; }
template <class T, int C, int R> mat<T, R, C> transpose(const mat<T, C, R>& m) { mat<T, R, C> r; for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r[i].v[j] = m[j].v[i]; return r; }

// ------------------------------------------------------------------------------------------------ built-in functions
inline f32 wgsl_sin(f32 a) { f32 s, c; hk::sincos_(a, &s, &c); return s; }
inline f32 wgsl_cos(f32 a) { f32 s, c; hk::sincos_(a, &s, &c); return c; }
#define WGSL_MAP1(NAME, EXPR)                                                                                                     \
    inline f32 NAME(f32 a) { return EXPR; }                                                                                       \
    template <int N> vec<f32, N> NAME(const vec<f32, N>& x) { vec<f32, N> r; for (int i = 0; i < N; ++i) { const f32 a = x.v[i]; r.v[i] = EXPR; } return r; }
WGSL_MAP1(abs, ::fabsf(a)) WGSL_MAP1(floor, ::floorf(a)) WGSL_MAP1(ceil, ceilf(a)) WGSL_MAP1(trunc, truncf(a)) WGSL_MAP1(round, nearbyintf(a))
WGSL_MAP1(fract, hk::fract(a)) WGSL_MAP1(sqrt, sqrtf(a)) WGSL_MAP1(inverseSqrt, 1.0f / sqrtf(a)) WGSL_MAP1(exp, hk::exp_(a)) WGSL_MAP1(exp2, hk::exp2_(a))
WGSL_MAP1(log, logf(a)) WGSL_MAP1(log2, log2f(a)) WGSL_MAP1(sin, wgsl_sin(a)) WGSL_MAP1(cos, wgsl_cos(a)) WGSL_MAP1(tan, tanf(a)) WGSL_MAP1(acos, acosf(a))
WGSL_MAP1(asin, asinf(a)) WGSL_MAP1(atan, atanf(a)) WGSL_MAP1(sign, hk::signf(a))
inline i32 abs(i32 a) { return a < 0 ? -a : a; }
template <class T> T min_s(T a, T b);
template <class T> T max_s(T a, T b);
template <> inline f32 min_s<f32>(f32 a, f32 b) { return hk::fmin_(a, b); }       // IEEE minNum / maxNum (hk_math.h)
template <> inline f32 max_s<f32>(f32 a, f32 b) { return hk::fmax_(a, b); }
#define WGSL_MAP2(NAME, EXPR)                                                                                                     \
    template <class T> T NAME##_s(T a, T b) { return EXPR; }                                                                      \
    inline f32 NAME(f32 a, f32 b) { return NAME##_s<f32>(a, b); }                                                                 \
    inline i32 NAME(i32 a, i32 b) { return NAME##_s<i32>(a, b); }                                                                 \
    inline u32 NAME(u32 a, u32 b) { return NAME##_s<u32>(a, b); }                                                                 \
    template <class T, int N> vec<T, N> NAME(const vec<T, N>& x, const vec<T, N>& y) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = NAME##_s<T>(x.v[i], y.v[i]); return r; }
WGSL_MAP2(min, (b < a) ? b : a) WGSL_MAP2(max, (a < b) ? b : a)

// the exponents the shaders use have exact multiplicative forms (hk_math.h); anything else would go to libm
inline f32 pow(f32 a, f32 b) {
    if (b == 16.0f) return hk::pow16(a);
    if (b == 0.25f) return hk::pow025(a);
    if (b == 5.0f) return hk::pow5(a);
    if (b == 2.0f) return a * a;
    return powf(a, b);
}
template <int N> vec<f32, N> pow(const vec<f32, N>& a, const vec<f32, N>& b) { vec<f32, N> r; for (int i = 0; i < N; ++i) r.v[i] = pow(a.v[i], b.v[i]); return r; }
inline f32 atan2(f32 a, f32 b) { return atan2f(a, b); }
inline f32 step(f32 edge, f32 x) { return x < edge ? 0.0f : 1.0f; }
template <class T> T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
template <class T, int N> vec<T, N> clamp(const vec<T, N>& x, typename wgsl_id<T>::type lo, typename wgsl_id<T>::type hi) { return min(max(x, vec<T, N>(lo)), vec<T, N>(hi)); }
inline f32 mix(f32 a, f32 b, f32 t) { return hk::mixf(a, b, t); }
template <int N> vec<f32, N> mix(const vec<f32, N>& a, const vec<f32, N>& b, f32 t) { vec<f32, N> r; for (int i = 0; i < N; ++i) r.v[i] = hk::mixf(a.v[i], b.v[i], t); return r; }
template <int N> vec<f32, N> mix(const vec<f32, N>& a, const vec<f32, N>& b, const vec<f32, N>& t) { vec<f32, N> r; for (int i = 0; i < N; ++i) r.v[i] = hk::mixf(a.v[i], b.v[i], t.v[i]); return r; }
inline f32 smoothstep(f32 lo, f32 hi, f32 x) { const f32 t = clamp((x - lo) / (hi - lo), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
template <int N> f32 dot(const vec<f32, N>& a, const vec<f32, N>& b) { f32 s = a.v[0] * b.v[0]; for (int i = 1; i < N; ++i) s = fmaf(a.v[i], b.v[i], s); return s; }   // FMA chain
template <class T, int N> T dot(const vec<T, N>& a, const vec<T, N>& b) { T s = a.v[0] * b.v[0]; for (int i = 1; i < N; ++i) s = s + a.v[i] * b.v[i]; return s; }
inline vec<f32, 3> cross(const vec<f32, 3>& a, const vec<f32, 3>& b) { return vec<f32, 3>(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))); }
template <int N> f32 length(const vec<f32, N>& a) { return sqrtf(dot(a, a)); }
inline f32 length(f32 a) { return ::fabsf(a); }
template <int N> f32 distance(const vec<f32, N>& a, const vec<f32, N>& b) { return length(a - b); }
template <int N> vec<f32, N> normalize(const vec<f32, N>& a) { const f32 inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
template <int N> vec<f32, N> reflect(const vec<f32, N>& e1, const vec<f32, N>& e2) { return e1 - e2 * (2.0f * dot(e2, e1)); }
// select(f, t, cond): t where cond
template <class T> T select(const T& f, const T& t, bool c) { return c ? t : f; }
inline f32 select(f32 f, f32 t, bool c) { return c ? t : f; }
template <class T, int N> vec<T, N> select(const vec<T, N>& f, const vec<T, N>& t, const vec<bool, N>& c) { vec<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = c.v[i] ? t.v[i] : f.v[i]; return r; }
template <class T, class U> T bitcast(const U& u) { static_assert(sizeof(T) == sizeof(U), "bitcast size"); T t; memcpy(&t, &u, sizeof(T)); return t; }
inline u32 countOneBits(u32 v) { return (u32)__builtin_popcount(v); }

// f16 conversions (round to nearest even; overflow -> infinity like the GPU's pack does for finite inputs beyond 65504)
inline uint16_t wgsl_f32_to_f16(f32 f) {
    u32 x; memcpy(&x, &f, 4);
    const u32 sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u | ((x >> 13) & 0x3ffu) : 0u));
    if (x >= 0x47800000u) return (uint16_t)(sign | 0x7c00u);                      // >= 65536: infinity
    if (x >= 0x38800000u) {                                                       // normal half
        const u32 m = x - 0x38000000u, r = m + 0xfffu + ((m >> 13) & 1u);
        return (uint16_t)(sign | (r >> 13));
    }
    if (x < 0x33000000u) return (uint16_t)sign;                                   // below half of the smallest subnormal
    const u32 e = x >> 23, m = (x & 0x7fffffu) | 0x800000u, shift = 126u - e;     // subnormal half: m * 2^(e-150) in units of 2^-24
    const u32 q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1u);
    return (uint16_t)(sign | (q + ((rem > half || (rem == half && (q & 1u))) ? 1u : 0u)));
}
inline f32 wgsl_f16_to_f32(uint16_t h) {
    const u32 sign = ((u32)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    u32 x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { f32 f = (f32)m * 5.9604644775390625e-08f; memcpy(&x, &f, 4); x |= sign; }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    f32 f; memcpy(&f, &x, 4);
    return f;
}
inline u32 pack2x16float(const vec2<f32>& v) { return (u32)wgsl_f32_to_f16(v.x) | ((u32)wgsl_f32_to_f16(v.y) << 16); }
inline vec2<f32> unpack2x16float(u32 p) { return vec2<f32>(wgsl_f16_to_f32((uint16_t)(p & 0xffffu)), wgsl_f16_to_f32((uint16_t)(p >> 16))); }
inline u32 pack2x16unorm(const vec2<f32>& v) {
    u32 r = 0;
    for (int i = 0; i < 2; ++i) r |= (u32)floorf(0.5f + 65535.0f * fminf(fmaxf(v.v[i], 0.0f), 1.0f)) << (16 * i);
    return r;
}
inline vec2<f32> unpack2x16unorm(u32 p) { return vec2<f32>((f32)(p & 0xffffu) / 65535.0f, (f32)(p >> 16) / 65535.0f); }
inline u32 pack4x8snorm(const vec4<f32>& v) {
    u32 r = 0;
    for (int i = 0; i < 4; ++i) r |= ((u32)(i32)floorf(0.5f + 127.0f * fminf(fmaxf(v.v[i], -1.0f), 1.0f)) & 0xffu) << (8 * i);
    return r;
}
inline vec4<f32> unpack4x8snorm(u32 p) {
    vec4<f32> r;
    for (int i = 0; i < 4; ++i) r.v[i] = fmaxf((f32)(int8_t)((p >> (8 * i)) & 0xffu) / 127.0f, -1.0f);
    return r;
}
inline u32 pack4x8unorm(const vec4<f32>& v) {
    u32 r = 0;
    for (int i = 0; i < 4; ++i) r |= (u32)floorf(0.5f + 255.0f * fminf(fmaxf(v.v[i], 0.0f), 1.0f)) << (8 * i);
    return r;
}
inline vec4<f32> unpack4x8unorm(u32 p) { vec4<f32> r; for (int i = 0; i < 4; ++i) r.v[i] = (f32)((p >> (8 * i)) & 0xffu) / 255.0f; return r; }

// ------------------------------------------------------------------------------------------------ textures and samplers
enum { WGSL_RGBA32F = 0, WGSL_RGBA16F = 1, WGSL_R32F = 2, WGSL_RG32F = 3, WGSL_RGBA8SNORM = 4, WGSL_RGBA8UNORM = 5, WGSL_RGBA8SRGB = 6, WGSL_RG32U = 7 };
struct rgba16float {}; struct r32float {}; struct rgba32float {}; struct rgba8unorm {}; struct read_write {}; struct write {}; struct read {};

struct wgsl_texture {
    void* p = nullptr;
    int w = 0, h = 0, format = 0;
    void bind(void* ptr, int width, int height, int fmt) { p = ptr; w = width; h = height; format = fmt; }
    bool inside(int x, int y) const { return p && x >= 0 && y >= 0 && x < w && y < h; }
    vec4<f32> load(int x, int y) const {
        if (!inside(x, y)) return vec4<f32>(0.0f);
        const size_t i = (size_t)y * w + x;
        switch (format) {
            case WGSL_RGBA32F: { const f32* t = (const f32*)p + 4 * i; return vec4<f32>(t[0], t[1], t[2], t[3]); }
            case WGSL_RGBA16F: { const uint16_t* t = (const uint16_t*)p + 4 * i; return vec4<f32>(wgsl_f16_to_f32(t[0]), wgsl_f16_to_f32(t[1]), wgsl_f16_to_f32(t[2]), wgsl_f16_to_f32(t[3])); }
            case WGSL_R32F: return vec4<f32>(((const f32*)p)[i], 0.0f, 0.0f, 1.0f);
            case WGSL_RG32F: { const f32* t = (const f32*)p + 2 * i; return vec4<f32>(t[0], t[1], 0.0f, 1.0f); }
            case WGSL_RGBA8SNORM: return unpack4x8snorm(((const u32*)p)[i]);
            case WGSL_RGBA8UNORM: case WGSL_RGBA8SRGB: {      // texel decode as the sampler unit's table would: from the byte, in double, rounded once
                const uint8_t* t = (const uint8_t*)p + 4 * i;
                vec4<f32> c;
                for (int k = 0; k < 4; ++k) {
                    const double v = t[k] / 255.0;
                    c.v[k] = (f32)((format == WGSL_RGBA8SRGB && k < 3) ? (v <= 0.04045 ? v / 12.92 : ::pow((v + 0.055) / 1.055, 2.4)) : v);
                }
                return c;
            }
            default: return vec4<f32>(0.0f);
        }
    }
    vec4<u32> load_u(int x, int y) const {
        if (!inside(x, y)) return vec4<u32>(0u);
        const u32* t = (const u32*)p + 2 * ((size_t)y * w + x);
        return vec4<u32>(t[0], t[1], 0u, 1u);
    }
    void store(int x, int y, const vec4<f32>& c) const {
        if (!inside(x, y)) return;
        const size_t i = (size_t)y * w + x;
        switch (format) {
            case WGSL_RGBA32F: { f32* t = (f32*)p + 4 * i; for (int k = 0; k < 4; ++k) t[k] = c.v[k]; break; }
            case WGSL_RGBA16F: { uint16_t* t = (uint16_t*)p + 4 * i; for (int k = 0; k < 4; ++k) t[k] = wgsl_f32_to_f16(c.v[k]); break; }
            case WGSL_R32F: ((f32*)p)[i] = c.x; break;
            case WGSL_RG32F: { f32* t = (f32*)p + 2 * i; t[0] = c.x; t[1] = c.y; break; }
            case WGSL_RGBA8SNORM: ((u32*)p)[i] = pack4x8snorm(c); break;
            case WGSL_RGBA8UNORM: ((u32*)p)[i] = pack4x8unorm(c); break;
            default: break;
        }
    }
};
template <class T> struct texture_2d : wgsl_texture {};
template <class F, class A> struct texture_storage_2d : wgsl_texture {};
struct sampler {
    int mode_u = 1, mode_v = 1, linear = 0;          // 0 repeat, 1 clamp to edge, 2 mirror repeat
    void set(int u, int v, int l) { mode_u = u; mode_v = v; linear = l; }
};
template <class T> struct binding_array {
    std::vector<T> items;
    T& at(int i) { if ((size_t)i >= items.size()) items.resize((size_t)i + 1); return items[(size_t)i]; }
    T& operator[](u32 i) { static T none; return i < items.size() ? items[i] : none; }
    T& operator[](i32 i) { return (*this)[(u32)i]; }
};
inline vec2<i32> textureDimensions(const wgsl_texture& t) { return vec2<i32>(t.w, t.h); }
inline vec2<i32> textureDimensions(const wgsl_texture& t, i32) { return vec2<i32>(t.w, t.h); }
inline vec4<f32> textureLoad(const texture_2d<f32>& t, const vec2<i32>& c, i32) { return t.load(c.x, c.y); }
inline vec4<u32> textureLoad(const texture_2d<u32>& t, const vec2<i32>& c, i32) { return t.load_u(c.x, c.y); }
template <class F, class A> vec4<f32> textureLoad(const texture_storage_2d<F, A>& t, const vec2<i32>& c) { return t.load(c.x, c.y); }
template <class F, class A> void textureStore(const texture_storage_2d<F, A>& t, const vec2<i32>& c, const vec4<f32>& v) { t.store(c.x, c.y, v); }
inline int wgsl_wrap(int i, int n, int mode) {
    if (mode == 0) { i %= n; return i < 0 ? i + n : i; }
    if (mode == 2) { const int period = 2 * n; i %= period; if (i < 0) i += period; return i < n ? i : period - 1 - i; }
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
// mip level 0 only (the path samples level 0 everywhere); nearest = the texel containing the coordinate, linear = the usual four taps
inline vec4<f32> textureSampleLevel(const wgsl_texture& t, const sampler& s, const vec2<f32>& uv, f32) {
    if (!t.p || t.w <= 0 || t.h <= 0) return vec4<f32>(0.0f);
    const f32 fx = uv.x * (f32)t.w, fy = uv.y * (f32)t.h;
    if (!s.linear) return t.load(wgsl_wrap((int)floorf(fx), t.w, s.mode_u), wgsl_wrap((int)floorf(fy), t.h, s.mode_v));
    const f32 px = fx - 0.5f, py = fy - 0.5f, x0 = floorf(px), y0 = floorf(py), ax = px - x0, ay = py - y0;
    const int ix0 = wgsl_wrap((int)x0, t.w, s.mode_u), ix1 = wgsl_wrap((int)x0 + 1, t.w, s.mode_u);
    const int iy0 = wgsl_wrap((int)y0, t.h, s.mode_v), iy1 = wgsl_wrap((int)y0 + 1, t.h, s.mode_v);
    const vec4<f32> top = t.load(ix0, iy0) * (1.0f - ax) + t.load(ix1, iy0) * ax, bottom = t.load(ix0, iy1) * (1.0f - ax) + t.load(ix1, iy1) * ax;
    return top * (1.0f - ay) + bottom * ay;
}

// textureGather(component, t, s, uv): the four texels a bilinear sample at uv would blend, in WGSL's order
// x = (u_min, v_max), y = (u_max, v_max), z = (u_max, v_min), w = (u_min, v_min); footprint as textureSampleLevel's linear branch
inline vec4<f32> textureGather(i32 component, const wgsl_texture& t, const sampler& s, const vec2<f32>& uv) {
    if (!t.p || t.w <= 0 || t.h <= 0) return vec4<f32>(0.0f);
    const f32 px = uv.x * (f32)t.w - 0.5f, py = uv.y * (f32)t.h - 0.5f;
    const int x0 = (int)floorf(px), y0 = (int)floorf(py);
    const int ix0 = wgsl_wrap(x0, t.w, s.mode_u), ix1 = wgsl_wrap(x0 + 1, t.w, s.mode_u);
    const int iy0 = wgsl_wrap(y0, t.h, s.mode_v), iy1 = wgsl_wrap(y0 + 1, t.h, s.mode_v);
    const vec4<f32> a = t.load(ix0, iy1), b = t.load(ix1, iy1), c = t.load(ix1, iy0), d = t.load(ix0, iy0);
    const f32 av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cv[4] = {c.x, c.y, c.z, c.w}, dv[4] = {d.x, d.y, d.z, d.w};
    const int k = component & 3;
    return vec4<f32>(av[k], bv[k], cv[k], dv[k]);
}

// ------------------------------------------------------------------------------------------------ derivatives (fragment stage)
// dpdx / dpdy of an arbitrary expression = the difference of that expression between two pixels of the invocation's 2 x 2 quad.  The
// raster harness (oracle/wgsl/raster_prepass.py) evaluates the fragment function at the quad's pixels first (mode 1: every derivative
// call records its argument, in call order, into the slot of the pixel being evaluated: 0 / 1 = left / right pixel of the row, 2 / 3 =
// top / bottom pixel of the column) and then at the pixel itself (mode 2: call k returns the recorded difference) — fine derivatives.
struct wgsl_derivatives { int mode = 0, slot = 0, n = 0, nx = 0, ny = 0; f32 rec[4][32]; };
inline wgsl_derivatives& wgsl_derivative_state() { static thread_local wgsl_derivatives d; return d; }
inline f32 wgsl_derivative(f32 v, int lo, int hi) {
    wgsl_derivatives& D = wgsl_derivative_state();
    if (D.mode == 1) { if (D.n < 32) D.rec[D.slot][D.n] = v; D.n += 1; return 0.0f; }
    if (D.mode == 2) { const int k = D.nx++; return k < 32 ? D.rec[hi][k] - D.rec[lo][k] : 0.0f; }
    return 0.0f;
}
inline f32 dpdx(f32 v) { return wgsl_derivative(v, 0, 1); }
inline f32 dpdy(f32 v) { return wgsl_derivative(v, 2, 3); }

// ------------------------------------------------------------------------------------------------ dispatch
struct wgsl_ids { vec3<u32> global, local, group, num_groups; u32 local_index = 0; };
inline wgsl_ids& wgsl_tls() { static thread_local wgsl_ids ids; return ids; }
inline vec3<u32> wgsl_global_id() { return wgsl_tls().global; }
inline vec3<u32> wgsl_local_id() { return wgsl_tls().local; }
inline vec3<u32> wgsl_group_id() { return wgsl_tls().group; }
inline vec3<u32> wgsl_num_groups() { return wgsl_tls().num_groups; }
inline u32 wgsl_local_index() { return wgsl_tls().local_index; }
static pthread_barrier_t* wgsl_barrier = nullptr;
static int wgsl_parallel = 0;
static const u32* wgsl_order = nullptr;      // optional explicit invocation order: (x, y) global ids, every invocation of the dispatch once
static size_t wgsl_order_count = 0;
extern "C" __attribute__((used)) void wgsl_set_parallel(int on) { wgsl_parallel = on; }
extern "C" __attribute__((used)) void wgsl_set_order(const u32* xy, size_t count) { wgsl_order = xy; wgsl_order_count = count; }
inline void workgroupBarrier() { if (wgsl_barrier) pthread_barrier_wait(wgsl_barrier); }
inline void storageBarrier() { if (wgsl_barrier) pthread_barrier_wait(wgsl_barrier); }

inline void wgsl_set_ids(unsigned gx, unsigned gy, unsigned gz, unsigned wx, unsigned wy, unsigned wz, unsigned ngx, unsigned ngy, unsigned ngz,
                         unsigned lx, unsigned ly, unsigned lz) {
    wgsl_ids& t = wgsl_tls();
    t.group = vec3<u32>(gx, gy, gz); t.local = vec3<u32>(lx, ly, lz); t.num_groups = vec3<u32>(ngx, ngy, ngz);
    t.global = vec3<u32>(gx * wx + lx, gy * wy + ly, gz * wz + lz);
    t.local_index = lx + wx * (ly + wy * lz);
}
// Entry points that call workgroupBarrier() (spatial_reuse): the invocations of a workgroup are real threads meeting at a pthread
// barrier, workgroups one after the other.  All others: see below.
template <bool COOPERATIVE, class F> void wgsl_dispatch(unsigned ngx, unsigned ngy, unsigned ngz, unsigned wx, unsigned wy, unsigned wz, F body) {
  if (COOPERATIVE) {
    const unsigned n = wx * wy * wz;
    pthread_barrier_t inner, outer;
    pthread_barrier_init(&inner, nullptr, n);
    pthread_barrier_init(&outer, nullptr, n);
    wgsl_barrier = &inner;
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < n; ++t)
        threads.emplace_back([=, &outer]() {
            const unsigned lx = t % wx, ly = (t / wx) % wy, lz = t / (wx * wy);
            for (unsigned gz = 0; gz < ngz; ++gz)
                for (unsigned gy = 0; gy < ngy; ++gy)
                    for (unsigned gx = 0; gx < ngx; ++gx) {
                        wgsl_set_ids(gx, gy, gz, wx, wy, wz, ngx, ngy, ngz, lx, ly, lz);
                        body();
                        pthread_barrier_wait(&outer);          // workgroup-shared variables are reused by the next workgroup
                    }
        });
    for (auto& th : threads) th.join();
    wgsl_barrier = nullptr;
    pthread_barrier_destroy(&inner);
    pthread_barrier_destroy(&outer);
    return;
  }
    // invocations one after the other in raster order of their global ids (row by row), or in the order the caller gives.  The
    // shader text has one data race: direct_lit / indirect_lit_ambient scatter reservoirs to reprojected pixels
    // (light.wgsl:1094,1201,1458) into the buffer in which background pixels store their own (light.wgsl:1063,1282), and several
    // writers can name one target.  Any winner is a valid execution of the reference; raster order makes "the last writer in raster
    // order wins" — the rule the oracle and the CUDA path fix.  wgsl_parallel = 1 spreads rows over cores (racy: timing only).
    const long long rows = (long long)ngy * wy * ngz * wz;
    const unsigned width = ngx * wx;
    if (wgsl_order && wgsl_order_count > 0 && ngz * wz == 1) {      // the caller's order (any order is a valid execution; a subset = debugging)
        for (size_t k = 0; k < wgsl_order_count; ++k) {
            const unsigned x = wgsl_order[2 * k], y = wgsl_order[2 * k + 1];
            wgsl_set_ids(x / wx, y / wy, 0, wx, wy, wz, ngx, ngy, ngz, x % wx, y % wy, 0);
            body();
        }
        return;
    }
#pragma omp parallel for schedule(dynamic, 1) if (wgsl_parallel)
    for (long long row = 0; row < rows; ++row) {
        const unsigned y = (unsigned)(row % ((long long)ngy * wy)), z = (unsigned)(row / ((long long)ngy * wy));
        for (unsigned x = 0; x < width; ++x) {
            wgsl_set_ids(x / wx, y / wy, z / wz, wx, wy, wz, ngx, ngy, ngz, x % wx, y % wy, z % wz);
            body();
        }
    }
}

}  // namespace wgsl
