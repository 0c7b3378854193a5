// kernels_scene.cu — SURVEY 8(f) rank 2: the per-frame half of the scene rebuilt ON THE DEVICE.
//
// The reference rebuilds, on the CPU and whenever anything moves (src/mesh_material/instance.rs:352-437): every instance's world
// AABB, model and inverse-transpose matrix (:286-323), the TLAS over the instances (bvh 0.7.1 BVH::build + flatten_custom, :365-371),
// every emissive's bounding sphere and surface area (:399-418) and the BVH over the emissives (:420-427) — and then re-uploads all
// of it (:430-437); its README lists "asynchronous building of acceleration structures" as to do.  hk_scene_update_instances takes
// those arrays from a host that has done this work (host/hikari.cpp); the kernels below do the work themselves from what actually
// changed — one model matrix per instance — and leave the results where the light kernels read them, stream-ordered, without a host
// round trip:
//   k_scene_instances    per instance: world AABB, inverse transpose, previous model + "moved" flag, compact traversal record
//   kc_build_flat_bvh    one CTA: bvh 0.7.1's bucketed SAH build + flatten_custom, one WARP per tree node, level by level
//   kc_scene_emissives   one warp per emissive: bounding sphere, surface area (areas in parallel, summed in the reference's order)
// The result is the reference's, bit for bit (tests/test_gpu_scene_update.py compares every record with the host mirror's):
//   * the arithmetic is host/hikari.cpp's, statement for statement, compiled without FMA contraction;
//   * a node's split (axis, buckets, SAH costs) depends on min / max / counts over the node's shapes, which no order of evaluation
//     changes (boxes never hold -0: every bound is a sum whose exact-zero result rounds to +0), and the one thing that is
//     order-sensitive — the order of the shapes inside a child, bucket by bucket, stable — is reproduced by a ballot-ranked
//     counting sort;
//   * a node's place in the flat array follows from how many tree nodes and leaves precede it in pre-order: a subtree of m shapes
//     has 2m - 1 tree nodes and writes 3m - 1 records, so both children's places are known when the parent splits.
// What this path does NOT rebuild: alias tables (their cache rule — rebuild when the scale moved by more than 0.01,
// instance.rs:385-397 — is the caller's to check: hikari::MeshMaterialWorld::prepare_instance_transforms) and anything that changes
// the SET of instances, meshes or materials; those go through hk_scene_update_instances / hk_scene_upload.
#include "hk_device.cuh"
#include "hk_kernels.h"

namespace hkd {

namespace {

constexpr float SCENE_BVH_EPSILON = 0.00001f;   // bvh 0.7.1 EPSILON
constexpr int SCENE_BUCKETS = 6;                // bvh 0.7.1 NUM_BUCKETS
constexpr uint32_t SCENE_LEAF = 0x80000000u;

// std::min / std::max as the host mirror uses them (first argument wins ties)
__device__ __forceinline__ float smin(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float smax(float a, float b) { return a < b ? b : a; }

struct SBox {
    float mn[3], mx[3];
    __device__ __forceinline__ void clear() {
        mn[0] = mn[1] = mn[2] = __uint_as_float(0x7f800000u);
        mx[0] = mx[1] = mx[2] = __uint_as_float(0xff800000u);
    }
    __device__ __forceinline__ void join(const SBox& o) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { mn[k] = smin(mn[k], o.mn[k]); mx[k] = smax(mx[k], o.mx[k]); }
    }
    __device__ __forceinline__ float surface_area() const {
        const float sx = mx[0] - mn[0], sy = mx[1] - mn[1], sz = mx[2] - mn[2];
        return 2.0f * (sx * sy + sx * sz + sy * sz);
    }
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = smin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = smax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void warp_join(SBox& b) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { b.mn[k] = warp_min(b.mn[k]); b.mx[k] = warp_max(b.mx[k]); }
}

__device__ __forceinline__ void store_node(hk_node* out, uint32_t at, const SBox& b, uint32_t entry, uint32_t exit) {
    float4* p = reinterpret_cast<float4*>(out + at);
    p[0] = make_float4(b.mn[0], b.mn[1], b.mn[2], __uint_as_float(entry));
    p[1] = make_float4(b.mx[0], b.mx[1], b.mx[2], __uint_as_float(exit));
}

// glam Mat4::transform_point3 / transform_vector3 as restated in host/hikari.cpp (column-major, no perspective divide)
__device__ __forceinline__ void xf_point3(const float* m, const float* p, float* out) {
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = m[k] * p[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = m[4 + k] * p[1] + r[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = m[8 + k] * p[2] + r[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = m[12 + k] + r[k];
}
__device__ __forceinline__ void xf_vector3(const float* m, const float* p, float* out) {
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = m[k] * p[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = m[4 + k] * p[1] + r[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = m[8 + k] * p[2] + r[k];
}
__device__ __forceinline__ float len3(const float* a) { return sqrtf((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }

// glam Mat4::inverse, scalar path (host/hikari.cpp mat4_inverse)
__device__ void mat4_inverse_dev(const float* m, float* out) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3];
    const float m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const float m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11];
    const float m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    const float coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13, coef03 = m12 * m23 - m22 * m13;
    const float coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13, coef07 = m11 * m23 - m21 * m13;
    const float coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12, coef11 = m11 * m22 - m21 * m12;
    const float coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13, coef15 = m10 * m23 - m20 * m13;
    const float coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12, coef19 = m10 * m22 - m20 * m12;
    const float coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11, coef23 = m10 * m21 - m20 * m11;
    const float fac0[4] = {coef00, coef00, coef02, coef03}, fac1[4] = {coef04, coef04, coef06, coef07};
    const float fac2[4] = {coef08, coef08, coef10, coef11}, fac3[4] = {coef12, coef12, coef14, coef15};
    const float fac4[4] = {coef16, coef16, coef18, coef19}, fac5[4] = {coef20, coef20, coef22, coef23};
    const float vec0[4] = {m10, m00, m00, m00}, vec1[4] = {m11, m01, m01, m01};
    const float vec2[4] = {m12, m02, m02, m02}, vec3_[4] = {m13, m03, m03, m03};
    const float sign_a[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sign_b[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
    float inv[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float inv0 = vec1[k] * fac0[k] - vec2[k] * fac1[k] + vec3_[k] * fac2[k];
        const float inv1 = vec0[k] * fac0[k] - vec2[k] * fac3[k] + vec3_[k] * fac4[k];
        const float inv2 = vec0[k] * fac1[k] - vec1[k] * fac3[k] + vec3_[k] * fac5[k];
        const float inv3 = vec0[k] * fac2[k] - vec1[k] * fac4[k] + vec2[k] * fac5[k];
        inv[0][k] = inv0 * sign_a[k]; inv[1][k] = inv1 * sign_b[k]; inv[2][k] = inv2 * sign_a[k]; inv[3][k] = inv3 * sign_b[k];
    }
    const float d0 = m00 * inv[0][0], d1 = m01 * inv[1][0], d2 = m02 * inv[2][0], d3 = m03 * inv[3][0];
    const float dot1 = ((d0 + d1) + d2) + d3;
    const float rcp_det = 1.0f / dot1;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) out[4 * c + k] = inv[c][k] * rcp_det;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ instances
// instance.rs:286-323 for instance i: the mesh's Aabb (centre, half extents) carried through the new GlobalTransform.
// `previous` = PreviousMeshUniform::transform (instance.rs:111-128); nullptr = the model the record held until now.
__global__ void __launch_bounds__(128) k_scene_instances(uint32_t n, const float4* __restrict__ models, const float4* __restrict__ previous,
                                                         const float* __restrict__ mesh_aabbs, hk_instance* instances, hk_instance_trav* trav,
                                                         float4* previous_out, uint32_t* moved_out, float4* box_lo, float4* box_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m[16], p[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 v = models[4u * i + c];
        m[4 * c] = v.x; m[4 * c + 1] = v.y; m[4 * c + 2] = v.z; m[4 * c + 3] = v.w;
    }
    const float4* old = previous ? previous + 4u * i : reinterpret_cast<const float4*>(instances[i].model);
    uint32_t moved = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 v = old[c];
        p[4 * c] = v.x; p[4 * c + 1] = v.y; p[4 * c + 2] = v.z; p[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) moved |= (__float_as_uint(p[k]) != __float_as_uint(m[k])) ? 1u : 0u;   // memcmp, as the host path
    const float center[3] = {mesh_aabbs[6u * i], mesh_aabbs[6u * i + 1], mesh_aabbs[6u * i + 2]};
    const float half[3] = {mesh_aabbs[6u * i + 3], mesh_aabbs[6u * i + 4], mesh_aabbs[6u * i + 5]};
    float c[3];
    xf_point3(m, center, c);
    float lo[3] = {0.0f, 0.0f, 0.0f}, hi[3] = {0.0f, 0.0f, 0.0f};       // instance.rs:298-303 starts from ZERO
#pragma unroll
    for (int index = 0; index < 8; ++index) {
        const float vtx[3] = {half[0] * (float)(2 * (index & 1) - 1), half[1] * (float)(2 * ((index >> 1) & 1) - 1),
                              half[2] * (float)(2 * ((index >> 2) & 1) - 1)};
        float t[3];
        xf_vector3(m, vtx, t);
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = smin(lo[k], t[k]); hi[k] = smax(hi[k], t[k]); }
    }
    float inv[16];
    mat4_inverse_dev(m, inv);
    hk_instance* rec = instances + i;
    float4* r4 = reinterpret_cast<float4*>(rec);
    const uint32_t material = rec->material, node_index = rec->node_index;
    const float mn[3] = {lo[0] + c[0], lo[1] + c[1], lo[2] + c[2]}, mx[3] = {hi[0] + c[0], hi[1] + c[1], hi[2] + c[2]};
    r4[0] = make_float4(mn[0], mn[1], mn[2], __uint_as_float(material));
    r4[1] = make_float4(mx[0], mx[1], mx[2], __uint_as_float(node_index));
    float4* t4 = reinterpret_cast<float4*>(trav + i);
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        r4[2 + col] = make_float4(m[4 * col], m[4 * col + 1], m[4 * col + 2], m[4 * col + 3]);
        // inverse().transpose(): element (row, col) of the transpose = element (col, row) of the inverse
        const float4 it = make_float4(inv[col], inv[4 + col], inv[8 + col], inv[12 + col]);
        r4[6 + col] = it;
        t4[col] = it;
        previous_out[4u * i + col] = make_float4(p[4 * col], p[4 * col + 1], p[4 * col + 2], p[4 * col + 3]);
    }
    moved_out[i] = moved;
    box_lo[i] = make_float4(mn[0], mn[1], mn[2], 0.0f);
    box_hi[i] = make_float4(mx[0], mx[1], mx[2], 0.0f);
}

// ------------------------------------------------------------------------------------------------ SAH build + flatten
// One tree node of the level being split.  `t` = its index among the tree nodes in pre-order (what bvh calls the node index and
// writes back through set_bh_node_index), `leaves_before` = shapes that precede it in pre-order; its navigator record sits at
// t - 1 + leaves_before (the root has none).
struct BuildSeg {
    uint32_t begin, count, t, leaves_before;
    float lo[3], hi[3];          // the box the PARENT computed for this child (l_box / r_box of BVH::build)
};
static_assert(sizeof(BuildSeg) == 40, "scratch sizing in context.cu");

__global__ void __launch_bounds__(256) kc_build_flat_bvh(uint32_t n, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                                                         float4* center, uint32_t* idx_a, uint32_t* idx_b, BuildSeg* seg_a, BuildSeg* seg_b,
                                                         hk_node* out, uint8_t* index_base, uint32_t index_stride) {
    __shared__ uint32_t s_cur_count, s_next_count;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, warps = blockDim.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    if (n == 0u) return;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        idx_a[i] = i;
        const float4 lo = box_lo[i], hi = box_hi[i];
        center[i] = make_float4(lo.x + (hi.x - lo.x) / 2.0f, lo.y + (hi.y - lo.y) / 2.0f, lo.z + (hi.z - lo.z) / 2.0f, 0.0f);   // AABB::center
    }
    if (tid == 0u) {
        s_next_count = 0u;
        if (n == 1u) {      // a single shape: one leaf record without a navigator
            SBox e; e.clear();
            store_node(out, 0u, e, SCENE_LEAF | 0u, 1u);
            *reinterpret_cast<uint32_t*>(index_base) = 0u;
            s_cur_count = 0u;
        } else {
            BuildSeg root;
            root.begin = 0u; root.count = n; root.t = 0u; root.leaves_before = 0u;
            for (int k = 0; k < 3; ++k) { root.lo[k] = 0.0f; root.hi[k] = 0.0f; }
            seg_a[0] = root;
            s_cur_count = 1u;
        }
    }
    __syncthreads();
    BuildSeg* cur = seg_a; BuildSeg* nxt = seg_b;
    uint32_t* idx_cur = idx_a; uint32_t* idx_nxt = idx_b;
    for (;;) {
        const uint32_t cnt = s_cur_count;
        if (cnt == 0u) break;
        for (uint32_t s = warp; s < cnt; s += warps) {
            const BuildSeg sg = cur[s];
            const uint32_t at = sg.t - 1u + sg.leaves_before;        // navigator record (not for the root)
            SBox own;
            for (int k = 0; k < 3; ++k) { own.mn[k] = sg.lo[k]; own.mx[k] = sg.hi[k]; }
            if (sg.count == 1u) {
                if (lane == 0u) {
                    const uint32_t shape = idx_cur[sg.begin];
                    SBox e; e.clear();
                    store_node(out, at, own, at + 1u, at + 2u);
                    store_node(out, at + 1u, e, SCENE_LEAF | shape, at + 2u);
                    *reinterpret_cast<uint32_t*>(index_base + (size_t)shape * index_stride) = sg.t;
                }
                continue;
            }
            if (sg.t != 0u && lane == 0u) store_node(out, at, own, at + 1u, at + 3u * sg.count - 1u);
            // bounds of the shapes' boxes and of their centres
            SBox ab, cb;
            ab.clear(); cb.clear();
            for (uint32_t j = lane; j < sg.count; j += 32u) {
                const uint32_t shape = idx_cur[sg.begin + j];
                const float4 lo = box_lo[shape], hi = box_hi[shape], c = center[shape];
                ab.mn[0] = smin(ab.mn[0], lo.x); ab.mn[1] = smin(ab.mn[1], lo.y); ab.mn[2] = smin(ab.mn[2], lo.z);
                ab.mx[0] = smax(ab.mx[0], hi.x); ab.mx[1] = smax(ab.mx[1], hi.y); ab.mx[2] = smax(ab.mx[2], hi.z);
                cb.mn[0] = smin(cb.mn[0], c.x); cb.mn[1] = smin(cb.mn[1], c.y); cb.mn[2] = smin(cb.mn[2], c.z);
                cb.mx[0] = smax(cb.mx[0], c.x); cb.mx[1] = smax(cb.mx[1], c.y); cb.mx[2] = smax(cb.mx[2], c.z);
            }
            warp_join(ab); warp_join(cb);
            const float sx = cb.mx[0] - cb.mn[0], sy = cb.mx[1] - cb.mn[1], sz = cb.mx[2] - cb.mn[2];
            const int axis = (sx > sy && sx > sz) ? 0 : (sy > sz ? 1 : 2);       // AABB::largest_axis
            const float split_axis_size = axis == 0 ? sx : (axis == 1 ? sy : sz);
            const float axis_min = axis == 0 ? cb.mn[0] : (axis == 1 ? cb.mn[1] : cb.mn[2]);
            SBox l_box, r_box;
            l_box.clear(); r_box.clear();
            uint32_t nl;
            if (split_axis_size < SCENE_BVH_EPSILON) {
                // the centres coincide along every axis: halves in the order the shapes have
                nl = sg.count / 2u;
                for (uint32_t j = lane; j < sg.count; j += 32u) {
                    const uint32_t shape = idx_cur[sg.begin + j];
                    const float4 lo = box_lo[shape], hi = box_hi[shape];
                    SBox b;
                    b.mn[0] = lo.x; b.mn[1] = lo.y; b.mn[2] = lo.z; b.mx[0] = hi.x; b.mx[1] = hi.y; b.mx[2] = hi.z;
                    if (j < nl) l_box.join(b); else r_box.join(b);
                    idx_nxt[sg.begin + j] = shape;
                }
                warp_join(l_box); warp_join(r_box);
            } else {
                SBox bb[SCENE_BUCKETS];
                uint32_t bn[SCENE_BUCKETS];
#pragma unroll
                for (int b = 0; b < SCENE_BUCKETS; ++b) { bb[b].clear(); bn[b] = 0u; }
                for (uint32_t j = lane; j < sg.count; j += 32u) {
                    const uint32_t shape = idx_cur[sg.begin + j];
                    const float4 lo = box_lo[shape], hi = box_hi[shape], c = center[shape];
                    const float ca = axis == 0 ? c.x : (axis == 1 ? c.y : c.z);
                    const float rel = (ca - axis_min) / split_axis_size;
                    const uint32_t bucket = (uint32_t)(rel * ((float)SCENE_BUCKETS - 0.01f));
                    SBox b;
                    b.mn[0] = lo.x; b.mn[1] = lo.y; b.mn[2] = lo.z; b.mx[0] = hi.x; b.mx[1] = hi.y; b.mx[2] = hi.z;
#pragma unroll
                    for (int q = 0; q < SCENE_BUCKETS; ++q)
                        if (bucket == (uint32_t)q) { bb[q].join(b); bn[q] += 1u; }
                }
#pragma unroll
                for (int b = 0; b < SCENE_BUCKETS; ++b) { warp_join(bb[b]); bn[b] = warp_sum(bn[b]); }
                // the split with the least SAH cost (every lane evaluates the same five candidates)
                int min_bucket = 0;
                float min_cost = __uint_as_float(0x7f800000u);
                const float parent_area = ab.surface_area();
#pragma unroll
                for (int i = 0; i < SCENE_BUCKETS - 1; ++i) {
                    SBox cl, cr;
                    cl.clear(); cr.clear();
                    uint32_t nlc = 0u, nrc = 0u;
#pragma unroll
                    for (int b = 0; b < SCENE_BUCKETS; ++b) {
                        if (b <= i) { cl.join(bb[b]); nlc += bn[b]; } else { cr.join(bb[b]); nrc += bn[b]; }
                    }
                    const float cost = ((float)nlc * cl.surface_area() + (float)nrc * cr.surface_area()) / parent_area;
                    if (cost < min_cost) { min_bucket = i; min_cost = cost; l_box = cl; r_box = cr; }
                }
                // children = the buckets' shapes, bucket by bucket, each bucket in the order the shapes had: a stable counting sort
                uint32_t off[SCENE_BUCKETS];
                uint32_t acc = 0u;
                nl = 0u;
#pragma unroll
                for (int b = 0; b < SCENE_BUCKETS; ++b) {
                    off[b] = acc; acc += bn[b];
                    if (b <= min_bucket) nl += bn[b];
                }
                for (uint32_t base = 0u; base < sg.count; base += 32u) {
                    const uint32_t j = base + lane;
                    const bool valid = j < sg.count;
                    uint32_t shape = 0u, bucket = 0xFFFFFFFFu;
                    if (valid) {
                        shape = idx_cur[sg.begin + j];
                        const float4 c = center[shape];
                        const float ca = axis == 0 ? c.x : (axis == 1 ? c.y : c.z);
                        const float rel = (ca - axis_min) / split_axis_size;
                        bucket = (uint32_t)(rel * ((float)SCENE_BUCKETS - 0.01f));
                    }
#pragma unroll
                    for (int b = 0; b < SCENE_BUCKETS; ++b) {
                        const uint32_t mask = __ballot_sync(0xffffffffu, valid && bucket == (uint32_t)b);
                        if (valid && bucket == (uint32_t)b) idx_nxt[sg.begin + off[b] + (uint32_t)__popc(mask & lt)] = shape;
                        off[b] += (uint32_t)__popc(mask);
                    }
                }
            }
            if (lane == 0u) {
                const uint32_t slot = atomicAdd(&s_next_count, 2u);
                BuildSeg l, r;
                l.begin = sg.begin; l.count = nl; l.t = sg.t + 1u; l.leaves_before = sg.leaves_before;
                r.begin = sg.begin + nl; r.count = sg.count - nl; r.t = sg.t + 2u * nl; r.leaves_before = sg.leaves_before + nl;
                for (int k = 0; k < 3; ++k) { l.lo[k] = l_box.mn[k]; l.hi[k] = l_box.mx[k]; r.lo[k] = r_box.mn[k]; r.hi[k] = r_box.mx[k]; }
                nxt[slot] = l; nxt[slot + 1u] = r;
            }
        }
        __syncthreads();
        if (tid == 0u) { s_cur_count = s_next_count; s_next_count = 0u; }
        BuildSeg* ts = cur; cur = nxt; nxt = ts;
        uint32_t* ti = idx_cur; idx_cur = idx_nxt; idx_nxt = ti;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ emissives
// instance.rs:399-418 for emissive e (its instance, alias-table slice and material are those of the last full upload): bounding
// sphere from the instance's new AABB, surface area = sum of the transformed triangles' areas (mod.rs:318-328) in primitive order.
__global__ void __launch_bounds__(128) kc_scene_emissives(uint32_t ne, hk_emissive* emissives, const hk_instance* __restrict__ instances,
                                                          const hk_material* __restrict__ materials, const hk_primitive* __restrict__ primitives,
                                                          const hk_vertex* __restrict__ vertices, float4* box_lo, float4* box_hi) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= ne) return;                                   // whole warps leave together
    hk_emissive* em = emissives + e;
    const hk_instance* inst = instances + em->instance;
    float m[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m[k] = inst->model[k];
    const uint32_t prim0 = inst->mesh.primitive, vert0 = inst->mesh.vertex, count = em->alias_table_count;   // one alias entry per triangle
    float surface_area = 0.0f;
    for (uint32_t base = 0u; base < count; base += 32u) {
        const uint32_t i = base + lane;
        float area = 0.0f;
        if (i < count) {
            const hk_primitive* pr = primitives + prim0 + i;
            float v[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k) xf_point3(m, vertices[vert0 + pr->vertices[k].index].position, v[k]);
            const float a[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]};
            const float b[3] = {v[2][0] - v[0][0], v[2][1] - v[0][1], v[2][2] - v[0][2]};
            const float c[3] = {a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]};   // glam cross
            area = 0.5f * fabsf(len3(c));
        }
        const uint32_t in_chunk = count - base < 32u ? count - base : 32u;
        for (uint32_t l = 0u; l < 32u; ++l) {              // the reference's left-to-right sum, every lane the same
            const float a_l = __shfl_sync(0xffffffffu, area, (int)l);
            if (l < in_chunk) surface_area += a_l;
        }
    }
    if (lane == 0u) {
        const float* ec = materials[inst->material].emissive;
        const float e4[4] = {ec[0], ec[1], ec[2], ec[3]};
        const float intensity = 255.0f * e4[3] * len3(e4);
        float pos[3], ext[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { pos[k] = 0.5f * (inst->max[k] + inst->min[k]); ext[k] = inst->max[k] - inst->min[k]; }
        const float radius = 0.5f * len3(ext) + sqrtf(intensity);
#pragma unroll
        for (int k = 0; k < 4; ++k) em->emissive[k] = e4[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) em->position[k] = pos[k];
        em->radius = radius;
        em->surface_area = surface_area;
        box_lo[e] = make_float4(pos[0] - radius, pos[1] - radius, pos[2] - radius, 0.0f);
        box_hi[e] = make_float4(pos[0] + radius, pos[1] + radius, pos[2] + radius, 0.0f);
    }
}

}  // namespace hkd

// ------------------------------------------------------------------------------------------------ launchers
using namespace hkd;
void hk_launch_scene_instances(uint32_t n, const float4* models, const float4* previous, const float* mesh_aabbs, hk_instance* instances,
                               hkd::hk_instance_trav* trav, float4* previous_out, uint32_t* moved_out, float4* box_lo, float4* box_hi, cudaStream_t st) {
    if (n == 0u) return;
    k_scene_instances<<<(n + 127u) / 128u, 128, 0, st>>>(n, models, previous, mesh_aabbs, instances, trav, previous_out, moved_out, box_lo, box_hi);
}
void hk_launch_build_flat_bvh(uint32_t n, const float4* box_lo, const float4* box_hi, void* scratch, hk_node* out, void* index_base,
                              uint32_t index_stride, cudaStream_t st) {
    if (n == 0u) return;
    // scratch: centre float4[n] | idx uint32[2][n] | segments BuildSeg[2][n]   (hk_scene_bvh_scratch_bytes)
    uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
    float4* center = reinterpret_cast<float4*>(p); p += sizeof(float4) * (size_t)n;
    uint32_t* idx_a = reinterpret_cast<uint32_t*>(p); p += sizeof(uint32_t) * (size_t)n;
    uint32_t* idx_b = reinterpret_cast<uint32_t*>(p); p += sizeof(uint32_t) * (size_t)n;
    p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 15u) & ~(uintptr_t)15u);
    hkd::BuildSeg* seg_a = reinterpret_cast<hkd::BuildSeg*>(p); p += sizeof(hkd::BuildSeg) * (size_t)n;
    hkd::BuildSeg* seg_b = reinterpret_cast<hkd::BuildSeg*>(p);
    kc_build_flat_bvh<<<1, 256, 0, st>>>(n, box_lo, box_hi, center, idx_a, idx_b, seg_a, seg_b, out, reinterpret_cast<uint8_t*>(index_base), index_stride);
}
size_t hk_scene_bvh_scratch_bytes(uint32_t n) {
    return sizeof(float4) * (size_t)n + 2u * sizeof(uint32_t) * (size_t)n + 16u + 2u * sizeof(hkd::BuildSeg) * (size_t)n;
}
void hk_launch_scene_emissives(uint32_t ne, hk_emissive* emissives, const hk_instance* instances, const hk_material* materials,
                               const hk_primitive* primitives, const hk_vertex* vertices, float4* box_lo, float4* box_hi, cudaStream_t st) {
    if (ne == 0u) return;
    kc_scene_emissives<<<(ne + 3u) / 4u, 128, 0, st>>>(ne, emissives, instances, materials, primitives, vertices, box_lo, box_hi);
}
