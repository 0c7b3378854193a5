// hk_pool.cuh — the per-CTA ray pool of the cooperative light kernels (kc_*).
//
// Why: the per-pixel light kernels of round 1 traced one ray per lane, inlined in the shading code.  ncu showed 11.6 active
// lanes per instruction in k_indirect: the skip-link walk has a data-dependent trip count (mean / max over an 8x4 tile 0.54 on
// cornell, 0.25 in the city), so a warp waits for its longest ray while the rest of its lanes idle (experiment 4).
//
// Design (Aila-Laine style dynamic fetch, per CTA, in shared memory):
//   * a CTA of POOL_THREADS threads owns POOL_THREADS pixels, one per thread; all SHADING state of a pixel stays in its owner's
//     registers for the whole kernel;
//   * whenever the pixels need rays traced, every owner writes its ray into its slot of the pool (SoA arrays in shared memory)
//     and the slot index into a queue, compacted with __ballot_sync / __popc so that the queue is dense;
//   * pool_traverse(): only POOL_TRAVERSE_WARPS of the CTA's warps walk the BVH.  Each lane pulls the next queued ray the moment
//     its current one terminates (warp-aggregated fetch: one ballot of the idle lanes, one shared-memory atomicAdd by their
//     leader, ranks by popc), so traversal lanes stay full until the pool runs dry; with P pixels per CTA and T traversal lanes a
//     lane walks P / T rays back to back, which is what averages the trip-count variance out.  The other warps wait at the CTA
//     barrier and cost no issue slots; their registers ARE the storage of the pixels' shading state;
//   * results (u, v, distance, instance, primitive) come back through the ray's slot; the owner picks them up after the barrier.
//   The walk of ONE ray is exactly the reference's (same records, same visit order, same strict '<' updates, same early-outs), so
//   results are bit-identical no matter which lane walks it.
//   * scene records that every ray touches are staged ONCE per CTA into shared memory with TMA bulk copies
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS: UBLKCP): the TLAS records and a compact per-instance
//     traversal record (object-space transform + mesh index, 80 of the instance's 176 bytes); for small scenes (cornell: 3.3 KB of
//     BLAS records, 1.5 KB of triangles) the BLAS records and triangles as well, so that such a scene is walked entirely out of
//     shared memory.  What does not fit stays in global memory (L1 / L2); the walk uses generic pointers and does not care.
#pragma once
#include "hk_device.cuh"

#ifndef HK_POOL_TRAVERSE_WARPS
#define HK_POOL_TRAVERSE_WARPS 4
#endif
#ifndef HK_NOINLINE
#define HK_NOINLINE __noinline__
#endif
#ifndef HK_POOL_MINB
#define HK_POOL_MINB 3
#endif

namespace hkd {

constexpr int POOL_THREADS = 256;                   // 16 x 16 pixels, 8 warps of 8 x 4
constexpr int POOL_TILE_W = 16, POOL_TILE_H = 16;
constexpr int POOL_TRAVERSE_WARPS = HK_POOL_TRAVERSE_WARPS;
constexpr uint32_t FULL_MASK = 0xffffffffu;
constexpr uint32_t RAY_KIND_TLAS = 0u, RAY_KIND_BLAS = 1u;

// ------------------------------------------------------------------------------------- TMA bulk copy + mbarrier
#ifdef HK_EMU
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { *reinterpret_cast<volatile uint32_t*>(bar) = 0u; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_complete_emulated(uint64_t* bar) { *reinterpret_cast<volatile uint32_t*>(bar) = 1u; }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t) { emu_wait_changed(reinterpret_cast<volatile uint32_t*>(bar), 0u); }
__device__ __forceinline__ uint32_t lanemask_lt() { return emu_lanemask_lt(); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(arrivals), "r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy by the TMA unit; completion is signalled on `bar` as `bytes` of transaction count
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_complete_emulated(uint64_t*) {}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "HK_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra HK_MBAR_DONE;\n\t"
        "bra HK_MBAR_WAIT;\n\t"
        "HK_MBAR_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
#endif

// ------------------------------------------------------------------------------------------------- the pool
constexpr int STAGE_F4 = HK_STAGE_F4;   // 16 KB of staged scene records per CTA

struct PoolShared {
    // ray slots, one per thread of the CTA (slot = owner's thread index).  A slot's result overwrites its ray:
    //   ray:  ox oy oz | dx dy dz | tmax early | arg (exclude instance, or the instance of a BLAS-only ray) | kind
    //   hit:  u  v  distance | instance primitive (as bits in dx, dy)
    float ox[POOL_THREADS], oy[POOL_THREADS], oz[POOL_THREADS];
    float dx[POOL_THREADS], dy[POOL_THREADS], dz[POOL_THREADS];
    float tmax[POOL_THREADS], early[POOL_THREADS];
    uint32_t arg[POOL_THREADS], kind[POOL_THREADS];
    uint16_t queue[POOL_THREADS];
    int q_count[2], q_next[2];       // double-buffered by phase parity: the next phase is filled while nobody reads the old counters
    uint64_t stage_bar;              // mbarrier of the scene staging copies
    float4 stage[STAGE_F4];
};

// per-thread view of the scene for the walk: generic pointers into shared memory where staged, into global memory otherwise
struct WalkScene {
    const float4* tlas;        // 2 per record
    const float4* itrav;       // 5 per instance: transpose(inverse_transpose_model) rows as stored (4) | mesh index (1)
    const float4* blas;        // 2 per record, whole asset_nodes buffer
    const float4* prims;       // 3 per triangle, whole primitives buffer
    const hk_instance* instances;
    uint32_t tlas_count, leaf_boxes_match;
};

// Thread 0 starts the staging copies; every thread waits for them in pool_stage_wait() before its first walk.
__device__ __forceinline__ void pool_stage_begin(PoolShared& S, const DeviceScene& sc, const StagePlan& plan) {
    if (threadIdx.x == 0) {
        S.q_count[0] = S.q_count[1] = 0; S.q_next[0] = S.q_next[1] = 0;
        mbar_init(&S.stage_bar, 1u);
        const uint32_t b_tlas = plan.tlas_count * 32u, b_itrav = plan.itrav_count * 80u, b_blas = plan.blas_count * 32u, b_prim = plan.prim_count * 48u;
        mbar_expect_tx(&S.stage_bar, b_tlas + b_itrav + b_blas + b_prim);
        if (b_tlas) bulk_copy_g2s(&S.stage[plan.tlas_f4], sc.instance_nodes, b_tlas, &S.stage_bar);
        if (b_itrav) bulk_copy_g2s(&S.stage[plan.itrav_f4], sc.instance_trav, b_itrav, &S.stage_bar);
        if (b_blas) bulk_copy_g2s(&S.stage[plan.blas_f4], sc.asset_nodes, b_blas, &S.stage_bar);
        if (b_prim) bulk_copy_g2s(&S.stage[plan.prim_f4], sc.primitives, b_prim, &S.stage_bar);
        mbar_complete_emulated(&S.stage_bar);
    }
}
__device__ __forceinline__ WalkScene pool_stage_wait(PoolShared& S, const DeviceScene& sc, const StagePlan& plan) {
    mbar_wait(&S.stage_bar, 0u);
    WalkScene w;
    w.tlas = plan.tlas_count ? &S.stage[plan.tlas_f4] : reinterpret_cast<const float4*>(sc.instance_nodes);
    w.itrav = plan.itrav_count ? &S.stage[plan.itrav_f4] : reinterpret_cast<const float4*>(sc.instance_trav);
    w.blas = plan.blas_count ? &S.stage[plan.blas_f4] : reinterpret_cast<const float4*>(sc.asset_nodes);
    w.prims = plan.prim_count ? &S.stage[plan.prim_f4] : reinterpret_cast<const float4*>(sc.primitives);
    w.instances = sc.instances;
    w.tlas_count = sc.instance_node_count;
    w.leaf_boxes_match = sc.leaf_boxes_match;
    return w;
}

// Queue a ray for the next pool_traverse().  Called by EVERY thread of the CTA in uniform control flow; `emit` says whether this
// thread has a ray.  The queue is compacted per warp: one ballot, one shared-memory atomicAdd by the first emitting lane.
__device__ __forceinline__ void pool_push(PoolShared& S, int phase, bool emit, uint32_t kind, vec3 origin, vec3 direction, float max_distance,
                                          float early_distance, uint32_t arg) {
    const uint32_t m = __ballot_sync(FULL_MASK, emit);
    if (m == 0u) return;
    const int leader = __ffs((int)m) - 1;
    const int lane = (int)(threadIdx.x & 31u);
    int base = 0;
    if (lane == leader) base = atomicAdd(&S.q_count[phase & 1], __popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (emit) {
        const int t = (int)threadIdx.x;
        S.queue[base + __popc(m & lanemask_lt())] = (uint16_t)t;
        S.ox[t] = origin.x; S.oy[t] = origin.y; S.oz[t] = origin.z;
        S.dx[t] = direction.x; S.dy[t] = direction.y; S.dz[t] = direction.z;
        S.tmax[t] = max_distance; S.early[t] = early_distance;
        S.arg[t] = arg; S.kind[t] = kind;
    }
}
__device__ __forceinline__ Hit pool_result(const PoolShared& S) {
    const int t = (int)threadIdx.x;
    Hit h;
    h.u = S.ox[t]; h.v = S.oy[t]; h.distance = S.oz[t];
    h.instance_index = __float_as_uint(S.dx[t]); h.primitive_index = __float_as_uint(S.dy[t]);
    return h;
}

// world -> object space of instance `i` from the (possibly staged) compact traversal record; light.wgsl:306-316.
// Same arithmetic as instance_ray() in hk_device.cuh.
__device__ __forceinline__ uint4 walk_instance_ray(const WalkScene& W, uint32_t i, const Ray& ray, Ray& r) {
    const float4* m = W.itrav + 5u * (size_t)i;
    const vec4 c0 = f4v(m[0]), c1 = f4v(m[1]), c2 = f4v(m[2]), c3 = f4v(m[3]);
    const float4 mesh_bits = m[4];
    const vec4 o = v4(ray.origin, 1.0f), d = v4(ray.direction, 0.0f);
    const vec4 po = v4(dot(c0, o), dot(c1, o), dot(c2, o), dot(c3, o));
    r.origin = (po.w == 1.0f) ? xyz(po) : xyz(po) / po.w;
    r.direction = v3(dot(c0, d), dot(c1, d), dot(c2, d));
    r.inv_direction = 1.0f / r.direction;
    return make_uint4(__float_as_uint(mesh_bits.x), __float_as_uint(mesh_bits.y), __float_as_uint(mesh_bits.z), __float_as_uint(mesh_bits.w));
}

// The walk of all queued rays of phase `phase` by the CTA's traversal warps.  EVERY thread of the CTA calls this (uniform control
// flow); on return every queued slot holds its Hit.  One definition, not inlined: the walk is the same ~700 instructions for every
// caller, and all warps of the CTA that execute it do so at the same time, which is what the instruction caches want.
static __device__ HK_NOINLINE void pool_traverse(PoolShared& S, const WalkScene& W, int phase) {
    __syncthreads();                                   // every ray of this phase is in its slot, q_count is final
    const int n = S.q_count[phase & 1];
    if (threadIdx.x == 0) { S.q_count[(phase + 1) & 1] = 0; S.q_next[(phase + 1) & 1] = 0; }
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    if (warp < POOL_TRAVERSE_WARPS && n > 0) {
        bool have = false, drained = false;
        int slot = 0;
        Ray ray, cur;                                  // world-space ray | the ray of the level being walked
        Hit hit;
        const float4* nodes = W.tlas;
        uint32_t count = 0, index = 0, tlas_resume = 0, instance_index = 0, mesh_primitive = 0, exclude = 0;
        float early = 0.0f;
        bool in_blas = false, blas_hit = false;
        hit.u = hit.v = hit.distance = 0.0f; hit.instance_index = hit.primitive_index = U32_MAX;
        ray.origin = ray.direction = ray.inv_direction = v3(0.0f); cur = ray;
        for (;;) {
            // ---- dynamic fetch: lanes without a ray pull the next queued ones
            const uint32_t idle = drained ? 0u : __ballot_sync(FULL_MASK, !have);
            if (idle) {
                const int leader = __ffs((int)idle) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&S.q_next[phase & 1], __popc(idle));
                base = __shfl_sync(FULL_MASK, base, leader);
                if (base + __popc(idle) > n) drained = true;             // warp-uniform: the queue cannot serve all of us any more
                if (!have) {
                    const int i = base + __popc(idle & lanemask_lt());
                    if (i < n) {
                        slot = (int)S.queue[i];
                        ray.origin = v3(S.ox[slot], S.oy[slot], S.oz[slot]);
                        ray.direction = v3(S.dx[slot], S.dy[slot], S.dz[slot]);
                        ray.inv_direction = 1.0f / ray.direction;
                        hit.u = 0.0f; hit.v = 0.0f; hit.distance = S.tmax[slot];
                        hit.instance_index = U32_MAX; hit.primitive_index = U32_MAX;
                        early = S.early[slot];
                        const uint32_t arg = S.arg[slot];
                        have = true;
                        if (S.kind[slot] == RAY_KIND_TLAS) {             // traverse_top, light.wgsl:442-486
                            exclude = arg;
                            nodes = W.tlas; count = W.tlas_count; index = 0;
                            cur = ray; in_blas = false; blas_hit = false;
                        } else {                                          // stand-alone traverse_bottom of select_light_candidate, :687
                            exclude = DONT_EXCLUDE;
                            const uint4 mesh = walk_instance_ray(W, arg, ray, cur);
                            in_blas = true; blas_hit = false;
                            tlas_resume = W.tlas_count;                   // "returns" into an exhausted TLAS: the walk ends with the BLAS
                            instance_index = arg; mesh_primitive = mesh.y;
                            nodes = W.blas + 2u * (size_t)mesh.z; count = mesh.w; index = 0;
                        }
                    }
                }
            }
            if (!__any_sync(FULL_MASK, have)) break;
            if (!have) continue;
            // ---- one step of the unified TLAS / BLAS walk (hk_device.cuh traverse_top): interior records until a leaf record or
            // the end of the level, then the leaf / level event
            bool done = false;
            uint32_t entry = 0, exit_index = 0;
            while (index < count) {
                const float4 n0 = nodes[2u * index], n1 = nodes[2u * index + 1u];
                entry = __float_as_uint(n0.w);
                exit_index = __float_as_uint(n1.w);
                if (entry >= BVH_LEAF_FLAG) break;
                index = (slab(cur, f4xyz(n0), f4xyz(n1)) < hit.distance) ? entry : exit_index;
            }
            if (index >= count) {
                if (!in_blas) done = true;
                else {
                    in_blas = false;
                    if (blas_hit) {
                        hit.instance_index = instance_index;
                        if (hit.distance < early) done = true;
                    }
                    nodes = W.tlas; count = W.tlas_count; index = tlas_resume;
                    cur = ray;
                }
            } else {
                const bool via_navigator = index != 0u && W.leaf_boxes_match != 0u;
                index = exit_index;
                if (!in_blas) {
                    const uint32_t candidate = entry - BVH_LEAF_FLAG;
                    if (candidate != exclude) {
                        bool pass = via_navigator;
                        if (!pass) {
                            const hk_instance* inst = W.instances + candidate;
                            const float4 imin = ldg4(inst->min), imax = ldg4(inst->max);
                            pass = slab(ray, f4xyz(imin), f4xyz(imax)) < hit.distance;
                        }
                        if (pass) {
                            const uint4 mesh = walk_instance_ray(W, candidate, ray, cur);
                            in_blas = true; blas_hit = false;
                            tlas_resume = exit_index; instance_index = candidate; mesh_primitive = mesh.y;
                            nodes = W.blas + 2u * (size_t)mesh.z; count = mesh.w; index = 0;
                        }
                    }
                } else {
                    const uint32_t primitive_index = mesh_primitive + entry - BVH_LEAF_FLAG;
                    const float4* prim = W.prims + 3u * (size_t)primitive_index;
                    const float4 a = prim[0], b = prim[1], c = prim[2];
                    const vec3 p0 = f4xyz(a), p1 = f4xyz(b), p2 = f4xyz(c);
                    if (via_navigator || slab(cur, vmin(p0, vmin(p1, p2)), vmax(p0, vmax(p1, p2))) < hit.distance) {
                        float u, v;
                        const float distance = triangle(cur, p0, p1, p2, u, v);
                        if (distance < hit.distance) {
                            hit.u = u; hit.v = v; hit.distance = distance;
                            hit.primitive_index = primitive_index;
                            blas_hit = true;
                            if (distance < early) { hit.instance_index = instance_index; done = true; }
                        }
                    }
                }
            }
            if (done) {
                S.ox[slot] = hit.u; S.oy[slot] = hit.v; S.oz[slot] = hit.distance;
                S.dx[slot] = __uint_as_float(hit.instance_index); S.dy[slot] = __uint_as_float(hit.primitive_index);
                have = false;
            }
        }
    }
    __syncthreads();                                   // every hit is in its slot
}

// 8x4-pixel sub-tiles per warp inside the CTA's 16x16 tile (ray coherence + whole-sector plane accesses, as tile_pixel())
__device__ __forceinline__ void pool_pixel(int& x, int& y, const KParams& P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = P.col_lo + blockIdx.x * POOL_TILE_W + (warp & 1) * 8 + (lane & 7);
    y = P.row_lo + blockIdx.y * POOL_TILE_H + (warp >> 1) * 4 + (lane >> 3);
}

}  // namespace hkd
