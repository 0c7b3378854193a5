// cuda_emu.h — TEST INFRASTRUCTURE.  A host shim for the handful of CUDA constructs the kernels use, so that the kernel
// SOURCES (bevy_hikari_b200/csrc/*.cu, unmodified except for the launch syntax, see build_emu.py) can be compiled with g++
// and their LOGIC compared with the oracle without a GPU (tests/test_emulated_kernels.py).  It is not a fallback: the
// package never loads the library built from it, performance is irrelevant, and nothing here is shipped.
// Execution model, two kinds of launch:
//   EMU_LAUNCH       every thread of every block runs to completion, one after the other (blocks in parallel with OpenMP).
//                    Equivalent for kernels without shared memory, barriers or warp collectives (k_*).
//   EMU_LAUNCH_COOP  for the cooperative kernels (kc_*: shared-memory ray pool, __syncthreads, ballots, shuffles, mbarrier):
//                    the threads of a block are fibers (ucontext) on one host thread; a fiber runs until it reaches a collective
//                    that is not complete yet, then the scheduler switches to the next runnable one.  __shared__ variables are
//                    static thread_local, i.e. one copy per host thread = per block in flight.  A collective that can never
//                    complete (divergent barrier, a lane missing from a ballot) is reported as a deadlock instead of hanging.
#pragma once
#define HK_EMU 1
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__

struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(4) uchar4 { uint8_t x, y, z, w; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

// CUDA's global min / max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }

struct EmuIdx { unsigned x, y, z; };
extern thread_local EmuIdx threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// ------------------------------------------------------------------------------------------------ runtime
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef struct EmuStream* cudaStream_t;
typedef struct EmuEvent* cudaEvent_t;
typedef unsigned long long cudaTextureObject_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; };
struct cudaIpcMemHandle_t { char reserved[64]; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t bytes) { return posix_memalign(p, 256, bytes ? bytes : 256) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.01f; return cudaSuccess; }   // a constant: nothing is timed here
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeDevice; a->device = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h, &p, sizeof(p)); return cudaSuccess; }
// one process only: "opening" a handle gives the exporter's pointer back (real CUDA refuses this inside the exporting process;
// the two-process paths are exercised on the device, tests/test_gpu_frame_assembly.py, tests/test_gpu_zz_halo.py)
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, &h, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }


// ------------------------------------------------------------------------------------- cooperative kernels (fibers)
#include <ucontext.h>
#include <stdio.h>
#include <sys/mman.h>

#include <functional>
#include <vector>

#define __shared__ static thread_local
#define HK_NOINLINE __attribute__((noinline))

struct EmuWarpState {
    uint32_t gen = 0, arrived = 0, pending_mask = 0, alive_mask = 0;
    uint32_t vals[2][32];
};
struct EmuFiber {
    ucontext_t ctx;
    void* stack = nullptr;
    bool done = true;
    const volatile uint32_t* wait_ptr = nullptr;   // runnable again when *wait_ptr != wait_val
    uint32_t wait_val = 0;
};
struct EmuCta {
    static const size_t STACK = 512 * 1024;
    std::vector<EmuFiber> fibers;
    std::vector<EmuWarpState> warps;
    ucontext_t sched;
    unsigned n = 0, current = 0, alive = 0;
    dim3 bdim;
    uint32_t bar_gen = 0, bar_arrived = 0, bar_acc[2] = {0, 0}, bar_cnt[2] = {0, 0};
    std::function<void()> body;
    void run(const dim3& block, const std::function<void()>& f, bool reverse);
    void yield_until_changed(const volatile uint32_t* p, uint32_t v);
    void on_exit_thread();
};
extern int emu_reverse_order();
extern thread_local EmuCta* emu_cta;      // the block this host thread is running cooperatively, or nullptr
[[noreturn]] void emu_deadlock(const char* what);

static inline unsigned emu_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }

// arrive with a 32-bit contribution; returns the generation whose vals[gen & 1][lane] hold every participant's contribution
static inline uint32_t emu_warp_collect(uint32_t mask, uint32_t value) {
    EmuCta* c = emu_cta;
    if (!c) emu_deadlock("warp collective in a kernel launched with EMU_LAUNCH (name it kc_* for a cooperative launch)");
    const unsigned tid = emu_tid(), lane = tid & 31u;
    EmuWarpState& w = c->warps[tid >> 5];
    if (!((mask >> lane) & 1u)) emu_deadlock("lane executes a *_sync collective it is not named in");
    const uint32_t g = w.gen;
    w.vals[g & 1u][lane] = value;
    w.arrived |= 1u << lane;
    w.pending_mask = mask;
    const uint32_t need = mask & w.alive_mask;
    if ((w.arrived & need) == need) { w.arrived = 0; w.gen = g + 1u; }
    else c->yield_until_changed(&w.gen, g);
    return g;
}
// every emulated thread is "converged" only with itself: code that votes over __activemask() (traverse_top) degenerates to a
// one-lane warp, in sequential and in cooperative launches alike
static inline uint32_t __activemask() { return 1u << (emu_tid() & 31u); }
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    if (mask == (1u << (emu_tid() & 31u))) return pred ? mask : 0u;
    const uint32_t g = emu_warp_collect(mask, pred ? 1u : 0u);
    const EmuWarpState& w = emu_cta->warps[emu_tid() >> 5];
    uint32_t r = 0;
    for (unsigned l = 0; l < 32; ++l) if (((mask & w.alive_mask) >> l) & 1u) r |= (w.vals[g & 1u][l] & 1u) << l;
    return r;
}
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, !pred) == 0u; }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu_warp_collect(mask, 0u); }
template <class T> static inline T emu_shfl_from(uint32_t mask, T v, unsigned src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t bits; memcpy(&bits, &v, 4);
    const uint32_t g = emu_warp_collect(mask, bits);
    const uint32_t out = emu_cta->warps[emu_tid() >> 5].vals[g & 1u][src_lane & 31u];
    T r; memcpy(&r, &out, 4);
    return r;
}
template <class T> static inline T __shfl_sync(uint32_t mask, T v, int src) { return emu_shfl_from(mask, v, (unsigned)src); }
// sequential launches: flush_counters is patched to per-thread adds (build_emu.py) and never reaches this; cooperative launches shuffle for real
template <class T> static inline T __shfl_xor_sync(uint32_t mask, T v, int lane_mask) { return emu_cta ? emu_shfl_from(mask, v, (emu_tid() & 31u) ^ (unsigned)lane_mask) : T(0); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline uint32_t emu_lanemask_lt() { return (1u << (emu_tid() & 31u)) - 1u; }

static inline void emu_block_barrier(int pred, uint32_t* or_out, uint32_t* count_out) {
    EmuCta* c = emu_cta;
    if (!c) emu_deadlock("__syncthreads in a kernel launched with EMU_LAUNCH (name it kc_* for a cooperative launch)");
    const uint32_t g = c->bar_gen;
    if (pred) { c->bar_acc[g & 1u] = 1u; c->bar_cnt[g & 1u] += 1u; }
    c->bar_arrived += 1;
    if (c->bar_arrived >= c->alive) { c->bar_arrived = 0; c->bar_acc[(g + 1u) & 1u] = 0; c->bar_cnt[(g + 1u) & 1u] = 0; c->bar_gen = g + 1u; }
    else c->yield_until_changed(&c->bar_gen, g);
    if (or_out) *or_out = c->bar_acc[g & 1u];
    if (count_out) *count_out = c->bar_cnt[g & 1u];
}
static inline void __syncthreads() { emu_block_barrier(0, nullptr, nullptr); }
static inline int __syncthreads_or(int pred) { uint32_t r; emu_block_barrier(pred, &r, nullptr); return (int)r; }
static inline int __syncthreads_count(int pred) { uint32_t r; emu_block_barrier(pred, nullptr, &r); return (int)r; }

// shared-memory atomics of the cooperative kernels (one host thread per block: plain read-modify-write)
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }

// wait until a 32-bit word written by another thread of the block changes (the emulated mbarrier of hk_pool.cuh)
static inline void emu_wait_changed(const volatile uint32_t* p, uint32_t v) {
    if (*p != v) return;
    if (!emu_cta) emu_deadlock("wait on another thread's write in a kernel launched with EMU_LAUNCH");
    emu_cta->yield_until_changed(p, v);
}

#define EMU_LAUNCH_COOP(GRID, BLOCK, ...)                                                 \
    do {                                                                                  \
        const dim3 emu_g = (GRID), emu_b = (BLOCK);                                       \
        const long long emu_n = (long long)emu_g.x * emu_g.y * emu_g.z;                   \
        const bool emu_rev = emu_reverse_order() != 0;                                    \
        _Pragma("omp parallel for schedule(dynamic, 1) if (!emu_rev)")                    \
        for (long long emu_j = 0; emu_j < emu_n; ++emu_j) {                               \
            const long long emu_i = emu_rev ? emu_n - 1 - emu_j : emu_j;                  \
            gridDim = emu_g; blockDim = emu_b;                                            \
            blockIdx.x = (unsigned)(emu_i % emu_g.x);                                     \
            blockIdx.y = (unsigned)((emu_i / emu_g.x) % emu_g.y);                         \
            blockIdx.z = (unsigned)(emu_i / ((long long)emu_g.x * emu_g.y));              \
            static thread_local EmuCta emu_block;                                         \
            emu_block.run(emu_b, [&]() { __VA_ARGS__; }, emu_rev);                        \
        }                                                                                 \
    } while (0)

// ------------------------------------------------------------------------------------------------ launches
// EMU_LAUNCH(grid, block, kernel_call): every thread of every block, blocks distributed over the host cores
// HK_EMU_REVERSE=1 runs blocks and threads in descending order on one host thread: if a launch's result depended on the order
// in which its threads run (a read of something another thread of the same launch writes), the parity tests would fail in one
// of the two orders.  They pass in both.
extern int emu_reverse_order();
#define EMU_LAUNCH(GRID, BLOCK, ...)                                                      \
    do {                                                                                  \
        const dim3 emu_g = (GRID), emu_b = (BLOCK);                                       \
        const long long emu_n = (long long)emu_g.x * emu_g.y * emu_g.z;                   \
        const unsigned emu_tn = emu_b.x * emu_b.y * emu_b.z;                              \
        const bool emu_rev = emu_reverse_order() != 0;                                    \
        _Pragma("omp parallel for schedule(dynamic, 4) if (!emu_rev)")                    \
        for (long long emu_j = 0; emu_j < emu_n; ++emu_j) {                               \
            const long long emu_i = emu_rev ? emu_n - 1 - emu_j : emu_j;                  \
            gridDim = emu_g; blockDim = emu_b;                                            \
            blockIdx.x = (unsigned)(emu_i % emu_g.x);                                     \
            blockIdx.y = (unsigned)((emu_i / emu_g.x) % emu_g.y);                         \
            blockIdx.z = (unsigned)(emu_i / ((long long)emu_g.x * emu_g.y));              \
            for (unsigned emu_u = 0; emu_u < emu_tn; ++emu_u) {                           \
                const unsigned emu_t = emu_rev ? emu_tn - 1 - emu_u : emu_u;              \
                threadIdx.x = emu_t % emu_b.x;                                            \
                threadIdx.y = (emu_t / emu_b.x) % emu_b.y;                                \
                threadIdx.z = emu_t / (emu_b.x * emu_b.y);                                \
                __VA_ARGS__;                                                              \
            }                                                                             \
        }                                                                                 \
    } while (0)
