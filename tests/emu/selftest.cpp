// selftest.cpp — TEST INFRASTRUCTURE: the cooperative part of cuda_emu.h (fibers, __syncthreads, ballots, shuffles, shared memory,
// the wait primitive) checked against closed-form answers, in ascending and descending thread order.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "cuda_emu.h"

// per block: compaction of the even-valued inputs through ballot + popc + a shared counter, a block sum through shuffles and shared
// memory, a producer/consumer hand-off through a waited word
static void kc_selftest(const int* in, int n, int* compacted, int* counts, int* sums, int* handoff) {
    __shared__ int s_count;
    __shared__ int s_partial[8];
    __shared__ uint32_t s_flag;
    __shared__ int s_value;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = (int)blockIdx.x * (int)blockDim.x + tid;
    if (tid == 0) { s_count = 0; s_flag = 0u; }
    __syncthreads();
    const int v = i < n ? in[i] : 1;
    const bool keep = i < n && (v % 2) == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&s_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) compacted[(size_t)blockIdx.x * blockDim.x + base + __popc(m & ((1u << lane) - 1u))] = v;
    int s = i < n ? v : 0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_partial[warp] = s;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (unsigned w = 0; w < blockDim.x / 32; ++w) t += s_partial[w];
        sums[blockIdx.x] = t;
        counts[blockIdx.x] = s_count;
    }
    // hand-off: the LAST thread produces, everybody else waits for the flag (a thread that runs first must really block)
    if (tid == (int)blockDim.x - 1) { s_value = 1000 + (int)blockIdx.x; s_flag = 1u; }
    else emu_wait_changed(&s_flag, 0u);
    if (tid == 0) handoff[blockIdx.x] = s_value;
    const int any_big = __syncthreads_or(v > 90);
    if (tid == 1) handoff[blockIdx.x] += any_big ? 100000 : 0;
}

int main() {
    const int n = 1000, B = 256, G = (n + B - 1) / B;
    std::vector<int> in(n), compacted((size_t)G * B, -1), counts(G), sums(G), handoff(G);
    for (int i = 0; i < n; ++i) in[i] = (i * 37 + 11) % 97;
    EMU_LAUNCH_COOP(dim3(G), dim3(B), kc_selftest(in.data(), n, compacted.data(), counts.data(), sums.data(), handoff.data()));
    int bad = 0;
    for (int b = 0; b < G; ++b) {
        int want_sum = 0, want_count = 0, big = 0;
        std::vector<int> evens;
        for (int i = b * B; i < (b + 1) * B && i < n; ++i) { want_sum += in[i]; if (in[i] % 2 == 0) { ++want_count; evens.push_back(in[i]); } if (in[i] > 90) big = 1; }
        if (sums[b] != want_sum || counts[b] != want_count || handoff[b] != 1000 + b + 100000 * big) ++bad;
        // compaction is order preserving inside a warp and warps land in the order their leaders reach the counter: compare as multisets
        std::vector<int> got(compacted.begin() + (size_t)b * B, compacted.begin() + (size_t)b * B + want_count);
        std::sort(got.begin(), got.end()); std::sort(evens.begin(), evens.end());
        if (got != evens) ++bad;
    }
    printf("selftest %s\n", bad ? "FAILED" : "ok");
    return bad ? 1 : 0;
}
