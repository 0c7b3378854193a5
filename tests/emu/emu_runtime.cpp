// emu_runtime.cpp — TEST INFRASTRUCTURE: globals of cuda_emu.h and the fiber scheduler of cooperative launches.
#include <stdlib.h>

#include "cuda_emu.h"

thread_local EmuIdx threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
thread_local EmuCta* emu_cta = nullptr;

int emu_reverse_order() { static int v = getenv("HK_EMU_REVERSE") ? 1 : 0; return v; }

[[noreturn]] void emu_deadlock(const char* what) {
    fprintf(stderr, "cuda_emu: %s (block %u,%u thread %u)\n", what, blockIdx.x, blockIdx.y, threadIdx.x);
    abort();
}

static void emu_trampoline() {
    EmuCta* c = emu_cta;
    c->body();
    c->on_exit_thread();
    c->fibers[c->current].done = true;
    swapcontext(&c->fibers[c->current].ctx, &c->sched);
}

void EmuCta::on_exit_thread() {
    // an exited thread no longer takes part in collectives: complete the ones that were only waiting for it
    alive -= 1;
    EmuWarpState& w = warps[current >> 5];
    w.alive_mask &= ~(1u << (current & 31u));
    if (w.arrived) {
        const uint32_t need = w.pending_mask & w.alive_mask;
        if ((w.arrived & need) == need) { w.arrived = 0; w.gen += 1u; }
    }
    if (alive > 0 && bar_arrived >= alive) { bar_arrived = 0; bar_acc[(bar_gen + 1u) & 1u] = 0; bar_cnt[(bar_gen + 1u) & 1u] = 0; bar_gen += 1u; }
}

void EmuCta::yield_until_changed(const volatile uint32_t* p, uint32_t v) {
    EmuFiber& f = fibers[current];
    f.wait_ptr = p; f.wait_val = v;
    swapcontext(&f.ctx, &sched);
    f.wait_ptr = nullptr;
}

void EmuCta::run(const dim3& block, const std::function<void()>& f, bool reverse) {
    n = block.x * block.y * block.z;
    bdim = block;
    body = f;
    if (fibers.size() < n) fibers.resize(n);
    warps.assign((n + 31) / 32, EmuWarpState());
    for (unsigned t = 0; t < n; ++t) {
        EmuFiber& fb = fibers[t];
        if (!fb.stack) {
            fb.stack = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (fb.stack == MAP_FAILED) emu_deadlock("cannot allocate a fiber stack");
        }
        getcontext(&fb.ctx);
        fb.ctx.uc_stack.ss_sp = fb.stack;
        fb.ctx.uc_stack.ss_size = STACK;
        fb.ctx.uc_link = nullptr;
        makecontext(&fb.ctx, emu_trampoline, 0);
        fb.done = false; fb.wait_ptr = nullptr;
        warps[t >> 5].alive_mask |= 1u << (t & 31u);
    }
    alive = n;
    bar_gen = 0; bar_arrived = 0; bar_acc[0] = bar_acc[1] = 0; bar_cnt[0] = bar_cnt[1] = 0;
    EmuCta* outer = emu_cta;
    emu_cta = this;
    while (alive > 0) {
        bool progressed = false;
        for (unsigned u = 0; u < n; ++u) {
            const unsigned t = reverse ? n - 1 - u : u;
            EmuFiber& fb = fibers[t];
            if (fb.done) continue;
            if (fb.wait_ptr && *fb.wait_ptr == fb.wait_val) continue;   // still blocked
            current = t;
            threadIdx.x = t % bdim.x; threadIdx.y = (t / bdim.x) % bdim.y; threadIdx.z = t / (bdim.x * bdim.y);
            swapcontext(&sched, &fb.ctx);
            progressed = true;
        }
        if (!progressed) emu_deadlock("cooperative launch deadlocked: a barrier or warp collective can never complete");
    }
    emu_cta = outer;
}
