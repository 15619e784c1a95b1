// cuda_emu.cpp — fiber scheduler of the test-only SIMT emulator (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {

thread_local Warp* g_warp = nullptr;
thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local uint32_t* g_shared = nullptr;

static const size_t STACK_BYTES = 256 * 1024;

static void lane_entry() {
    Warp* w = g_warp;
    const int lane = w->cur;
    w->body(w->arg);
    w->done[lane] = true;
    w->progress++;
    swapcontext(&w->ctx[lane], &w->sched); // never resumed
}

void run_warp(void (*body)(void*), void* arg, const emu_dim3 tids[32], emu_dim3 bid, emu_dim3 bdim, emu_dim3 gdim, int lanes) {
    static thread_local Warp* w = nullptr;
    if (!w) {
        w = new Warp();
        w->stacks = (char*)malloc(STACK_BYTES * 32);
    }
    g_warp = w;
    w->body = body; w->arg = arg; w->ncolls = 0; w->progress = 0;
    blockIdx = bid; blockDim = bdim; gridDim = gdim;
    for (int i = 0; i < 32; i++) {
        w->done[i] = i >= lanes;
        w->tid[i] = tids[i];
        if (i >= lanes) continue;
        getcontext(&w->ctx[i]);
        w->ctx[i].uc_stack.ss_sp = w->stacks + STACK_BYTES * i;
        w->ctx[i].uc_stack.ss_size = STACK_BYTES;
        w->ctx[i].uc_link = nullptr;
        makecontext(&w->ctx[i], (void (*)())lane_entry, 0);
    }
    for (;;) {
        bool alive = false;
        const unsigned long before = w->progress;
        for (int i = 0; i < 32; i++) {
            if (w->done[i]) continue;
            alive = true;
            w->cur = i;
            threadIdx = w->tid[i];
            swapcontext(&w->sched, &w->ctx[i]);
        }
        if (!alive) break;
        if (w->progress == before) {
            fprintf(stderr, "emu: warp deadlock (block %d): lanes wait on collectives that can never complete\n", bid.x);
            for (int c = 0; c < w->ncolls; c++) fprintf(stderr, "  mask %08x arrived %08x gen %u\n", w->colls[c].mask, w->colls[c].arrived, w->colls[c].gen);
            abort();
        }
    }
}

} // namespace emu
