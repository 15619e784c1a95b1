/*
 * cuda_emu.h — a tiny SIMT emulator for TEST use only: runs a warp-synchronous CUDA kernel source on the CPU, one fiber
 * (ucontext) per lane, so that kernel logic can be checked against the oracle in a container without a GPU.
 * Never part of the product: libcpuvox_b200.so does not contain or call any of this (the product has no CPU path).
 *
 * Model: one OS thread runs one warp at a time; its 32 lanes are fibers scheduled round-robin. A warp collective
 * (__shfl*_sync, __ballot_sync, __syncwarp, __reduce_or_sync) deposits the lane's value and yields until every lane named in
 * the mask has arrived, exactly the contract of the *_sync intrinsics. Shared memory is a per-CTA buffer; the kernels here
 * never use __syncthreads, so the warps of a CTA run one after another.
 */
#pragma once
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct emu_dim3 { int x, y, z; };
typedef int cudaError_t;
typedef void* cudaStream_t;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__

namespace emu {

struct Coll { unsigned mask, arrived, gen; unsigned vals[32]; unsigned res[2][32]; };

struct Warp {
    ucontext_t sched;
    ucontext_t ctx[32];
    char* stacks = nullptr;
    bool done[32];
    int cur = 0;
    unsigned long progress = 0;
    Coll colls[8];
    int ncolls = 0;
    void (*body)(void*) = nullptr;
    void* arg = nullptr;
    emu_dim3 tid[32];
};

extern thread_local Warp* g_warp;
extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local uint32_t* g_shared;

static inline void yield_lane() {
    Warp* w = g_warp;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

static inline const unsigned* exchange(unsigned mask, unsigned v) {
    Warp* w = g_warp;
    Coll* c = nullptr;
    for (int i = 0; i < w->ncolls; i++) if (w->colls[i].mask == mask) { c = &w->colls[i]; break; }
    if (!c) {
        if (w->ncolls >= 8) { fprintf(stderr, "emu: too many distinct collective masks\n"); abort(); }
        c = &w->colls[w->ncolls++];
        memset(c, 0, sizeof *c);
        c->mask = mask;
    }
    const int lane = w->cur;
    if (!((mask >> lane) & 1u)) { fprintf(stderr, "emu: lane %d calls a collective whose mask %08x excludes it\n", lane, mask); abort(); }
    c->vals[lane] = v;
    c->arrived |= 1u << lane;
    const unsigned g = c->gen;
    if (c->arrived == mask) {
        memcpy(c->res[g & 1u], c->vals, sizeof c->vals);
        c->arrived = 0;
        c->gen = g + 1;
        w->progress++;
    } else {
        while (c->gen == g) yield_lane();
    }
    return c->res[g & 1u];
}

template <class T> static inline unsigned bits_of(T v) { static_assert(sizeof(T) == 4, "32-bit values only"); unsigned u; memcpy(&u, &v, 4); return u; }
template <class T> static inline T from_bits(unsigned u) { T v; memcpy(&v, &u, 4); return v; }

void run_warp(void (*body)(void*), void* arg, const emu_dim3 tids[32], emu_dim3 bid, emu_dim3 bdim, emu_dim3 gdim, int lanes);

} // namespace emu

using emu::threadIdx;
using emu::blockIdx;
using emu::blockDim;
using emu::gridDim;

// ---- warp collectives ------------------------------------------------------------------------------------------------
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const unsigned* all = emu::exchange(mask, emu::bits_of(v));
    const int lane = emu::g_warp->cur;
    const int s = (lane & ~(width - 1)) + (src & (width - 1));
    return emu::from_bits<T>(all[s]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const unsigned* all = emu::exchange(mask, emu::bits_of(v));
    const int lane = emu::g_warp->cur;
    const int s = lane - (int)delta;
    return s < (lane & ~(width - 1)) ? v : emu::from_bits<T>(all[s]);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const unsigned* all = emu::exchange(mask, emu::bits_of(v));
    const int lane = emu::g_warp->cur;
    const int s = lane + (int)delta;
    return s > (lane | (width - 1)) ? v : emu::from_bits<T>(all[s]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const unsigned* all = emu::exchange(mask, pred ? 1u : 0u);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (((mask >> i) & 1u) && all[i]) r |= 1u << i;
    return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    const unsigned* all = emu::exchange(mask, v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1u) r |= all[i];
    return r;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)emu::exchange(mask, 0u); }

// ---- scalar intrinsics -----------------------------------------------------------------------------------------------
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __ffs(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __fns(unsigned mask, unsigned base, int offset) { // position of the offset-th set bit at or above base (offset > 0), 0xffffffff if none
    if (offset <= 0) return 0xffffffffu; // (negative / zero offsets are not used by the kernels)
    for (unsigned i = base; i < 32; i++) if ((mask >> i) & 1u) { if (--offset == 0) return i; }
    return 0xffffffffu;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) { // prmt.b32, default mode: nibble i of sel picks a byte of {b,a}
    const unsigned long long pool = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((pool >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
static inline int __float2int_rz(float f) { // cvt.rzi.s32.f32: saturating, NaN -> 0
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int)f;
}
static inline float __int_as_float(int i) { return emu::from_bits<float>((unsigned)i); }
static inline int __float_as_int(float f) { return (int)emu::bits_of(f); }
static inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(u); }
static inline unsigned __float_as_uint(float f) { return emu::bits_of(f); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline long long clock64() { return 0; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
