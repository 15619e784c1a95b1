/*
 * nvtx_ranges.h — NVTX ranges with the reference's own profiler sample names (Profiler.BeginSample in
 * Assets/Code/RenderManager.cs:97,119,127,154,173,178 and :277), so a timeline of this library reads like the Unity profiler's.
 * NVTX v3 is header-only: without a tool attached a range costs one predictable branch. Build with -DCVX_NVTX=0 to compile them out.
 */
#pragma once
#ifndef CVX_NVTX
#define CVX_NVTX 1
#endif
#if CVX_NVTX
#include <nvtx3/nvToolsExt.h>
struct cvx_nvtx_range {
    explicit cvx_nvtx_range(const char* name) { nvtxRangePushA(name); }
    ~cvx_nvtx_range() { nvtxRangePop(); }
    cvx_nvtx_range(const cvx_nvtx_range&) = delete;
};
#define CVX_RANGE_CAT2(a, b) a##b
#define CVX_RANGE_CAT(a, b) CVX_RANGE_CAT2(a, b)
#define CVX_RANGE(name) cvx_nvtx_range CVX_RANGE_CAT(cvx_range_, __LINE__)(name)
#else
#define CVX_RANGE(name) do { } while (0)
#endif
