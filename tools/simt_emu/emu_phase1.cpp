// emu_phase1.cpp — TEST-ONLY: compiles cpuvox_b200/csrc/raybuffer_kernels.cu for the CPU through the SIMT emulator
// (cuda_emu.h) and exposes Phase 1 to the tests, so kernel logic can be compared with the oracle without a GPU.
// Not linked into libcpuvox_b200.so and never used by the product path.
#define CVX_EMU 1
#include "cuda_emu.h"

#include <atomic>
#include <thread>
#include <vector>

#ifdef CVX_EMU_STATS
unsigned long long* emu_stats = nullptr;
#endif
#include "../../cpuvox_b200/csrc/raybuffer_kernels.cu"
#include "../../cpuvox_b200/csrc/host_frame.h"
#include "../../cpuvox_b200/csrc/world_transcode.h"

struct emu_world {
    cvxh_lod_tables tables[CVXD_LODS];
    std::vector<uint32_t> elements[CVXD_LODS];
    cvxd_world w;
};

namespace {

struct LaunchArgs { const cvxd_world* world; const cvxd_frame* frame; int variant, group; bool counters; };

template <int G, bool INV>
void body_gi(const LaunchArgs* a) {
    if (a->variant == 0) {
        if (a->counters) phase1_kernel<G, true, false, false, INV>(*a->world, *a->frame);
        else phase1_kernel<G, false, false, false, INV>(*a->world, *a->frame);
    } else {
        if (a->counters) phase1_kernel<G, true, false, true, INV>(*a->world, *a->frame);
        else phase1_kernel<G, false, false, true, INV>(*a->world, *a->frame);
    }
}

template <int G>
void body_g(void* p) {
    const LaunchArgs* a = (const LaunchArgs*)p;
    if (a->frame->inverse) body_gi<G, true>(a); else body_gi<G, false>(a);
}

} // namespace

extern "C" {

emu_world* emu_world_create(int lods, const int32_t dims[3], const void* const* blobs, const int64_t* bytes, const int32_t* column_counts) {
    emu_world* w = new emu_world();
    memset(&w->w, 0, sizeof w->w);
    cvxd_world_set_dims(&w->w, dims[0], dims[1], dims[2]);
    w->w.lod_count = lods;
    w->w.regular = 1;
    for (int l = 0; l < lods; l++) {
        const int64_t needCols = (int64_t)(dims[0] >> l) * (dims[2] >> l);
        const int64_t headerBytes = 12 * (int64_t)column_counts[l];
        const int64_t cells = (bytes[l] - headerBytes) / 4;
        if (!cvxh_transcode_lod(blobs[l], needCols, column_counts[l], cells, l, dims[1], w->tables[l])) { delete w; return nullptr; }
        w->elements[l].assign((const uint32_t*)((const uint8_t*)blobs[l] + headerBytes), (const uint32_t*)((const uint8_t*)blobs[l] + headerBytes) + cells);
        cvxd_lod& d = w->w.lods[l];
        d.headers = (const uint4*)w->tables[l].headers.data();
        d.elements = w->elements[l].data();
        d.bounds = (const uint2*)w->tables[l].bounds.data();
        d.mul_x = dims[2] >> l;
        d.lod = l;
        if (!w->tables[l].regular) w->w.regular = 0;
    }
    return w;
}

#ifdef CVX_EMU_STATS
void emu_set_stats(unsigned long long* p) { emu_stats = p; } // 16 counters per flat ray
#endif
void emu_world_destroy(emu_world* w) { delete w; }
int emu_world_regular(const emu_world* w) { return w->w.regular; }

// Renders rays [ray_begin, ray_end) of the frame into td / lr (host memory, same layout as the device raybuffers).
int emu_phase1(const emu_world* w, const cvx_frame_setup* setup, int W, int H, uint32_t* td, uint32_t* lr, cvxd_counters* counters,
               int variant, int group, int threads, int ray_begin, int ray_end) {
    cvxd_frame f;
    cvxh::frame_from_setup(setup, W, H, f);
    cvxd_frame_set_world(&f, &w->w);
    f.td = td; f.lr = lr; f.counters = counters;
    if (ray_end < 0 || ray_end > f.total_rays) ray_end = f.total_rays;
    if (ray_begin < 0) ray_begin = 0;
    f.ray_begin = ray_begin; f.ray_end = ray_end;
    const int n = ray_end - ray_begin;
    if (n <= 0) return 0;
    if (group != 8 && group != 16 && group != 32) group = 32;
    const int groupsPerCta = CVXD_THREADS_PER_CTA / group;
    const int blocks = (n + groupsPerCta - 1) / groupsPerCta;
    const int seenWords = ((W > H ? W : H) + 31) >> 5;
    const size_t smemWords = (size_t)groupsPerCta * (seenWords + 9 * group) + 64;
    LaunchArgs args{&w->w, &f, variant, group, counters != nullptr};
    std::atomic<int> next{0};
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads > blocks) threads = blocks;
    auto worker = [&]() {
        std::vector<uint32_t> smem(smemWords);
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= blocks) break;
            emu::g_shared = smem.data();
            for (int warp = 0; warp < CVXD_THREADS_PER_CTA / 32; warp++) {
                emu_dim3 tids[32];
                for (int i = 0; i < 32; i++) tids[i] = emu_dim3{warp * 32 + i, 0, 0};
                void (*body)(void*) = group == 32 ? body_g<32> : (group == 16 ? body_g<16> : body_g<8>);
                emu::run_warp(body, &args, tids, emu_dim3{b, 0, 0}, emu_dim3{CVXD_THREADS_PER_CTA, 1, 1}, emu_dim3{blocks, 1, 1}, 32);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    return n;
}

} // extern "C"
