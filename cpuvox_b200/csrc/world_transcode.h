/*
 * world_transcode.h — host-side transcoding of one World LOD blob (the reference's allocator blob: ColumnCount 12-byte
 * RLEColumn headers, then the element area; Assets/Code/World.cs:161-209,285-313) into the device layout.
 * Header-only. The product builds the same tables on the device (transcode_*_kernel in world_builder_gpu.cu, used by cvx_world_upload
 * and cvx_world_build_from_mesh); this host version feeds the test-only SIMT emulator (tools/simt_emu) and documents the layout.
 *
 *   headers  one uint4 per addressed column (index (x >> lod) * (dimZ >> lod) + (z >> lod), World.cs:145-149):
 *              x = element offset (4-byte cells into the element area; colours at x + runCount + 2)
 *              y = runCount | worldMin << 16
 *              z = worldMax
 *              w = offset of the column's first record in `bounds`
 *   bounds   runCount + 1 records per non-empty column, top to bottom. Record i is the boundary above run i:
 *              x = world-Y of the boundary (y_0 = dimY, y_i = dimY - voxelScale * (Length_0 + .. + Length_{i-1}))
 *              y = RLEElement i verbatim (ColorsIndex | Length << 16); 0 for the last record (the column's floor)
 *            so a lane that owns a boundary has, with its neighbour lane, the world-Y extent of a run without the running
 *            sums of DrawSegmentRayJob.cs:449-455, and adjacent runs share the projection of their common boundary.
 *   regular  every non-empty column consists of runCount valid elements (Length > 0) whose lengths add up to the full
 *            column height, which is what WorldBuilder.ToFinalColumn (WordBuilder.cs:232-256) emits. Only then do the
 *            top-down and bottom-up running sums of :449-455 meet at the same integers, and only then is `bounds` used
 *            (the kernels fall back to the element area otherwise).
 */
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

struct cvxh_u4 { uint32_t x, y, z, w; };
struct cvxh_u2 { uint32_t x, y; };

struct cvxh_lod_tables {
    std::vector<cvxh_u4> headers;
    std::vector<cvxh_u2> bounds;
    bool regular = true;
    int64_t bad_column = -1; // first column whose offset/run count points outside the element area
};

// blob: headers (12 bytes each, `column_count` of them) followed by `element_cells` 4-byte cells.
inline bool cvxh_transcode_lod(const void* blob, int64_t need_cols, int64_t column_count, int64_t element_cells, int lod, int dim_y,
                               cvxh_lod_tables& out) {
    const uint8_t* p = (const uint8_t*)blob;
    const uint32_t* elements = (const uint32_t*)(p + 12 * column_count);
    out.headers.assign((size_t)need_cols, cvxh_u4{0, 0, 0, 0});
    out.bounds.clear();
    out.regular = dim_y <= 65535;
    const int scale = 1 << lod;
    for (int64_t i = 0; i < need_cols; i++) {
        uint32_t w0, w1, w2;
        memcpy(&w0, p + 12 * i, 4); memcpy(&w1, p + 12 * i + 4, 4); memcpy(&w2, p + 12 * i + 8, 4);
        const uint32_t rc = w1 & 0xffffu;
        cvxh_u4 h{w0, w1, w2 & 0xffffu, 0};
        if (rc) {
            const int32_t off = (int32_t)w0;
            if (off < 0 || (int64_t)off + rc + 2 > element_cells) { out.bad_column = i; return false; }
            // colour ranges of the solid runs (gathered by Phase 1) must stay inside the element area too
            const int64_t colour_cells = element_cells - ((int64_t)off + rc + 2);
            for (uint32_t k = 0; k < rc; k++) {
                const uint32_t e = elements[(int64_t)off + 1 + k];
                const int ci = (int)(int16_t)(e & 0xffffu), len = (int)(int16_t)(e >> 16);
                if (len == 0) break;
                if (ci >= 0 && (len < 0 || (int64_t)ci + len > colour_cells)) { out.bad_column = i; return false; }
            }
            h.w = (uint32_t)out.bounds.size();
            int64_t y = dim_y;
            bool ok = true;
            for (uint32_t k = 0; k < rc; k++) {
                const uint32_t el = elements[(int64_t)off + 1 + k];
                const int len = (int)(int16_t)(el >> 16);
                out.bounds.push_back(cvxh_u2{(uint32_t)(y < 0 ? 0 : y), el});
                if (len <= 0) ok = false;
                y -= (int64_t)len * scale;
                if (y < 0) ok = false;
            }
            out.bounds.push_back(cvxh_u2{(uint32_t)(y < 0 ? 0 : y), 0u});
            if (y != 0) ok = false;
            if (!ok) out.regular = false;
        }
        out.headers[(size_t)i] = h;
    }
    if (out.bounds.size() >= 0xffffffffull) out.regular = false;
    return true;
}
