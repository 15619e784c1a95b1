/*
 * host_frame.h — host-side construction of the per-frame kernel parameters (cvxd_frame) from the caller's
 * cvx_frame_setup: RenderManager.DrawSegments' context fill (Assets/Code/RenderManager.cs:281-318).
 * Header-only so that the C ABI (capi.cu) and the test-only SIMT emulator (tools/simt_emu) share one definition.
 */
#pragma once
#include <math.h>
#include <string.h>

#include "../../include/cpuvox_b200.h"
#include "device_types.h"

namespace cvxh {

inline int f2i(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000; }
inline int clampi(int x, int a, int b) { return x < a ? a : (x > b ? b : x); }

// RenderManager.DrawSegments context fill, RenderManager.cs:281-318
inline int fill_segments(const cvx_frame_setup* s, int W, int H, cvxd_segment out[4]) {
    int total = 0;
    const float vx = s->vanishing_point_screen[0], vy = s->vanishing_point_screen[1];
    for (int k = 0; k < 4; k++) {
        cvxd_segment& c = out[k];
        memset(&c, 0, sizeof c);
        const cvx_segment& in = s->segments[k];
        c.ray_count = in.ray_count;
        total += in.ray_count > 0 ? in.ray_count : 0;
        for (int i = 0; i < 2; i++) {
            c.ray_min[i] = in.cam_local_plane_ray_min[i]; c.ray_max[i] = in.cam_local_plane_ray_max[i];
            c.min_screen[i] = in.min_screen[i]; c.max_screen[i] = in.max_screen[i];
        }
        if (in.ray_count <= 0) continue;
        c.axis_mapped_to_y = k > 1 ? 0 : 1;
        c.ray_index_offset = k == 1 ? s->segments[0].ray_count : (k == 3 ? s->segments[2].ray_count : 0);
        if (k < 2) {
            c.buffer = 0;
            int v = clampi(f2i(rintf(vy)), 0, H - 1); // Mathf.RoundToInt, half-to-even
            c.pix_min = k == 0 ? v : 0;
            c.pix_max = k == 0 ? H - 1 : v;
        } else {
            c.buffer = 1;
            int v = clampi(f2i(rintf(vx)), 0, W - 1);
            c.pix_min = k == 3 ? 0 : v;
            c.pix_max = k == 3 ? v : W - 1;
        }
    }
    return total;
}

// everything of cvxd_frame that does not depend on the context's device buffers
inline void frame_from_setup(const cvx_frame_setup* s, int W, int H, cvxd_frame& f) {
    memset(&f, 0, sizeof f);
    memcpy(f.wts, s->camera.world_to_screen, sizeof f.wts);
    f.pos_x = s->camera.position_xz[0]; f.pos_z = s->camera.position_xz[1]; f.pos_y = s->camera.position_y;
    f.inverse = s->camera.inverse_element_iteration_direction ? 1 : 0;
    f.far_clip = s->camera.far_clip;
    memcpy(f.lod_dist, s->camera.lod_distances, sizeof f.lod_dist);
    f.total_rays = fill_segments(s, W, H, f.seg);
    f.vp_x = s->vanishing_point_screen[0]; f.vp_y = s->vanishing_point_screen[1];
    f.width = W; f.height = H;
    f.ray_begin = 0; f.ray_end = f.total_rays;
}

} // namespace cvxh
