"""Developer tool (GPU box): where the end-to-end leg of bench.py loses against the device-resident one. Host-timed steps of 60 views at 4K:
(a) cvx_draw_batch of ready-made setups, no copies; (b) cvx_draw_world_batch (host setup inside), no copies; (c) the same with frames
delivered to pinned host memory; (d) as (c) with 120 / 240 views per call (the drain at the end of a call amortised)."""
from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv  # noqa: E402

W, H = 3840, 2160
world = cv.World.from_obj(os.path.join(ROOT, "tests", "data", "mill.obj"), 1024)
poses = cv.benchmark_path(world.dims, 60, far_clip=2.0 * world.max_dimension)
rm = cv.RenderManager(0)
rm.upload_world(world)
rm.set_resolution(W, H)
setups = [rm.make_setup(p) for p in poses]
pinned = cv.alloc_pinned((240, H, W))


def timed(fn, views, steps=6):
    fn(); fn(); rm.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
        rm.sync()
    dt = time.perf_counter() - t0
    return views * steps / dt


for k in (6, 8):
    rm.set_frames_in_flight(k)
    print(f"in flight {k}:")
    print("  (a) draw_batch(setups), device only      %.1f frames/s" % timed(lambda: rm.draw_batch(setups), 60), flush=True)
    print("  (b) draw_world_batch(poses), device only %.1f frames/s" % timed(lambda: rm.draw_world_batch(poses), 60), flush=True)
    print("  (c) draw_world_batch(poses, pinned)      %.1f frames/s" % timed(lambda: rm.draw_world_batch(poses, pinned[:60]), 60), flush=True)
    print("  (d) 120 views per call, pinned           %.1f frames/s" % timed(lambda: rm.draw_world_batch(poses * 2, pinned[:120]), 120, 3), flush=True)
    print("  (d) 240 views per call, pinned           %.1f frames/s" % timed(lambda: rm.draw_world_batch(poses * 4, pinned), 240, 2), flush=True)
    print("  (e) 240 views per call, device only      %.1f frames/s" % timed(lambda: rm.draw_world_batch(poses * 4), 240, 2), flush=True)
