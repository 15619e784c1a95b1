"""Developer tool (GPU box): lanes per ray x views in flight on BASELINE config 2 (terrain 2048^3, 1080p, pitched down) with batches long enough
(64 views = the 16 yaws four times) that the drain of a batch does not dominate: frames/s of cvx_draw_batch, device resident."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv

world = cv.World.synthetic(0, (2048, 2048, 2048), seed=1234)
rm = cv.RenderManager(0)
rm.upload_world(world)
rm.set_resolution(1920, 1080)
poses = [cv.CameraPose.from_euler((1024.0, 1700.0, 1024.0), (60.0, 30.0 + 22.5 * i, 0.0), far_clip=4096.0) for i in range(16)] * 4
setups = [rm.make_setup(p) for p in poses]
for g in (32, 16, 8):
    rm.set_group_size(g)
    for k in (6, 12, 16):
        rm.set_frames_in_flight(k)
        rm.draw_batch(setups); rm.sync()
        t0 = time.perf_counter()
        for _ in range(3):
            rm.draw_batch(setups)
        rm.sync()
        dt = time.perf_counter() - t0
        print(f"lanes per ray {g:2d}, {k:2d} views in flight: {3 * len(setups) / dt:7.1f} frames/s", flush=True)
