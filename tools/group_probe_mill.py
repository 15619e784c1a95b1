"""Developer tool (GPU box, under ncu): mill 1024^3 at 4K, poses 30 and 59, with 32, 16 and 8 lanes per ray (instruction counts, active threads per instruction)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv

world = cv.World.from_obj(os.path.join(ROOT, "tests", "data", "mill.obj"), 1024)
rm = cv.RenderManager(0)
rm.upload_world(world)
rm.set_resolution(3840, 2160)
poses = cv.benchmark_path(world.dims, 60, far_clip=2.0 * world.max_dimension)
for i in (5, 30, 59):
    s = rm.make_setup(poses[i])
    for g in (32, 16, 8):
        rm.set_group_size(g)
        rm.draw_setup(s); rm.sync()
        print("pose", i, "group", g, "phase1 %.3f ms" % rm.last_draw_ms()[0], flush=True)
