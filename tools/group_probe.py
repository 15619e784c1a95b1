"""Developer tool (GPU box, under ncu): one frame of BASELINE config 2 (terrain 2048^3, 1080p, pitched down) with 32, 16 and 8 lanes per ray.
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum -k regex:phase1 python tools/group_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv

world = cv.World.synthetic(0, (2048, 2048, 2048), seed=1234)
rm = cv.RenderManager(0)
rm.upload_world(world)
W, H = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '1920x1080').split('x')]
rm.set_resolution(W, H)
pose = cv.CameraPose.from_euler((1024.0, 1700.0, 1024.0), (60.0, 30.0, 0.0), far_clip=4096.0)
s = rm.make_setup(pose)
for g in (32, 16, 8):
    rm.set_group_size(g)
    for _ in range(2):
        rm.draw_setup(s); rm.sync()
    print("group", g, "phase1 %.3f ms" % rm.last_draw_ms()[0], flush=True)
