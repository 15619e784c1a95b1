"""Developer tool (GPU box): render single poses of the benchmark path repeatedly (no oracle) — the command ncu wraps.
Usage: python tools/one_frame.py [--res 1920x1080] [--poses 0,59] [--reps 3] [--maxdim 1024] [--synthetic DIM]"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cpuvox_b200 as cv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--poses", default="0,59")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--maxdim", type=int, default=1024)
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--frames", type=int, default=60)
    a = ap.parse_args()
    W, H = [int(x) for x in a.res.split("x")]
    if a.synthetic:
        world = cv.World.synthetic(0, (a.synthetic,) * 3, seed=1234)
    else:
        world = cv.World.from_obj(os.path.join(ROOT, "tests", "data", "mill.obj"), a.maxdim)
    rm = cv.RenderManager(0, counters=False)
    rm.upload_world(world)
    rm.set_resolution(W, H)
    poses = cv.benchmark_path(world.dims, a.frames, far_clip=2.0 * world.max_dimension)
    for i in [int(x) for x in a.poses.split(",")]:
        setup = rm.make_setup(poses[i])
        best = 1e9
        for _ in range(a.reps):
            rm.draw_setup(setup)
            rm.sync()
            p1, p2 = rm.last_draw_ms()
            best = min(best, p1)
        print(f"pose {i}: phase1 best {best:.3f} ms, phase2 {p2:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
