"""Developer tool (GPU box): per-ray cycle breakdown of Phase 1 (cvx_debug_ray_timing) for poses of the benchmark path."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv  # noqa: E402

NAMES = ["other", "walk+hdr", "select", "renarrow", "runs", "geometry", "resolve+ctl", "sky", "retest", "horizon", "cap-px", "side-setup", "side-px", "mark", "-", "-"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--poses", default="0,12,26,38,48,59")
    ap.add_argument("--maxdim", type=int, default=1024)
    ap.add_argument("--config2", action="store_true", help="fBm terrain 2048^3, camera (1024,1700,1024) pitch 60 (BASELINE config 2) instead of the mill path")
    a = ap.parse_args()
    W, H = [int(x) for x in a.res.split("x")]
    if a.config2:
        world = cv.World.synthetic(0, (2048, 2048, 2048), seed=1234)
        poses = [cv.CameraPose.from_euler((1024.0, 1700.0, 1024.0), (60.0, 30.0 + 22.5 * i, 0.0), far_clip=4096.0) for i in range(16)]
    else:
        world = cv.World.from_obj(os.path.join(ROOT, "tests", "data", "mill.obj"), a.maxdim)
        poses = cv.benchmark_path(world.dims, 60, far_clip=2.0 * world.max_dimension)
    rm = cv.RenderManager(0)
    rm.upload_world(world)
    rm.set_resolution(W, H)
    for i in [int(x) for x in a.poses.split(",")]:
        s = rm.make_setup(poses[i])
        rm.draw_setup(s); rm.sync()
        t = rm.ray_timing(s).astype(np.float64)
        tot = t.sum(axis=1)
        k = int(np.argmax(tot))
        print(f"pose {i}: rays {len(t)}  sum-cycles {tot.sum():.3e}  mean/ray {tot.mean():.0f}  max/ray {tot.max():.0f} (ray {k})")
        print("   share of all cycles : " + "  ".join(f"{n} {100 * t[:, j].sum() / tot.sum():.1f}%" for j, n in enumerate(NAMES)))
        print("   slowest ray         : " + "  ".join(f"{n} {100 * t[k, j] / tot[k]:.1f}%" for j, n in enumerate(NAMES)))
        top = np.sort(tot)[::-1]
        print("   ray cycles pctl     : " + "  ".join(f"p{p} {np.percentile(tot, p):.0f}" for p in (50, 90, 99)) + f"  top5 {top[:5].astype(int).tolist()}")


if __name__ == "__main__":
    main()
