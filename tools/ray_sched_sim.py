"""Developer tool (GPU box): is a frame's Phase-1 time set by its slowest rays or by its total work, and what would launching the
rays in a different order buy? Per pose: per-ray cycles (cvx_debug_ray_timing), the kernel's exclusive time, and a list-scheduling
simulation (3700 resident warps = 148 SMs x 25) of the launch order: flat index order (what the kernel does), longest first (oracle
LPT), and longest-first by the PREVIOUS pose's cost (what a renderer could know)."""
from __future__ import annotations

import argparse
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv  # noqa: E402


def makespan(costs, slots):
    h = [0.0] * min(slots, len(costs))
    heapq.heapify(h)
    for c in costs:
        heapq.heappush(h, heapq.heappop(h) + c)
    return max(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="3840x2160")
    ap.add_argument("--poses", default="0,6,12,18,24,30,36,42,48,54,59")
    ap.add_argument("--slots", type=int, default=148 * 25)
    a = ap.parse_args()
    W, H = [int(x) for x in a.res.split("x")]
    world = cv.World.from_obj(os.path.join(ROOT, "tests", "data", "mill.obj"), 1024)
    poses = cv.benchmark_path(world.dims, 60, far_clip=2.0 * world.max_dimension)
    rm = cv.RenderManager(0)
    rm.upload_world(world)
    rm.set_resolution(W, H)
    for i in [int(x) for x in a.poses.split(",")]:
        s = rm.make_setup(poses[i])
        best = 1e9
        for _ in range(3):
            rm.draw_setup(s); rm.sync()
            best = min(best, rm.last_draw_ms()[0])
        t = rm.ray_timing(s).astype(np.float64).sum(axis=1)
        n = len(t)
        order_lpt = np.argsort(-t)
        ms = lambda cyc: cyc / 1.965e6
        inorder, lpt = makespan(t, a.slots), makespan(t[order_lpt], a.slots)
        # interleaved: heavy clusters spread by a fixed stride permutation (no knowledge needed)
        stride = np.arange(n).reshape(-1, 1)
        perm = np.argsort((np.arange(n) * 2654435761) % n, kind="stable")
        strided = makespan(t[perm], a.slots)
        heavy_pos = np.mean(np.argsort(-t)[: max(1, n // 100)]) / n
        print(f"pose {i:2d}: rays {n:5d} kernel {best:.3f} ms | TIMING build: sum/slots {ms(t.sum() / min(a.slots, n)):.3f} ms  max ray {ms(t.max()):.3f} ms | sim makespan "
              f"in-order {ms(inorder):.3f}  LPT {ms(lpt):.3f}  hashed order {ms(strided):.3f} | top-1% rays sit at {heavy_pos:.2f} of the index range", flush=True)


if __name__ == "__main__":
    main()
