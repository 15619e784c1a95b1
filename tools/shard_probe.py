"""Developer tool (ONE GPU): what one rank of an N-GPU ray-sharded view costs. A ring for N ranks is created in one process and the share
of every rank r is rendered alone (cvx_draw_sharded, CUDA events around its Phase 1 and its owned-pixel Phase 2), for BASELINE config 4
(8K) or the mill at 4K:  python tools/shard_probe.py [--config 4|1] [--chunk 512] [--ranks 1,2,4,8]"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cpuvox_b200 as cv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--chunk", default="512")
    ap.add_argument("--ranks", default="1,2,4,8")
    ap.add_argument("--poses", default="0,1,2,3")
    a = ap.parse_args()
    b = types.SimpleNamespace(config=a.config, res={1: "3840x2160", 4: "7680x4320"}[a.config], maxdim=1024)
    world, poses, name, _ = bench.make_workload(cv, b)
    W, H = [int(x) for x in b.res.split("x")]
    rm = cv.RenderManager(0)
    rm.upload_world(world)
    rm.set_resolution(W, H)
    rm.set_frames_in_flight(1)
    print(name, flush=True)
    for pi in [int(x) for x in a.poses.split(",")]:
        setup = rm.make_setup(poses[pi])
        for chunk in [int(x) for x in a.chunk.split(",")]:
            for n in [int(x) for x in a.ranks.split(",")]:
                rm.ring_create(8, n)
                p1s, p2s = [], []
                view = 0
                for r in range(n):
                    best = (1e9, 1e9)
                    for _ in range(2):
                        rm.profile_begin(1)
                        rm.draw_sharded(setup, -1, chunk, view % 8, r)   # view < ring slots: no wait for a release
                        rm.sync()
                        p1, p2, _ = rm.profile_end()
                        if p1 + p2 < sum(best):
                            best = (p1, p2)
                    p1s.append(best[0]); p2s.append(best[1])
                rm.ring_close()
                p1s, p2s = np.array(p1s), np.array(p2s)
                tot = p1s + p2s
                print(f"pose {pi} chunk {chunk:5d} N={n}: per-rank P1 mean {p1s.mean():.3f} max {p1s.max():.3f} ms | P2 owned mean {p2s.mean():.3f} max {p2s.max():.3f} ms | "
                      f"share mean {tot.mean():.3f} max {tot.max():.3f} ms  (sum of shares / N = {tot.sum() / n:.3f})", flush=True)


if __name__ == "__main__":
    main()
