"""Developer tool (GPU box): compare the CUDA path with the CPU oracle over the benchmark path and print timings.
Usage: python tools/gpu_check.py [--res 1920x1080] [--frames 12] [--maxdim 1024] [--synthetic]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cpuvox_b200 as cv  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--maxdim", type=int, default=1024)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--out", default="gpurun_out/gpu_check.json")
    a = ap.parse_args()
    W, H = [int(x) for x in a.res.split("x")]
    t0 = time.time()
    if a.synthetic:
        world = cv.World.synthetic(0, (a.maxdim,) * 3, seed=1234)
    else:
        path = os.path.join(ROOT, "tests", "data", "mill.obj")
        world = cv.World.from_obj(path, a.maxdim)
    print(f"world {world.dims} built in {time.time() - t0:.2f}s, voxels {world.voxel_counts}", flush=True)
    rm = cv.RenderManager(0, counters=True)
    rm.upload_world(world)
    rm.set_resolution(W, H)
    rm.set_group_size(a.group)
    ow = None if a.no_oracle else orc.OracleWorld(world.dims, world.blobs, world.column_counts)
    poses = cv.benchmark_path(world.dims, a.frames, far_clip=2.0 * world.max_dimension)
    results = []
    for i, pose in enumerate(poses):
        setup = rm.make_setup(pose)
        rm.clear_raybuffers(0)
        rm.counters()
        rm.draw_setup(setup)
        rm.sync()
        p1, p2 = rm.last_draw_ms()
        cn = rm.counters()
        # warm timing with the counters compiled out (the product configuration)
        rm.set_counters(False)
        q1 = q2 = 1e9
        for _ in range(2):
            rm.draw_setup(setup)
            rm.sync()
            t1, t2 = rm.last_draw_ms()
            q1, q2 = min(q1, t1), min(q2, t2)
        rm.set_counters(True)
        rec = {"frame": i, "rays": cn["rays"], "p1_ms": p1, "p2_ms": p2, "p1_warm_ms": q1, "p2_warm_ms": q2, "counters": cn}
        if ow is not None:
            frame = rm.read_frame()
            td, lr = rm.read_raybuffers()
            os_ = orc.copy_setup(setup)
            t1 = time.time()
            otd, olr, ocn = orc.render_raybuffers(ow, os_, W, H)
            oframe = orc.blit(os_, W, H, otd, olr)
            rec["oracle_s"] = time.time() - t1
            # compare only what this frame wrote (buffers were cleared to 0 on both sides)
            rec["td_mismatch"] = int((td != otd).sum())
            rec["lr_mismatch"] = int((lr != olr).sum())
            rec["frame_mismatch"] = int((frame != oframe).sum())
            rec["counters_equal"] = cn == ocn
            if not rec["counters_equal"]:
                rec["oracle_counters"] = ocn
        results.append(rec)
        if a.verbose:
            print(json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(results, f, indent=1)
    tot = sum(r["p1_warm_ms"] + r["p2_warm_ms"] for r in results)
    bad = [r["frame"] for r in results if r.get("td_mismatch") or r.get("lr_mismatch") or r.get("frame_mismatch") or not r.get("counters_equal", True)]
    print(f"group {a.group}: mean frame {tot / len(results):.3f} ms -> {1000.0 * len(results) / tot:.1f} fps (warm, counters off); max p1 {max(r['p1_warm_ms'] for r in results):.3f} ms; parity-bad frames: {bad}")


if __name__ == "__main__":
    main()
