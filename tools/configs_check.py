"""Developer tool (GPU box): BASELINE.json configs 2-5 at FULL size — parity against the oracle on a sample of each config's
views plus throughput, frames/s and voxel-runs/s with the HBM-roofline fraction of SURVEY.md §8(d).

    python tools/configs_check.py [--configs 2,3,4,5] [--out gpurun_out/configs.jsonl]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/configs_check.py

config 2  fBm terrain 2048^3, 1920x1080, camera at the world centre, y 1700, pitch 60 down (VP on screen, 4 segments); 16 yaws
config 3  same world, 3840x2160, camera 40 above the ground, pitch 3 and pitch 0 (LimitRotationHorizon), clamped segments; 8 yaws each
config 4  boxes/pipes/slabs 4096x1024x4096, 7680x4320, far 8192; N > 1: the rays of each view sharded over the ranks, peer-store gather
config 5  256 cameras (seed 99) at 1280x720 over the config-2 world; N > 1: views sharded
One JSON line per config on rank 0. Single-GPU timing: CUDA events around cvx_draw_batch passes (device resident, L2 flushed
between passes); sharded single views: host clock around draw + barrier (max over ranks by construction)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ground_height(world, x, z):
    """worldMax of LOD-0 column (x, z): uint16 at byte 8 of its 12-byte RLEColumn header (World.cs:163-168)."""
    hdr = world.blobs[0][: 12 * world.column_counts[0]].view(np.uint32).reshape(-1, 3)
    return int(hdr[int(x) * world.dims[2] + int(z), 2] & 0xFFFF)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2,3,4,5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--check", type=int, default=2, help="views per config compared with the oracle")
    a = ap.parse_args()
    todo = [int(c) for c in a.configs.split(",")]

    import torch
    import torch.distributed as dist

    import cpuvox_b200 as cv
    from oracle import oracle as orc   # checker only

    rank, local, n = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if n > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = []

    def get_world(kind, dims, seed):
        w = cv.World.synthetic(kind, dims, seed=seed) if rank == 0 else None
        return cv.broadcast_world(w, src=0, device=dev) if n > 1 else w

    def counters_for(rm, setups):
        rm.set_counters(True)
        rm.counters()
        tot = None
        for s in setups:
            rm.draw_setup(s)
            c = rm.counters()
            tot = c if tot is None else {k: tot[k] + c[k] for k in c}
        rm.set_counters(False)
        return tot

    def parity(rm, world, setups, W, H, label):
        if rank != 0:
            return True
        ow = orc.OracleWorld(world.dims, world.blobs, world.column_counts)
        ok = True
        idx = sorted(set(np.linspace(0, len(setups) - 1, min(a.check, len(setups))).astype(int).tolist()))
        for i in idx:
            rm.clear_raybuffers(0)   # pixels outside the writable ranges are never touched: compare against zero-filled oracle buffers
            rm.draw_setup(setups[i])
            rm.sync()
            g = rm.read_frame()
            gtd, glr = rm.read_raybuffers()
            os_ = orc.copy_setup(setups[i])
            otd, olr, _ = orc.render_raybuffers(ow, os_, W, H)
            of = orc.blit(os_, W, H, otd, olr)
            same = np.array_equal(g, of) and np.array_equal(gtd, otd) and np.array_equal(glr, olr)
            if not same:
                print(f"MISMATCH {label} view {i}: {int((g != of).sum())} frame pixels", flush=True)
            ok &= same
        return ok

    def time_batch(rm, setups):
        stream = torch.cuda.Stream(local)
        torch.cuda.set_stream(stream)
        rm.set_stream(stream.cuda_stream)
        for _ in range(2):
            flush.zero_()
            rm.draw_batch(setups)
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.passes):
            flush.zero_()
            rm.draw_batch(setups)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if n > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / a.passes   # ms per pass (includes the 256 MiB flush write, ~0.05 ms)

    def report(cfg, desc, W, H, views_total, ms_per_pass, ctr_local, views_local, ok, extra=None):
        # counters are summed over the views THIS rank rendered; ranks render statistically alike shares, so scale to the job
        scale = views_total / max(1, views_local)
        fps = views_total / (ms_per_pass / 1000.0)
        bytes_total = cv.algorithmic_bytes(ctr_local, W, H, views_local) * scale
        line = {"config": cfg, "workload": desc, "resolution": [W, H], "n_gpus": n, "views": views_total, "parity_vs_oracle": "bit-exact" if ok else "MISMATCH",
                "frames_per_s": fps, "ms_per_frame": ms_per_pass / views_total, "runs_per_s": ctr_local["runs_visited"] * scale / (ms_per_pass / 1000.0),
                "runs_per_frame": ctr_local["runs_visited"] / max(1, views_local), "dda_steps_per_frame": ctr_local["dda_steps"] / max(1, views_local),
                "algorithmic_bytes_per_frame": bytes_total / views_total,
                "roofline": {"bound": "hbm", "achieved_gbs": bytes_total / (ms_per_pass / 1000.0) / 1e9, "peak_gbs_per_gpu": peak,
                             "frac": bytes_total / (ms_per_pass / 1000.0) / 1e9 / (peak * n)}}
        if extra:
            line.update(extra)
        if rank == 0:
            print(json.dumps(line), flush=True)
            lines.append(line)

    terrain = None
    if any(c in todo for c in (2, 3, 5)):
        terrain = get_world(0, (2048, 2048, 2048), 1234)
    rm = cv.RenderManager(local)

    if 2 in todo:
        W, H = 1920, 1080
        rm.upload_world(terrain)
        rm.set_resolution(W, H)
        poses = [cv.CameraPose.from_euler((1024.0, 1700.0, 1024.0), (60.0, 30.0 + 22.5 * i, 0.0), far_clip=4096.0) for i in range(16)]
        mine = cv.partition_views(len(poses), n, rank)
        setups = [rm.make_setup(poses[i]) for i in mine]
        assert all(s.segments[k].ray_count > 0 for s in setups for k in range(4))
        ok = parity(rm, terrain, setups, W, H, "config2")
        ctr = counters_for(rm, setups)
        ms = time_batch(rm, setups)
        report(2, "fBm terrain 2048^3 (seed 1234), camera (1024,1700,1024) pitch 60 down, 16 yaws, 4 segments", W, H, len(poses), ms, ctr, len(setups), ok)

    if 3 in todo:
        W, H = 3840, 2160
        rm.upload_world(terrain)
        rm.set_resolution(W, H)
        h = ground_height(terrain, 1024, 1024)
        for pitch in (3.0, 0.0):
            poses = [cv.CameraPose.from_euler((1024.5, h + 40.0, 1024.5), (pitch, 30.0 + 45.0 * i, 0.0), far_clip=4096.0) for i in range(8)]
            mine = cv.partition_views(len(poses), n, rank)
            setups = [rm.make_setup(poses[i]) for i in mine]
            segs = sorted({sum(1 for k in range(4) if s.segments[k].ray_count > 0) for s in setups})
            ok = parity(rm, terrain, setups, W, H, f"config3 pitch {pitch}")
            ctr = counters_for(rm, setups)
            ms = time_batch(rm, setups)
            report(3, f"fBm terrain 2048^3, camera 40 above ground (y={h + 40}), pitch {pitch} (forward.y clamped to +-0.001 when 0), 8 yaws", W, H,
                   len(poses), ms, ctr, len(setups), ok, {"active_segments": segs, "sharding": "views" if n > 1 else "none"})

    if 5 in todo:
        W, H = 1280, 720
        rm.upload_world(terrain)
        rm.set_resolution(W, H)
        rng = np.random.default_rng(99)
        poses = []
        for _ in range(256):
            x, z = rng.uniform(0, 2048, 2)
            hh = ground_height(terrain, min(2047, x), min(2047, z))
            y = rng.uniform(hh + 20.0, max(hh + 21.0, 1900.0))
            poses.append(cv.CameraPose.from_euler((float(x), float(y), float(z)), (float(rng.uniform(-30, 80)), float(rng.uniform(0, 360)), 0.0), far_clip=4096.0))
        mine = cv.partition_views(len(poses), n, rank)
        setups = [rm.make_setup(poses[i]) for i in mine]
        ok = parity(rm, terrain, setups, W, H, "config5")
        ctr = counters_for(rm, setups)
        ms = time_batch(rm, setups)
        report(5, "256 cameras (seed 99) over the fBm terrain 2048^3, y in [ground+20, 1900], pitch in [-30, 80], no roll; views sharded", W, H,
               len(poses), ms, ctr, len(setups), ok)
    rm.destroy()
    terrain = None

    if 4 in todo:
        W, H = 7680, 4320
        plant = get_world(1, (4096, 1024, 4096), 7)
        poses = [cv.CameraPose.from_euler((2048.5, 700.5, 2048.5), (35.0, 20.0, 0.0), far_clip=8192.0),
                 cv.CameraPose.from_euler((500.5, 400.5, 700.5), (12.0, 50.0, 0.0), far_clip=8192.0),
                 cv.CameraPose.from_euler((3000.5, 950.5, 1000.5), (70.0, 200.0, 10.0), far_clip=8192.0),
                 cv.CameraPose.from_euler((2048.5, 300.5, 100.5), (-10.0, 0.0, 0.0), far_clip=8192.0)]
        srm = cv.ShardedRenderManager(local, rank, n)
        srm.upload_world(plant)
        srm.set_resolution(W, H)
        rm = srm.rm
        setups = [rm.make_setup(p) for p in poses]
        ok = True
        ctr = None
        if rank == 0:
            ok = parity(rm, plant, setups, W, H, "config4")   # single-GPU frames vs the oracle
            ctr = counters_for(rm, setups)
        times = []
        for rep in range(a.passes + 1):
            for i, p in enumerate(poses):
                if n > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                srm.draw_world_sharded(p)
                rm.sync()
                dt = time.perf_counter() - t0
                if rep > 0:
                    times.append(dt)
                if rep == 0 and rank == 0 and n > 1:   # sharded frame == single-GPU frame, bit for bit
                    got = srm.read_frame().copy()
                    rm.draw_setup(setups[i])
                    rm.sync()
                    same = np.array_equal(got, rm.read_frame())
                    ok &= same
                    if not same:
                        print(f"MISMATCH config4 sharded view {i}", flush=True)
        t = torch.tensor([sum(times) / a.passes * 1000.0], dtype=torch.float64, device=dev)
        if n > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            report(4, f"boxes/pipes/slabs 4096x1024x4096 (seed 7, {plant.voxel_counts[0]} voxels), 4 views, far 8192" +
                   ("; rays of each view sharded, peer-store gather into rank 0" if n > 1 else ""), W, H, len(poses), float(t.item()), ctr, len(poses), ok,
                   {"timing": "host clock around draw + sync (+ barrier), one view at a time", "sharding": "rays" if n > 1 else "none"})
        srm.destroy()

    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "a") as fh:
            for ln in lines:
                fh.write(json.dumps(ln) + "\n")
        bad = [ln["config"] for ln in lines if ln["parity_vs_oracle"] != "bit-exact"]
        print("CONFIGS CHECK " + ("PASSED" if not bad else f"FAILED {bad}"), flush=True)
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
