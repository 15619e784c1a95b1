"""bench.py — the headline benchmark of the raybuffer path (BASELINE.json: frames/sec on mill 1024^3 at 1080p & 4K).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--res 3840x2160]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): datasets/mill.obj voxelized to 1024^3 (+5 LOD mips) by the voxelizer restatement, rendered
along the reference's own benchmark camera path (BenchmarkPath.anim sampled at 60 evenly spaced times, UnityManager.cs:86-87)
at 3840x2160 — BASELINE config 1 at the north_star's headline resolution; 1080p is reported beside it under "at_1080p".
A step = one pass over the 60 poses (60 frames). Synthetic data only in the sense of the fixed camera path; the world is
the reference's shipped dataset.

value    frames/s with everything resident in HBM: per frame one Phase-1 and one Phase-2 launch; a step is one
         cvx_draw_batch over the 60 poses (device only), which keeps up to `frames_in_flight` views in flight, each on its own
         stream with its own raybuffers and framebuffer (the reference double-buffers its raybuffers for the same overlap).
e2e      frames/s through the public C ABI with HOST buffers (cvx_draw_world_batch = RenderManager.DrawWorld per camera): per
         frame the host computes the segment/VP setup from the camera pose, passes it by value (kernel parameters are the only
         host->device bytes) and receives the finished frame in pinned host memory (copy overlapped with the next frames' kernels).
roofline Phase-1 kernel (dominant): algorithmic bytes of SURVEY.md §8(d) per launch / launch duration against the measured
         HBM copy bandwidth of MEASURED_PEAKS.json. Launches of different views overlap, so the duration used is the timed
         region's wall time per frame times Phase 1's share of the kernel time; the share comes from the "exclusive" pass, which
         repeats the launches with one view in flight (each launch alone on the GPU, CUDA events around every launch).
cpu_baseline / --impl reference: the CPU restatement of the reference's path (oracle/, "port": the reference is C# on
         Unity/Burst and cannot be built here) on all host threads.
N > 1    views are sharded (each rank renders the whole path for its own share of a global batch of N x 60 views), the
         world is built on rank 0, broadcast once over NCCL and replicated; no collective on the data path ("weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MILL = os.path.join(ROOT, "tests", "data", "mill.obj")
FRAMES_PER_STEP = 60
METRIC = "frames/sec @4K, mill 1024^3, BenchmarkPath 60 poses"
L2_FLUSH_BYTES = 256 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--res", default="3840x2160")
    ap.add_argument("--maxdim", type=int, default=1024)
    ap.add_argument("--group", type=int, default=0, help="Phase-1 lanes per ray (0 = library default)")
    ap.add_argument("--inflight", type=int, default=6, help="views in flight per cvx_draw_batch (1..8)")
    ap.add_argument("--inflight-e2e", type=int, default=8, help="views in flight for the e2e leg (frame copies occupy the slots longer)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-1080p", action="store_true")
    return ap.parse_args()


def workload_name(maxdim, W, H):
    return (f"datasets/mill.obj voxelized to {maxdim}^3 (+5 LODs), {W}x{H}, {FRAMES_PER_STEP}-pose BenchmarkPath.anim camera path, "
            "FOV 85, near 0.05, far 2*maxdim, lodError 1")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(W):
    """DRAM bytes per Phase-1 launch from the committed ncu --set full capture, if one exists for this resolution."""
    p = os.path.join(ROOT, "profiles", "phase1_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p)).get(str(W))
        return t
    return None


def ncu_instructions(W):
    """Mean warp instructions per Phase-1 launch over the 60 poses, from the committed ncu pass (profiles/phase1_inst.json)."""
    p = os.path.join(ROOT, "profiles", "phase1_inst.json")
    if os.path.exists(p):
        return json.load(open(p)).get(str(W))
    return None


def phase1_bytes(c):
    """Phase-1 share of the algorithmic bytes: headers, runs, colour gather, raybuffer writes."""
    return 12 * c["dda_steps"] + 4 * c["runs_visited"] + 4 * c["px_voxel"] + 4 * (c["px_voxel"] + c["px_sky"])


# ---------------------------------------------------------------------------------------------------------------------
def parse_obj_soup(path):
    """ObjModel.Import (ObjModel.cs:10-167) for `v x y z [r g b]` / `f a b c` files: the triangle soup it builds (one vertex per
    face corner, indices 0..n-1), colours through Unity's Color -> Color32 rounding. Pure Python: the CPU arm must not need the product."""
    vs, cs, tri = [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                q = line.split()
                vs.append([float(q[1]), float(q[2]), float(q[3])])
                cs.append([float(q[4]), float(q[5]), float(q[6])] if len(q) > 6 else [1.0, 1.0, 1.0])
            elif line.startswith("f "):
                tri.append([int(x.split("/")[0]) - 1 for x in line.split()[1:4]])
    vs, cs, tri = np.array(vs, dtype=np.float32), np.array(cs, dtype=np.float32), np.array(tri).ravel()
    col = np.rint(np.clip(cs, 0, 1) * np.float32(255)).astype(np.uint8)
    col = np.concatenate([col, np.full((len(col), 1), 255, np.uint8)], 1)
    return vs[tri], col[tri]


class CpuPath:
    """The reference's CPU implementation of the path, for `--impl reference` and the cpu_baseline leg. No product code:
    kind "reference": oracle/_ref — the reference's own C# (voxelizer, RLE builder, DownSample, segment setup, DrawSegments + the four
        jobs of DrawSegmentRayJob) translated to C++ and compiled (oracle/ref.py; no C# toolchain exists here), jobs spread over
        all host threads with a shared batch counter like Unity's IJobParallelFor;
    kind "port": oracle/cpuvox_oracle.cpp (hand restatement), only when the prebuilt oracle/_ref library is missing.
    Phase 2 is a GPU shader in the reference (RayBufferBlit.shader); on the CPU arm it is the oracle's per-pixel restatement
    (threaded), so that a CPU "frame" is the same product as a GPU frame. Camera path, LOD distances and LimitRotationHorizon
    (UnityEngine / MonoBehaviour code) come from the oracle library's restatements."""

    def __init__(self, maxdim, W, H, frames=FRAMES_PER_STEP):
        from oracle import oracle as orc
        from oracle import ref
        self.orc, self.W, self.H = orc, W, H
        self.kind = "reference" if ref.build() else "port"
        self.threads = orc.hardware_threads()
        P, C = parse_obj_soup(MILL)
        if self.kind == "reference":
            self.ref = ref
            dims, blobs, ccs, _ = ref.build_world_from_mesh(P, C, np.arange(P.shape[0]), maxdim)   # flip X only: the UI default
            self.world = ref.RefWorld(dims, blobs, ccs)
            self.buffers = ref.alloc_raybuffers(W, H)
        else:
            import cpuvox_b200 as cv  # fallback only: the oracle has no world builder of its own
            w = cv.World.from_obj(MILL, maxdim)
            dims, blobs, ccs = w.dims, w.blobs, w.column_counts
            self.buffers = (np.zeros((W + 2 * H, H), dtype=np.uint32), np.zeros((2 * W + H, W), dtype=np.uint32))
        self.oworld = orc.OracleWorld(dims, blobs, ccs)
        self.dims = dims
        lods = orc.setup_lods(max(dims), W, H)
        self.setups = []
        for i in range(frames):
            pos, rot = orc.benchmark_pose(1.15 * i / (frames - 1), dims)          # BenchmarkPath.anim, clip length 1.15
            if self.kind == "reference":
                rot = orc.limit_rotation_horizon(pos, rot)
                self.setups.append(ref.frame_setup(pos, rot, W, H, lods, dims[1], far=2.0 * max(dims)))
            else:
                self.setups.append(orc.frame_setup(pos, rot, W, H, lods, dims[1], far=2.0 * max(dims)))
        self.frame = np.zeros((H, W), dtype=np.uint32)

    def render(self, i):
        s = self.setups[i]
        if self.kind == "reference":
            td, lr = self.ref.render_raybuffers(self.world, s, self.W, self.H, threads=0, buffers=self.buffers)
            s = self.orc.copy_setup(s)
        else:
            td, lr, _ = self.orc.render_raybuffers(self.oworld, s, self.W, self.H, threads=0, td=self.buffers[0], lr=self.buffers[1])
        self.orc.blit(s, self.W, self.H, td, lr, threads=0, frame=self.frame)

    def describe(self):
        what = ("reference C# translated to C++ (oracle/_ref: cs2cpp.py, g++ -O2 -ffp-contract=off; not .NET, not Burst)" if self.kind == "reference"
                else "hand C++ restatement (oracle/cpuvox_oracle.cpp)")
        return f"Phase 1 = {what}, Phase 2 = per-pixel restatement of the blit shader; {self.threads} host threads"


def run_reference(a, rank, world_size):
    """--impl reference: the reference's own CPU implementation on all host threads, the same 60 poses per step as the repo arm."""
    if rank != 0:
        return
    W, H = [int(x) for x in a.res.split("x")]
    cpu = CpuPath(a.maxdim, W, H)
    n = len(cpu.setups)

    def step():
        for i in range(n):
            cpu.render(i)

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    fps = n * a.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1000.0 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "datasets/mill.obj (reference dataset) voxelized in-process, fixed 60-pose camera path",
        "config": {"workload": workload_name(a.maxdim, W, H), "resolution": [W, H], "frames_per_step": n},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cpu.threads, "kind": cpu.kind,
                         "sample": f"all {n} poses of the path per step; " + cpu.describe()},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def time_path(torch, dist, rm, setups, steps, warmup, device, flush, world_size, profile=True):
    """W warm-up steps, then exactly `steps` steps between barrier+synchronize pairs; CUDA events on the launching stream."""
    def one_step():
        flush.zero_()            # L2 flush between steps: 256 MiB written on the same stream
        rm.draw_batch(setups)    # device only: enqueues all views, joined back into this stream

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize(device)
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    if profile:
        rm.profile_begin(steps * len(setups))
    launches0 = rm.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        one_step()
    e1.record()
    torch.cuda.synchronize(device)
    t1 = time.perf_counter()
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    p1 = p2 = 0.0
    n = 0
    if profile:
        p1, p2, n = rm.profile_end()
    return ms, p1, p2, n, rm.launch_count() - launches0, (t0, t1)


def time_e2e(torch, dist, cv, rm, poses, steps, warmup, device, world_size, pinned):
    """Public-API path with host buffers: host setup per frame, frames delivered to pinned host memory."""
    def one_step():
        # RenderManager.DrawWorld per camera: the host part (LimitRotationHorizon, vanishing point, segments, CameraData) is computed
        # inside the call from the poses; kernels + device->host frame copies; returns when all frames are on the host
        rm.draw_world_batch(poses, pinned)

    for _ in range(max(1, warmup // 2)):
        one_step()
    torch.cuda.synchronize(device)
    if world_size > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    if world_size > 1:
        dist.barrier()
    return dt


def run_b200(a, rank, local_rank, world_size):
    import torch
    import torch.distributed as dist

    import cpuvox_b200 as cv
    from cpuvox_b200 import native as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cpuvox_b200 has no CPU fallback (use --impl reference for the CPU baseline)")
    device = local_rank
    torch.cuda.set_device(device)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{device}"))

    W, H = [int(x) for x in a.res.split("x")]
    world = cv.World.from_obj(MILL, a.maxdim) if rank == 0 else None
    if world_size > 1:
        world = cv.broadcast_world(world, src=0, device=torch.device(f"cuda:{device}"))  # once, then replicated per GPU
    rm = cv.RenderManager(device)
    # one explicit stream for everything that is timed: the library's launches, the L2 flush and the torch events
    # (torch's default stream is the NULL handle, which cvx_set_stream reads as "use the context's own stream")
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    rm.set_stream(stream.cuda_stream)
    rm.upload_world(world)
    rm.set_group_size(a.group)
    rm.set_frames_in_flight(a.inflight)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{device}")
    poses = cv.benchmark_path(world.dims, FRAMES_PER_STEP, far_clip=2.0 * world.max_dimension)

    def counters_for(setups):
        rm.set_counters(True)
        rm.counters()
        per = []
        for s in setups:
            rm.draw_setup(s)
            per.append(rm.counters())
        rm.set_counters(False)
        return per

    results = {}
    resolutions = [(W, H)] + ([] if a.no_1080p or (W, H) == (1920, 1080) else [(1920, 1080)])
    clocks = None
    for ri, (w, h) in enumerate(resolutions):
        rm.set_resolution(w, h)
        setups = [rm.make_setup(p) for p in poses]
        per = counters_for(setups)
        sampler = ClockSampler(device) if (ri == 0 and rank == 0) else None
        steps = a.steps if ri == 0 else max(2, a.steps // 2)
        ms, p1, p2, n, launches, span = time_path(torch, dist, rm, setups, steps, a.warmup, device, flush, world_size)
        if sampler:
            clocks = sampler.stop(*span)
        # the same launches one view at a time: exclusive kernel durations (outside the timed region, reported beside it)
        rm.set_frames_in_flight(1)
        _, x1, x2, xn, _, _ = time_path(torch, dist, rm, setups, 1, 1, device, flush, world_size)
        rm.set_frames_in_flight(a.inflight)
        pinned = cv.alloc_pinned((len(poses), h, w))
        rm.set_frames_in_flight(a.inflight_e2e)   # a slot is busy with its device->host copy too: more slots keep the GPU fed
        e2e_s = time_e2e(torch, dist, cv, rm, poses, steps, a.warmup, device, world_size, pinned)
        rm.set_frames_in_flight(a.inflight)
        N.lib.cvx_free_pinned(pinned.ctypes.data)
        t = torch.tensor([ms, e2e_s * 1000.0, p1, p2, x1, x2], dtype=torch.float64, device=f"cuda:{device}")
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        ms, e2e_ms, p1, p2, x1, x2 = [float(x) for x in t.cpu()]
        results[(w, h)] = dict(ms=ms, e2e_ms=e2e_ms, p1=p1, p2=p2, n=n, launches=launches, steps=steps, per=per, x1=x1, x2=x2, xn=xn)

    main = results[(W, H)]
    steps = main["steps"]
    frames = FRAMES_PER_STEP * steps * world_size
    fps = frames / (main["ms"] / 1000.0)
    e2e_fps = frames / (main["e2e_ms"] / 1000.0)
    per = main["per"]
    runs_per_frame = sum(c["runs_visited"] for c in per) / len(per)
    p1_bytes = sum(phase1_bytes(c) for c in per) / len(per)             # mean algorithmic bytes per Phase-1 launch
    p1_ev = main["p1"] / max(1, main["n"])                              # mean Phase-1 launch duration by CUDA events (launches overlap)
    p2_ev = main["p2"] / max(1, main["n"])
    frame_ms = main["ms"] / (steps * FRAMES_PER_STEP)                    # timed-region wall time per frame on this rank
    x1_ms, x2_ms = main["x1"] / max(1, main["xn"]), main["x2"] / max(1, main["xn"])   # one view in flight: each launch alone
    # Phase 1's share of the step: from the EXCLUSIVE durations (the overlapped event durations of the short Phase-2 launches are
    # inflated by the Phase-1 launches of the other views they share the GPU with, which would flatter Phase 1)
    share = x1_ms / (x1_ms + x2_ms) if x1_ms + x2_ms > 0 else (p1_ev / (p1_ev + p2_ev) if p1_ev + p2_ev > 0 else 1.0)
    p1_ms, p2_ms = frame_ms * share, frame_ms * (1.0 - share)           # effective duration per launch under overlap
    peak, peak_src = measured_hbm_peak()
    achieved = p1_bytes / (p1_ms * 1e-3) / 1e9 if p1_ms > 0 else 0.0
    frame_bytes = sum(cv.algorithmic_bytes(c, W, H) for c in per) / len(per)
    # what actually bounds Phase 1: issue slots (148 SMs x 4 schedulers x SM clock); instruction counts come from the committed ncu pass
    inst = ncu_instructions(W)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue = None
    if inst and p1_ms > 0:
        slots_per_s = 148 * 4 * sm_mhz * 1e6
        issue = {"warp_instructions_per_launch": inst, "source": "profiles/phase1_inst.json (ncu smsp__inst_executed.sum, mean over the 60 poses)",
                 "issue_slot_utilisation": inst / (p1_ms * 1e-3) / slots_per_s,
                 "note": "Phase 1 is latency/issue bound (DESIGN.md §7): this, not the HBM fraction, tracks kernel quality"}

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world_size, "steps": steps, "warmup": a.warmup,
        "ms_per_step": main["ms"] / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "datasets/mill.obj (reference dataset) voxelized in-process, fixed 60-pose camera path",
        "config": {
            "workload": workload_name(a.maxdim, W, H), "resolution": [W, H], "frames_per_step": FRAMES_PER_STEP * world_size,
            "parallelism": "1 GPU" if world_size == 1 else f"views sharded over {world_size} GPUs, world broadcast once and replicated, no data-path collective",
            "l2": "flushed between steps (256 MiB device write inside the timed region); frames of one step run back to back",
            "phase1_lanes_per_ray": a.group or 32, "frames_in_flight": a.inflight, "frames_in_flight_e2e": a.inflight_e2e,
        },
        "runs_per_s": runs_per_frame * fps,
        "ms_per_frame": {"total": frame_ms, "phase1_kernel": p1_ms, "phase2_kernel": p2_ms,
                         "basis": "wall time of the timed region per frame, split by the kernels' share of the exclusive CUDA-event durations "
                                  f"(up to {a.inflight} views in flight, launches overlap)",
                         "event_mean_overlapped": {"phase1_kernel": p1_ev, "phase2_kernel": p2_ev},
                         "exclusive_one_view_in_flight": {"phase1_kernel": x1_ms, "phase2_kernel": x2_ms}},
        "roofline": {
            "bound": "hbm", "kernel": "phase1_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": p1_bytes, "traffic": ncu_traffic(W),
            "exclusive": {"launch_ms": x1_ms, "achieved": p1_bytes / (x1_ms * 1e-3) / 1e9 if x1_ms > 0 else 0.0,
                          "note": "same launches with one view in flight (each alone on the GPU), measured after the timed region"},
            "issue": issue,
            "whole_frame": {"algorithmic_bytes": frame_bytes, "achieved": frame_bytes * fps / world_size / 1e9,
                            "frac": frame_bytes * fps / world_size / 1e9 / peak},
        },
        "e2e": {"value": e2e_fps, "unit": "frames/s",
                "h2d_bytes_per_step": FRAMES_PER_STEP * world_size * __import__("ctypes").sizeof(N.FrameSetup),
                "d2h_bytes_per_step": FRAMES_PER_STEP * world_size * W * H * 4},
        "gpu_launches": main["launches"],
        "clocks": clocks,
    }
    if (1920, 1080) in results and (W, H) != (1920, 1080):
        r = results[(1920, 1080)]
        fr = FRAMES_PER_STEP * r["steps"] * world_size
        line["at_1080p"] = {"value": fr / (r["ms"] / 1000.0), "unit": "frames/s", "e2e": fr / (r["e2e_ms"] / 1000.0),
                            "phase1_kernel_ms_exclusive": r["x1"] / max(1, r["xn"]), "phase2_kernel_ms_exclusive": r["x2"] / max(1, r["xn"]),
                            "runs_per_s": sum(c["runs_visited"] for c in r["per"]) / len(r["per"]) * fr / (r["ms"] / 1000.0)}

    if world_size == 1 and not a.no_cpu_baseline:
        cpu = CpuPath(a.maxdim, W, H)   # the checker as the timed CPU arm (oracle/_ref), never on the product path
        t0 = time.perf_counter()
        done = 0
        for _ in range(8):
            for i in range(len(cpu.setups)):
                cpu.render(i)
                done += 1
            if time.perf_counter() - t0 > 12.0:
                break
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": done / dt, "unit": "frames/s", "cores": cpu.threads, "kind": cpu.kind,
                                "sample": f"{done} frames: whole passes over the same {FRAMES_PER_STEP}-pose path at {W}x{H}; " + cpu.describe()}
    print(json.dumps(line), flush=True)
    rm.destroy()
    if world_size > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world_size)
        return
    if world_size != a.gpus and world_size == 1 and a.gpus > 1:
        raise SystemExit(f"--gpus {a.gpus} needs torchrun: python -m torch.distributed.run --nnodes=1 --nproc-per-node {a.gpus} "
                         f"--master-addr 127.0.0.1 --master-port 29511 bench.py --gpus {a.gpus} ...")
    run_b200(a, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
