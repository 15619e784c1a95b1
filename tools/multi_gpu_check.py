"""Developer tool (GPU box, N >= 2 GPUs, run under torchrun): the multi-GPU paths of cpuvox_b200/parallel.py on real devices.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
Checks, bit for bit against a single-GPU render on rank 0: rays sharded with the CUDA-IPC peer-store gather ("p2p"), rays sharded
with the ncclReduce gather ("reduce"), and views sharded (draw_views). Prints per-mode frame times at 4K and 8K."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import cpuvox_b200 as cv

    rank, local, n = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    world = cv.World.from_obj(os.path.join(ROOT, "tests", "data", "mill.obj"), 512) if rank == 0 else None
    world = cv.broadcast_world(world, src=0, device=torch.device(f"cuda:{local}"))
    poses = cv.benchmark_path(world.dims, 12, far_clip=2.0 * world.max_dimension)
    ok = True
    ref = None
    if rank == 0:
        ref = cv.RenderManager(local)
        ref.upload_world(world)
    for gather in ("p2p", "reduce"):
        srm = cv.ShardedRenderManager(local, rank, n, gather=gather)
        srm.upload_world(world)
        for (W, H) in ((640, 360), (3840, 2160), (7680, 4320)):
            srm.set_resolution(W, H)
            if rank == 0:
                ref.set_resolution(W, H)
            times = []
            for i, pose in enumerate(poses if W < 3000 else poses[::4]):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                srm.draw_world_sharded(pose)
                if gather == "reduce":
                    torch.cuda.current_stream().synchronize()
                    dist.barrier()
                times.append(time.perf_counter() - t0)
                if rank == 0:
                    got = srm.read_frame()
                    ref.draw_world(pose)
                    want = ref.read_frame()
                    same = np.array_equal(got, want)
                    ok &= same
                    if not same:
                        print(f"MISMATCH gather={gather} {W}x{H} pose {i}: {int((got != want).sum())} pixels", flush=True)
            if rank == 0:
                print(f"rays sharded over {n} GPUs, gather={gather}, {W}x{H}: {1000 * np.median(times):.3f} ms/frame (host-timed, incl. barrier)", flush=True)
        srm.destroy()
    # views sharded
    srm = cv.ShardedRenderManager(local, rank, n)
    srm.upload_world(world)
    W, H = 1280, 720
    srm.set_resolution(W, H)
    mine = cv.partition_views(len(poses), n, rank)
    dst = cv.alloc_pinned((len(mine), H, W))
    got_idx = srm.draw_views(poses, dst)
    frames = [None] * n
    dist.all_gather_object(frames, (got_idx, np.array(dst)))
    if rank == 0:
        ref.set_resolution(W, H)
        for idx, fr in frames:
            for j, i in enumerate(idx):
                ref.draw_world(poses[i])
                same = np.array_equal(ref.read_frame(), fr[j])
                ok &= same
                if not same:
                    print(f"MISMATCH views sharded: view {i}", flush=True)
        print("views sharded: ok" if ok else "views sharded: FAILED", flush=True)
        print("MULTI-GPU CHECK " + ("PASSED" if ok else "FAILED"), flush=True)
        ref.destroy()
    srm.destroy()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
