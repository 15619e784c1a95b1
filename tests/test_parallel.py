"""Multi-GPU host logic on CPU: world_size-2 `gloo` process groups exercise the ray/view partition, the world broadcast and
the disjoint-framebuffer gather algebra (the oracle stands in for the per-rank renderer)."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

from conftest import POSES, ROOT, setup_for


def test_partition_rays_covers_everything_once(cv):
    for total in (0, 1, 7, 6000, 12001):
        for n in (1, 2, 3, 8):
            parts = cv.partition_rays(total, n)
            assert len(parts) == n and parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
            assert max(e - b for b, e in parts) - min(e - b for b, e in parts) <= 1
    w = np.concatenate([np.full(1000, 1.0), np.full(1000, 3.0)])
    (b0, e0), (b1, e1) = cv.partition_rays(2000, 2, w)
    assert e0 == b1 and abs(w[b0:e0].sum() - w[b1:e1].sum()) <= 3.0 and e0 > 1000
    with pytest.raises(ValueError):
        cv.partition_rays(10, 2, [1.0] * 9)


def test_partition_views_round_robin(cv):
    views = [cv.partition_views(10, 4, r) for r in range(4)]
    assert sorted(sum(views, [])) == list(range(10)) and views[1] == [1, 5, 9]


def test_ray_weights_follow_segment_pixel_ranges(cv, terrain_world):
    W, H = 320, 180
    s = setup_for(cv, terrain_world, POSES[0], W, H)
    w = cv.ray_weights(s, W, H)
    assert len(w) == sum(max(0, s.segments[k].ray_count) for k in range(4))
    vy = int(np.rint(np.float32(s.vanishing_point_screen[1])))
    assert w[0] == H - vy and w[s.segments[0].ray_count] == vy + 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import cpuvox_b200 as cv
    from conftest import POSES, setup_for
    from oracle import oracle as orc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    try:
        # world built on rank 0 only, broadcast once, replicated
        world = cv.World.synthetic(0, (128, 128, 128), seed=5, lods=3) if rank == 0 else None
        world = cv.broadcast_world(world, src=0)
        W, H = 200, 120
        ow = orc.OracleWorld(world.dims, world.blobs, world.column_counts)
        s = setup_for(cv, world, POSES[0], W, H)
        total = sum(max(0, s.segments[k].ray_count) for k in range(4))
        begin, end = cv.partition_rays(total, world_size, cv.ray_weights(s, W, H))[rank]
        # this rank's rays only; rows of other ranks stay zero, so the blit picks up zeros for pixels it does not own
        td, lr, cn = orc.render_raybuffers(ow, orc.copy_setup(s), W, H, ray_begin=begin, ray_end=end)
        frame = orc.blit(orc.copy_setup(s), W, H, td, lr)
        t = torch.from_numpy(frame.view(np.int32).copy())
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)  # disjoint pixels: sum == copy
        counters = torch.tensor([cn[k] for k in sorted(cn)], dtype=torch.int64)
        dist.reduce(counters, dst=0, op=dist.ReduceOp.SUM)
        # batched views: round-robin
        mine = cv.partition_views(5, world_size, rank)
        crcs = torch.zeros(5, dtype=torch.int64)
        for i in mine:
            si = setup_for(cv, world, POSES[i], W, H)
            a, b, _ = orc.render_raybuffers(ow, orc.copy_setup(si), W, H)
            crcs[i] = int(orc.blit(orc.copy_setup(si), W, H, a, b).astype(np.uint64).sum())
        dist.all_reduce(crcs, op=dist.ReduceOp.SUM)
        if rank == 0:
            np.save(os.path.join(out_dir, "frame.npy"), t.numpy().view(np.uint32))
            np.save(os.path.join(out_dir, "counters.npy"), counters.numpy())
            np.save(os.path.join(out_dir, "views.npy"), crcs.numpy())
            np.save(os.path.join(out_dir, "blob0.npy"), world.blobs[0])
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_rank(cv, orc, tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    world = cv.World.synthetic(0, (128, 128, 128), seed=5, lods=3)
    assert np.array_equal(np.load(tmp_path / "blob0.npy"), world.blobs[0])
    W, H = 200, 120
    ow = orc.OracleWorld(world.dims, world.blobs, world.column_counts)
    s = setup_for(cv, world, POSES[0], W, H)
    td, lr, cn = orc.render_raybuffers(ow, orc.copy_setup(s), W, H)
    frame = orc.blit(orc.copy_setup(s), W, H, td, lr)
    assert np.array_equal(np.load(tmp_path / "frame.npy"), frame), "n-rank result must equal the 1-rank result bit for bit"
    assert list(np.load(tmp_path / "counters.npy")) == [cn[k] for k in sorted(cn)]
    sums = []
    for i in range(5):
        si = setup_for(cv, world, POSES[i], W, H)
        a, b, _ = orc.render_raybuffers(ow, orc.copy_setup(si), W, H)
        sums.append(int(orc.blit(orc.copy_setup(si), W, H, a, b).astype(np.uint64).sum()))
    assert list(np.load(tmp_path / "views.npy")) == sums
