"""Kernel LOGIC against the oracle without a GPU: the Phase-1 CUDA source (cpuvox_b200/csrc/raybuffer_kernels.cu) compiled
for the CPU through the test-only SIMT emulator (tools/simt_emu: one fiber per lane, warp collectives as rendezvous) must give
bit-identical raybuffers and work counters. Both kernels are covered: the general one (element area, any world) and the
boundary-table one (regular worlds). This checks the algorithm, not the GPU build — `-m gpu` does that through the C ABI."""
from __future__ import annotations

import numpy as np
import pytest

import emu
from conftest import COMB_POSES, POSES, comb_world, irregular_world, setup_for

MAGENTA = 0xFF | (255 << 8) | (20 << 16) | (147 << 24)


def _check(cv, orc, ow, ew, s, W, H, what, variants=(0, 1), groups=(32,)):
    td0 = np.full((W + 2 * H, H), MAGENTA, dtype=np.uint32)
    lr0 = np.full((2 * W + H, W), MAGENTA, dtype=np.uint32)
    otd, olr, ocn = orc.render_raybuffers(ow, orc.copy_setup(s), W, H, td=td0, lr=lr0)
    for variant in variants:
        for group in groups:
            for counters in (True, False):
                td, lr, cn = emu.render_raybuffers(ew, s, W, H, variant=variant, group=group, counters=counters, fill=MAGENTA)
                tag = f"{what} kernel={'general' if variant == 0 else 'boundary-table'} g{group} counters={counters}"
                assert np.array_equal(td, otd), f"{tag}: top/down raybuffer differs in {int((td != otd).sum())} pixels"
                assert np.array_equal(lr, olr), f"{tag}: left/right raybuffer differs in {int((lr != olr).sum())} pixels"
                assert cn is None or cn == ocn, f"{tag}: counters {cn} != {ocn}"


@pytest.mark.parametrize("world_name,res,poses", [
    ("terrain_world", (160, 90), ("down60", "up16", "horizon", "outside", "roll37")),
    ("structure_world", (128, 200), ("down85", "pitch3", "roll180", "up80")),
    ("mill_world", (167, 109), ("down60", "up16", "horizon", "outside_far", "roll37")),
])
def test_emulated_kernels_match_oracle(cv, orc, request, world_name, res, poses):
    world = request.getfixturevalue(world_name)
    ow = orc.OracleWorld(world.dims, world.blobs, world.column_counts)
    ew = emu.EmuWorld(world)
    assert ew.regular, "worlds made by the voxelizer/synthetic builders are regular: the boundary-table kernel must apply"
    W, H = res
    for spec in POSES:
        if spec[0] in poses:
            _check(cv, orc, ow, ew, setup_for(cv, world, spec, W, H), W, H, f"{world_name} {spec[0]} {W}x{H}")


def test_emulated_narrow_groups(cv, orc, terrain_world):
    """8 and 16 lanes per ray, both kernels."""
    ow = orc.OracleWorld(terrain_world.dims, terrain_world.blobs, terrain_world.column_counts)
    ew = emu.EmuWorld(terrain_world)
    W, H = 96, 64
    for spec in (POSES[0], POSES[4]):
        _check(cv, orc, ow, ew, setup_for(cv, terrain_world, spec, W, H), W, H, spec[0], groups=(8, 16))


def test_emulated_tall_columns_and_near_plane(cv, orc):
    """Columns of 70..128 runs (several passes per column) and cameras inside geometry (runs that straddle the near plane take the
    per-run clipping path of the boundary-table kernel)."""
    world, blob, cc = comb_world(cv)
    ow = orc.OracleWorld(world.dims, [blob], [cc])
    ew = emu.EmuWorld(world)
    assert ew.regular
    lods = np.full(6, 1e9, dtype=np.float32)
    W, H = 160, 90
    for pos, eul in COMB_POSES:
        s = cv.frame_setup(cv.CameraPose.from_euler(pos, eul, far_clip=200.0), W, H, lods, world.dims[1])
        _check(cv, orc, ow, ew, s, W, H, f"comb {pos} {eul}")


def test_irregular_world_is_detected_and_general_kernel_matches(cv, orc):
    world, blob, cc = irregular_world(cv)
    ow = orc.OracleWorld(world.dims, [blob], [cc])
    ew = emu.EmuWorld(world)
    assert not ew.regular
    lods = np.full(6, 1e9, dtype=np.float32)
    W, H = 128, 96
    for pos, eul in [((16.5, 40.5, 2.5), (10, 0, 0)), ((16.5, 70.5, 16.5), (75, 30, 0)), ((3.5, 20.5, 3.5), (-30, 45, 0))]:
        s = cv.frame_setup(cv.CameraPose.from_euler(pos, eul, far_clip=100.0), W, H, lods, world.dims[1])
        _check(cv, orc, ow, ew, s, W, H, f"irregular {pos} {eul}", variants=(0,))


def test_emulated_kernels_fuzz(cv, orc):
    """Seeded fuzz of the kernel logic on the CPU: random small worlds and cameras, both kernels, against the oracle."""
    from conftest import random_world_and_cameras
    rng = np.random.default_rng(99)
    lods = np.full(6, 1e9, dtype=np.float32)
    for it in range(12):
        world, blob, cc, W, H, poses = random_world_and_cameras(cv, rng, cameras=2)
        ow = orc.OracleWorld(world.dims, [blob], [cc])
        ew = emu.EmuWorld(world)
        for k, pose in enumerate(poses):
            s = cv.frame_setup(pose, W, H, lods, world.dims[1])
            _check(cv, orc, ow, ew, s, W, H, f"fuzz world {it} {world.dims} {W}x{H} camera {k}", variants=(0, 1) if ew.regular else (0,))
