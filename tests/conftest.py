"""Shared fixtures. `-m "not gpu"` runs here on CPU (oracle, host logic, ABI surface); `-m gpu` are the parity tests
proper and call the CUDA path through the C ABI on a B200."""
from __future__ import annotations

import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build the product library and the oracle once per session (nvcc cross-compiles without a GPU)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cpuvox_b200", "csrc"), "-s", "-j4"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


MILL = os.path.join(ROOT, "tests", "data", "mill.obj")

# (name, euler pitch/yaw/roll degrees, position as a fraction of the world dimensions) — the cases the reference's
# benchmark path and README exercise: looking down (4 segments), up (inverted run order), along the horizon
# (LimitRotationHorizon, vanishing point far off screen, clamped segments), from outside the world, rolled.
POSES = [
    ("down60", (60.0, 30.0, 0.0), (0.5, 0.9, 0.5)),
    ("down85", (85.0, -135.0, 0.0), (0.43, 0.95, 0.52)),
    ("up16", (-16.2, -135.0, 0.0), (0.9, 0.3, 0.9)),
    ("horizon", (0.0, 45.0, 0.0), (0.5, 0.5, 0.5)),
    ("pitch3", (3.0, 200.0, 0.0), (0.3, 0.6, 0.7)),
    ("outside", (0.0, 45.0, 0.0), (-0.1, 0.5, -0.1)),
    ("outside_far", (10.0, 225.0, 0.0), (-0.5, 0.7, -0.5)),
    ("roll180", (59.12, -135.0, 180.0), (0.9, 0.95, 0.9)),
    ("roll37", (40.0, 10.0, 37.0), (0.2, 0.8, 0.4)),
    ("up80", (-80.0, 0.0, 0.0), (0.5, 0.1, 0.5)),
]


@pytest.fixture(scope="session")
def cv():
    import cpuvox_b200
    return cpuvox_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref: the reference's own C# sources translated to C++ and compiled (oracle/ref.py). Built here from
    /root/reference; on the GPU box the prebuilt .so travels with the snapshot."""
    from oracle import ref as r
    if r.build() is None:
        pytest.skip("oracle/_ref/libcpuvox_ref.so absent and /root/reference not available to build it")
    r.lib()
    return r


@pytest.fixture(scope="session")
def terrain_world(cv):
    """Seeded fBm heightmap shell (BASELINE config 2 generator at test size)."""
    return cv.World.synthetic(0, (256, 256, 256), seed=1234)


@pytest.fixture(scope="session")
def structure_world(cv):
    """Seeded boxes/pipes/slabs world with many multi-run columns (BASELINE config 4 generator at test size), X != Z."""
    return cv.World.synthetic(1, (512, 128, 256), seed=7)


@pytest.fixture(scope="session")
def mill_world(cv):
    """datasets/mill.obj through the voxelizer restatement at maxDimension 256."""
    return cv.World.from_obj(MILL, 256)


def pose_for(cv, world, spec, far_scale=2.0):
    _, euler, frac = spec
    pos = tuple(frac[i] * world.dims[i] for i in range(3))
    return cv.CameraPose.from_euler(pos, euler, far_clip=far_scale * world.max_dimension)


def limited(cv, pose):
    """The pose after UnityManager.LimitRotationHorizon (UnityManager.cs:193-201), as LateUpdate hands it to DrawWorld."""
    import ctypes as C
    from cpuvox_b200.native import lib
    p = pose.to_native(16, 16)
    lib.cvx_host_limit_rotation_horizon(C.byref(p))
    return cv.CameraPose(tuple(p.position), tuple(p.rotation), pose.fov_y_degrees, pose.near_clip, pose.far_clip)


def parse_obj(path):
    """ObjModel.Import restated by the product's host code: triangle soup positions (n, 3) float32 and Color32 (n, 4) uint8."""
    import ctypes as C
    from cpuvox_b200.native import lib
    pos, col, n = C.c_void_p(), C.c_void_p(), C.c_int32()
    assert lib.cvx_obj_parse(path.encode(), 0, C.byref(pos), C.byref(col), C.byref(n)) == 0
    try:
        P = np.ctypeslib.as_array(C.cast(pos, C.POINTER(C.c_float)), (n.value, 3)).copy()
        Cc = np.ctypeslib.as_array(C.cast(col, C.POINTER(C.c_uint8)), (n.value, 4)).copy()
    finally:
        lib.cvx_host_free(pos)
        lib.cvx_host_free(col)
    return P, Cc


def setup_for(cv, world, spec, W, H):
    lods = cv.setup_lods(world.max_dimension, W, H)
    return cv.frame_setup(pose_for(cv, world, spec), W, H, lods, world.dims[1])


def comb_world(cv):
    """Hand-built 64x256x64 world (tests/rle.py encoder): comb columns of 128 one-voxel runs and 8x8 columns of ~70 random voxels
    (tall columns: several passes of the Phase-1 round cache), plus scattered short runs. Returns (World, blob, column_count)."""
    from rle import encode_world
    rng = np.random.default_rng(5)
    dims = (64, 256, 64)
    grid = np.zeros(dims, dtype=np.uint32)
    for (x, z) in [(20, 20), (21, 20), (40, 33), (10, 50), (33, 34)]:
        grid[x, ::2, z] = rng.integers(1, 2**32 - 1, size=128, dtype=np.uint64).astype(np.uint32) | 0xFF
    for _ in range(300):
        x, z = rng.integers(0, 64, 2)
        y0 = rng.integers(0, 250)
        h = rng.integers(1, 6)
        grid[x, y0:y0 + h, z] = rng.integers(1, 2**32 - 1, dtype=np.uint64).astype(np.uint32) | 0xFF
    for x in range(28, 36):
        for z in range(28, 36):
            ys = rng.integers(0, 256, size=70)
            grid[x, ys, z] = rng.integers(1, 2**32 - 1, size=70, dtype=np.uint64).astype(np.uint32) | 0xFF
    blob, cc = encode_world(grid)
    return cv.World(dims, [blob], [cc], [int((grid != 0).sum())]), blob, cc


# camera (position, euler) cases for comb_world: from the side, from above, inside the tall block (near-plane clipping of
# runs that straddle the camera plane), looking up, rolled, high above looking straight down
COMB_POSES = [((32.5, 128.5, 2.5), (0, 0, 0)), ((32.5, 200.5, 32.5), (80, 10, 0)), ((31.5, 100.3, 31.5), (30, 45, 0)),
              ((32.2, 40.5, 30.5), (-50, 200, 0)), ((5.5, 250.5, 5.5), (45, 45, 20)), ((32.5, 128.5, 32.5), (5, 90, 0)),
              ((32.5, 300.5, 32.5), (89, 0, 0))]


def irregular_world(cv):
    """A world whose columns are NOT all full-height runs of valid elements (a zero-length element inside one column, a
    column that stops short of the floor): the upload must classify it as irregular and Phase 1 must take the general kernel."""
    from rle import encode_world
    rng = np.random.default_rng(11)
    dims = (32, 64, 32)
    grid = np.zeros(dims, dtype=np.uint32)
    for _ in range(400):
        x, z = rng.integers(0, 32, 2)
        y0 = rng.integers(0, 60)
        grid[x, y0:y0 + rng.integers(1, 5), z] = rng.integers(1, 2**32 - 1, dtype=np.uint64).astype(np.uint32) | 0xFF
    blob, cc = encode_world(grid)
    words = blob.view(np.uint32)
    cells = words[3 * cc:]
    hdr = words[:3 * cc].reshape(cc, 3)
    multi = [i for i in range(32 * 32) if (hdr[i, 1] & 0xFFFF) >= 4]
    a, b = multi[0], multi[len(multi) // 2]
    cells[hdr[a, 0] + 3] = cells[hdr[a, 0] + 3] & 0xFFFF          # run 2 of column a: Length 0 -> the reference stops there
    last = hdr[b, 0] + (hdr[b, 1] & 0xFFFF)
    e = int(cells[last])
    cells[last] = (e & 0xFFFF) | ((((e >> 16) & 0xFFFF) + 3) << 16)  # column b: lengths no longer add up to the height
    return cv.World(dims, [blob], [cc], [int((grid != 0).sum())]), blob, cc


def random_world_and_cameras(cv, rng, cameras=4):
    """Fuzz case: a small random world (sparse voxels / heightmap / slabs with gaps; power-of-two dims 8..32, tests/rle.py encoder) and
    random cameras inside and outside it (any pitch short of vertical, any yaw, sometimes rolled), a random resolution and far clip.
    Returns (World, blob, column_count, W, H, [CameraPose])."""
    from rle import encode_world
    dx, dy, dz = (int(2 ** rng.integers(3, 6)) for _ in range(3))
    grid = np.zeros((dx, dy, dz), dtype=np.uint32)
    style = int(rng.integers(0, 3))
    if style == 0:
        m = rng.random((dx, dy, dz)) < rng.uniform(0.01, 0.3)
        grid[m] = rng.integers(1, 2 ** 32 - 1, size=int(m.sum()), dtype=np.uint64).astype(np.uint32) | 0xFF
    elif style == 1:
        h = rng.integers(1, dy, size=(dx, dz))
        for x in range(dx):
            for z in range(dz):
                grid[x, :h[x, z], z] = rng.integers(1, 2 ** 32 - 1, dtype=np.uint64).astype(np.uint32) | 0xFF
    else:
        for _ in range(int(rng.integers(1, 6))):
            y0 = int(rng.integers(0, dy))
            grid[:, y0:y0 + int(rng.integers(1, 4)), :] = rng.integers(1, 2 ** 32 - 1, dtype=np.uint64).astype(np.uint32) | 0xFF
            x0 = int(rng.integers(0, dx))
            grid[x0:x0 + 2, :, :] = 0
    blob, cc = encode_world(grid)
    world = cv.World((dx, dy, dz), [blob], [cc], [int((grid != 0).sum())])
    W, H = int(rng.integers(16, 80)), int(rng.integers(16, 80))
    poses = []
    for _ in range(cameras):
        inside = rng.random() < 0.6
        pos = tuple(float(rng.uniform(0, d)) if inside else float(rng.uniform(-0.5 * d, 1.5 * d)) for d in (dx, dy, dz))
        euler = (float(rng.uniform(-89.5, 89.5)), float(rng.uniform(0, 360)), float(rng.choice([0.0, 0.0, rng.uniform(-180, 180)])))
        poses.append(cv.CameraPose.from_euler(pos, euler, far_clip=float(rng.uniform(10, 4 * max(dx, dy, dz)))))
    return world, blob, cc, W, H, poses


def partition_rays_even(total: int, ranks: int, rank: int):
    """Contiguous, equally sized ray ranges (test helper for the sharded draws)."""
    return total * rank // ranks, total * (rank + 1) // ranks
