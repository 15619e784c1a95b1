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
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cpuvox_b200", "csrc"), "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
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


def setup_for(cv, world, spec, W, H):
    lods = cv.setup_lods(world.max_dimension, W, H)
    return cv.frame_setup(pose_for(cv, world, spec), W, H, lods, world.dims[1])
