"""Host-side pieces either side of the hot path: world production (WordBuilder/World.DownSample restatement), the .world
file format, LOD distances, the benchmark camera path — checked against the oracle's independent restatement and against
the layout invariants of Assets/Code/World.cs."""
from __future__ import annotations

import os

import numpy as np
import pytest

from conftest import POSES, pose_for
from rle import column_count, decode_world, encode_world


def test_rle_encode_decode_round_trip():
    rng = np.random.default_rng(3)
    grid = np.zeros((8, 16, 4), dtype=np.uint32)
    mask = rng.random(grid.shape) < 0.3
    grid[mask] = rng.integers(1, 2**32 - 1, size=int(mask.sum()), dtype=np.uint64).astype(np.uint32)
    grid[:, :, 1] = 0        # empty columns
    grid[2, :, 2] = 0x11223344  # full column: one run, no air
    blob, cc = encode_world(grid)
    assert cc == column_count(8, 4)
    back = decode_world(blob, cc, grid.shape)
    assert np.array_equal(back, grid)


@pytest.mark.parametrize("world_name", ["terrain_world", "structure_world", "mill_world"])
def test_builder_blobs_follow_the_reference_layout(request, world_name):
    """Every LOD blob decodes (guards, full-height runs, worldMin/Max in world units), its solid count equals the voxel
    count the reference logs per LOD (UnityManager.cs:326-331), and LOD j occupancy is the 2^j-cube OR of LOD 0
    (World.DownSample, World.cs:45-127)."""
    world = request.getfixturevalue(world_name)
    dims = world.dims
    lod0 = decode_world(world.blobs[0], world.column_counts[0], dims, 0)
    assert int((lod0 != 0).sum()) == world.voxel_counts[0]
    for lod in range(1, min(4, len(world.blobs))):
        assert world.column_counts[lod] == column_count(dims[0], dims[2], lod)
        g = decode_world(world.blobs[lod], world.column_counts[lod], dims, lod)
        assert int((g != 0).sum()) == world.voxel_counts[lod]
        s = 1 << lod
        pooled = (lod0 != 0).reshape(dims[0] // s, s, dims[1] // s, s, dims[2] // s, s).any(axis=(1, 3, 5))
        assert np.array_equal(pooled, g != 0), f"LOD {lod} occupancy"


def test_mesh_voxelizer_single_triangle(cv):
    """One axis-aligned triangle in the plane y = const voxelizes to a single-voxel-thick sheet with the vertex colour."""
    pos = np.array([[0, 0, 0], [0, 0, 8], [8, 0, 0],      # ground triangle
                    [0, 0, 0], [0, 8, 0], [0.001, 8, 0.001]], dtype=np.float32)  # sliver giving the mesh a height
    col = np.tile(np.array([[200, 100, 50, 255]], dtype=np.uint8), (6, 1))
    w = cv.World.from_mesh(pos, col, 16, lods=1)
    assert w.dims[0] == 16 and w.dims[2] == 16
    g = decode_world(w.blobs[0], w.column_counts[0], w.dims)
    sheet = g[:, 0, :] != 0
    assert sheet.sum() > 60 and sheet[1, 1] and not sheet[15, 15]
    v = int(g[2, 0, 2])
    assert (v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF, v >> 24) == (255, 200, 100, 50)  # bytes a, r, g, b


def test_world_file_round_trip(cv, terrain_world, tmp_path):
    """WorldSaveFile.Serialize/Deserialize (WorldSaveFile.cs:8-94): 24-byte header, (offset, length) table, raw blobs."""
    p = str(tmp_path / "t.world")
    terrain_world.save(p)
    raw = open(p, "rb").read()
    hdr = np.frombuffer(raw[:24], dtype=np.int32)
    assert list(hdr[2:5]) == list(terrain_world.dims) and hdr[5] == len(terrain_world.blobs)
    back = cv.World.load(p)
    assert back.dims == terrain_world.dims and len(back.blobs) == len(terrain_world.blobs)
    for a, b, ca, cb in zip(back.blobs, terrain_world.blobs, back.column_counts, terrain_world.column_counts):
        assert np.array_equal(a, b) and ca == cb


def test_setup_lods_matches_oracle_and_expected_magnitudes(cv, orc):
    """UnityManager.SetupLods (UnityManager.cs:417-458): LOD j starts where one pixel spans 1.41/lodError * (2<<j) world units."""
    for (dim, W, H) in [(1024, 1920, 1080), (1024, 3840, 2160), (2048, 1280, 720), (4096, 7680, 4320)]:
        a = cv.setup_lods(dim, W, H)
        b = orc.setup_lods(dim, W, H)
        assert np.array_equal(a, b)
        assert (np.diff(a) >= 0).all() and a[-1] == np.ceil(2.0 * 2 * dim)
    a = cv.setup_lods(1024, 1920, 1080)
    assert 1100 < a[0] < 1250  # ~1176 for FOV 85, lodError 1
    assert np.array_equal(cv.setup_lods(1024, 1920, 1080, lod_error=2.0) <= a, np.ones(6, dtype=bool))


def test_benchmark_path_matches_oracle_and_key_frames(cv, orc):
    """BenchmarkPath.anim sampled with cubic Hermite; position keys scale with the world dimensions (UnityManager.cs:86-87)."""
    dims = (1024, 1024, 1024)
    length = cv.benchmark_length()
    assert abs(length - 1.15) < 1e-6
    for i in range(60):
        t = length * i / 59
        p = cv.benchmark_pose(t, dims)
        op, oq = orc.benchmark_pose(t, dims)
        assert tuple(np.float32(p.position)) == tuple(np.float32(op)) and tuple(np.float32(p.rotation)) == tuple(np.float32(oq))
    first, last = cv.benchmark_pose(0.0, dims), cv.benchmark_pose(length, dims)
    np.testing.assert_allclose(first.position, (-102.4, 512.0, -102.4), rtol=1e-5)   # starts outside the world
    np.testing.assert_allclose(last.position, (0.427 * 1024, 0.95 * 1024, 0.52 * 1024), rtol=1e-5)
    q = orc.quat_euler(85.0, -225.5, 360.0)
    np.testing.assert_allclose(np.abs(last.rotation), np.abs(q), atol=1e-6)


def test_limit_rotation_horizon(cv):
    """|forward.y| < 0.001 is pushed to +-0.001 (UnityManager.cs:193-201) so the vanishing point stays finite."""
    W, H = 640, 360
    lods = cv.setup_lods(256, W, H)
    s = cv.frame_setup(cv.CameraPose.from_euler((10, 10, 10), (0.0, 45.0, 0.0)), W, H, lods, 256)
    assert np.isfinite(s.vanishing_point_screen[1]) and abs(s.vanishing_point_screen[1]) > 1e4
    assert s.camera.inverse_element_iteration_direction == 1  # Mathf.Sign(0) = +1 -> looks (just) up
    total = sum(max(0, s.segments[k].ray_count) for k in range(4))
    assert total == W  # one clamped segment spanning the screen width


def test_write_bmp_round_trip(cv, tmp_path):
    """cvx_host_write_bmp: 54-byte header, bottom-up BGRA rows = our row order with every pixel byte-reversed."""
    rng = np.random.default_rng(3)
    W, H = 37, 11
    frame = rng.integers(0, 2**32, size=(H, W), dtype=np.uint64).astype(np.uint32)
    path = str(tmp_path / "f.bmp")
    cv.write_bmp(path, frame)
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[:2].tobytes() == b"BM" and raw.size == 54 + W * H * 4
    assert int(raw[2:6].view(np.uint32)[0]) == raw.size and int(raw[10:14].view(np.uint32)[0]) == 54
    assert int(raw[18:22].view(np.int32)[0]) == W and int(raw[22:26].view(np.int32)[0]) == H and int(raw[28:30].view(np.uint16)[0]) == 32
    px = raw[54:].reshape(H, W, 4)                   # b, g, r, a
    ours = frame.view(np.uint8).reshape(H, W, 4)     # a, r, g, b
    assert np.array_equal(px, ours[..., ::-1])


def test_oracle_debug_view_hand_case(orc):
    """COPY_MAIN1/2 restatement (RayBufferBlit.shader:48-53): screen x picks the ray row, screen y (bottom-up) the pixel along it."""
    rows, row_len, W, H = 6, 4, 3, 2
    buf = (np.arange(rows * row_len, dtype=np.uint32) + 1).reshape(rows, row_len)
    fr = orc.blit_raybuffer(buf, W, H)
    for y in range(H):
        for x in range(W):
            u = 1.0 - (H - (y + 0.5)) / H      # = (y + 0.5) / H
            v = (x + 0.5) / W
            assert fr[y, x] == buf[int(v * rows), int(u * row_len)]
    # identity-sized view: a raybuffer with as many rows as screen columns and row_len == H shows up transposed
    buf = np.arange(5 * 7, dtype=np.uint32).reshape(5, 7)
    assert np.array_equal(orc.blit_raybuffer(buf, 5, 7), buf.T)


def test_algorithmic_bytes_formula(cv):
    c = {"dda_steps": 10, "runs_visited": 20, "px_voxel": 30, "px_sky": 40}
    assert cv.algorithmic_bytes(c, 8, 4) == 12 * 10 + 4 * 20 + 4 * 30 + 4 * 70 + 8 * 32
