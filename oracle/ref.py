"""TEST INFRASTRUCTURE ONLY — build recipe + ctypes wrapper of oracle/_ref/libcpuvox_ref.so: the REFERENCE'S OWN C# sources
(read from /root/reference where they lie) translated to C++ by oracle/refbuild/cs2cpp.py and compiled with g++ against
oracle/refbuild/unity_shim.hpp (our stand-in for the Unity packages the reference calls but does not vendor).

Why this way: the image has no C# toolchain (dotnet / mono / mcs / csc absent here and on the GPU box, probed), so the .NET
build north_star describes (oracle/dotnet/, shipped but never compiled) cannot run. This build CAN: it pins the hand-written
oracle (oracle/cpuvox_oracle.cpp) and the CUDA path to the reference's text instead of to our reading of it.

Only tests/, __graft_entry__ (build + smoke) and bench.py's CPU legs may import this. Nothing is copied from the reference:
generated C++ goes to a temporary directory and only the .so lands in oracle/_ref/ (git-ignored, travels with gpurun).
/root/reference does not exist on the GPU box — there the prebuilt .so is used as is.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CPUVOX_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(_HERE, "_ref")
LIB_PATH = os.path.join(OUT_DIR, "libcpuvox_ref.so")
RECIPE = [os.path.join(_HERE, "refbuild", f) for f in ("cs2cpp.py", "unity_shim.hpp", "ref_prelude.hpp", "ref_driver.cpp")]
LODS = 6
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-w"]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build(force: bool = False) -> str | None:
    """(Re)build oracle/_ref/libcpuvox_ref.so when the reference tree is present; otherwise keep what is there."""
    if not os.path.isdir(os.path.join(REF_SRC, "Assets", "Code")):
        return LIB_PATH if available() else None
    newest = max(os.path.getmtime(p) for p in RECIPE)
    if not force and available() and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    gen = os.environ.get("CPUVOX_REF_KEEP_GEN") or tempfile.mkdtemp(prefix="cpuvox_refgen_")
    try:
        os.makedirs(gen, exist_ok=True)
        subprocess.check_call([sys.executable, RECIPE[0], "--ref", REF_SRC, "-o", os.path.join(gen, "ref_gen.hpp")])
        subprocess.check_call(["g++", *CXXFLAGS, "-shared", "-I", os.path.join(_HERE, "refbuild"), "-I", gen,
                               "-o", LIB_PATH + ".tmp", RECIPE[3]])
        os.replace(LIB_PATH + ".tmp", LIB_PATH)
    finally:
        if not os.environ.get("CPUVOX_REF_KEEP_GEN"):
            shutil.rmtree(gen, ignore_errors=True)
    return LIB_PATH


class Segment(C.Structure):
    _fields_ = [("min_screen", C.c_float * 2), ("max_screen", C.c_float * 2),
                ("cam_local_plane_ray_min", C.c_float * 2), ("cam_local_plane_ray_max", C.c_float * 2),
                ("ray_count", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("world_to_screen", C.c_float * 16), ("position_xz", C.c_float * 2), ("position_y", C.c_float),
                ("inverse_element_iteration_direction", C.c_int32), ("far_clip", C.c_float),
                ("lod_distances", C.c_float * LODS)]


class FrameSetup(C.Structure):
    _fields_ = [("segments", Segment * 4), ("camera", Camera), ("vanishing_point_screen", C.c_float * 2)]


class Pose(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 4), ("fov_y_degrees", C.c_float),
                ("near_clip", C.c_float), ("far_clip", C.c_float), ("pixel_width", C.c_int32), ("pixel_height", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libcpuvox_ref.so is absent and /root/reference is not here to build it from")
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_describe.restype = C.c_char_p
        L.ref_world_create.restype = P
        L.ref_world_create.argtypes = [C.c_int32] * 3
        L.ref_world_set_lod.argtypes = [P, C.c_int32, P, C.c_int64, C.c_int32]
        L.ref_world_column_count.argtypes = [C.c_int32] * 4
        L.ref_world_free.argtypes = [P]
        L.ref_frame_setup_from_pose.argtypes = [C.POINTER(Pose), C.POINTER(C.c_float * LODS), C.c_int32, C.POINTER(FrameSetup)]
        L.ref_render_raybuffers.argtypes = [P, C.POINTER(FrameSetup), C.c_int32, C.c_int32, P, P, C.c_int32]
        L.ref_blit.argtypes = [C.POINTER(FrameSetup), C.c_int32, C.c_int32, P, P, P]
        L.ref_draw_world.argtypes = [P, C.POINTER(Pose), C.POINTER(C.c_float * LODS), P, P, P, C.c_int32]
        L.ref_dda_walk.argtypes = [C.POINTER(C.c_float * 2), C.POINTER(C.c_float * 2), C.POINTER(C.c_float * LODS), C.c_float, C.c_int32, P, P]
        L.ref_build_world_from_mesh.restype = P
        L.ref_build_world_from_mesh.argtypes = [P, P, C.c_int32, P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.ref_built_world_info.argtypes = [P, P, P, P, P]
        L.ref_built_world_copy_blob.argtypes = [P, C.c_int32, P, C.c_int64]
        L.ref_built_world_free.argtypes = [P]
        L.ref_built_world_save.argtypes = [P, C.c_char_p]
        L.ref_world_load.restype = P
        L.ref_world_load.argtypes = [C.c_char_p, P, P]
        L.ref_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().ref_last_error().decode()))


def describe() -> str:
    return lib().ref_describe().decode()


def _pad_rows(rows: int) -> int:
    return (rows + 255) // 256 * 256


class RefWorld:
    """World LODs handed to the reference's `new World(dimensions, lod, data)`; blobs are borrowed numpy uint8 arrays."""

    def __init__(self, dims, blobs, column_counts):
        L = lib()
        self.dims = tuple(int(d) for d in dims)
        self._blobs = [np.ascontiguousarray(b) for b in blobs]
        self._w = C.c_void_p(L.ref_world_create(*self.dims))
        for lod, (b, cc) in enumerate(zip(self._blobs, column_counts)):
            _check(L.ref_world_set_lod(self._w, lod, b.ctypes.data_as(C.c_void_p), b.nbytes, int(cc)), "ref_world_set_lod")

    def __del__(self):
        if getattr(self, "_w", None):
            lib().ref_world_free(self._w)
            self._w = None


def copy_setup(src) -> FrameSetup:
    out = FrameSetup()
    assert C.sizeof(out) == C.sizeof(src)
    C.memmove(C.byref(out), C.byref(src), C.sizeof(out))
    return out


def make_pose(position, rotation, width, height, fov=85.0, near=0.05, far=2048.0) -> Pose:
    p = Pose()
    p.position[:] = position
    p.rotation[:] = rotation
    p.fov_y_degrees, p.near_clip, p.far_clip, p.pixel_width, p.pixel_height = fov, near, far, width, height
    return p


def frame_setup(position, rotation, width, height, lod_distances, world_dim_y, fov=85.0, near=0.05, far=2048.0) -> FrameSetup:
    """a1-a5 through the reference's own functions. The rotation must already have LimitRotationHorizon applied."""
    p = make_pose(position, rotation, width, height, fov, near, far)
    lods = (C.c_float * LODS)(*[float(x) for x in lod_distances])
    out = FrameSetup()
    _check(lib().ref_frame_setup_from_pose(C.byref(p), C.byref(lods), world_dim_y, C.byref(out)), "ref_frame_setup_from_pose")
    return out


def alloc_raybuffers(width, height):
    """Zeroed flat raybuffers padded to whole 256-ray partial textures (RayBuffer.cs:18-19), reusable across frames."""
    W, H = width, height
    return np.zeros((_pad_rows(W + 2 * H), H), dtype=np.uint32), np.zeros((_pad_rows(2 * W + H), W), dtype=np.uint32)


def render_raybuffers(world: RefWorld, setup: FrameSetup, width, height, threads=0, buffers=None):
    W, H = width, height
    td, lr = buffers if buffers is not None else alloc_raybuffers(W, H)
    assert td.shape == (_pad_rows(W + 2 * H), H) and lr.shape == (_pad_rows(2 * W + H), W)
    _check(lib().ref_render_raybuffers(world._w, C.byref(setup), W, H, td.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p), threads),
           "ref_render_raybuffers")
    return td[:W + 2 * H], lr[:2 * W + H]


def blit(setup: FrameSetup, width, height, td, lr):
    frame = np.zeros((height, width), dtype=np.uint32)
    td = np.ascontiguousarray(td[:width + 2 * height], dtype=np.uint32)
    lr = np.ascontiguousarray(lr[:2 * width + height], dtype=np.uint32)
    _check(lib().ref_blit(C.byref(setup), width, height, td.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p),
                          frame.ctypes.data_as(C.c_void_p)), "ref_blit")
    return frame


def draw_world(world: RefWorld, position, rotation, width, height, lod_distances, fov=85.0, near=0.05, far=2048.0, threads=0):
    """RenderManager.DrawWorld end to end: (td, lr, frame)."""
    W, H = width, height
    p = make_pose(position, rotation, W, H, fov, near, far)
    lods = (C.c_float * LODS)(*[float(x) for x in lod_distances])
    td = np.zeros((W + 2 * H, H), dtype=np.uint32)
    lr = np.zeros((2 * W + H, W), dtype=np.uint32)
    frame = np.zeros((H, W), dtype=np.uint32)
    _check(lib().ref_draw_world(world._w, C.byref(p), C.byref(lods), td.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p),
                                frame.ctypes.data_as(C.c_void_p), threads), "ref_draw_world")
    return td, lr, frame


def dda_walk(start, direction, lod_distances, far_clip, max_steps=100000):
    cells = np.zeros((max_steps, 3), dtype=np.int32)
    dists = np.zeros((max_steps, 2), dtype=np.float32)
    n = lib().ref_dda_walk(C.byref((C.c_float * 2)(*start)), C.byref((C.c_float * 2)(*direction)),
                           C.byref((C.c_float * LODS)(*[float(x) for x in lod_distances])), far_clip, max_steps,
                           cells.ctypes.data_as(C.c_void_p), dists.ctypes.data_as(C.c_void_p))
    return cells[:n], dists[:n]


def load_world_file(path) -> "RefWorld":
    """WorldSaveFile.Deserialize (WorldSaveFile.cs:57-93) by the reference's own code."""
    L = lib()
    dims = (C.c_int32 * 3)()
    n = C.c_int32()
    h = L.ref_world_load(str(path).encode(), dims, C.byref(n))
    if not h:
        raise RuntimeError("ref_world_load: " + L.ref_last_error().decode())
    w = RefWorld.__new__(RefWorld)
    w.dims, w._blobs, w._w, w.world_count = tuple(dims), [], C.c_void_p(h), n.value
    return w


def build_world_from_mesh(positions, colors32, indices, max_dimension, flips=(True, False, False), lods=LODS, save_to=None):
    """UnityManager's Convert button (UnityManager.cs:340-366) with the reference's own builder: returns (dims, blobs, column_counts, voxels)."""
    L = lib()
    v = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    c = np.ascontiguousarray(colors32, dtype=np.uint8).reshape(-1, 4)
    assert c.shape[0] == v.shape[0]
    i = np.ascontiguousarray(indices, dtype=np.int32).ravel()
    h = L.ref_build_world_from_mesh(v.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), v.shape[0], i.ctypes.data_as(C.c_void_p), i.size, int(max_dimension),
                                    int(flips[0]), int(flips[1]), int(flips[2]), int(lods))
    if not h:
        raise RuntimeError("ref_build_world_from_mesh: " + L.ref_last_error().decode())
    h = C.c_void_p(h)
    try:
        dims = (C.c_int32 * 3)()
        nbytes = (C.c_int64 * LODS)()
        ccs = (C.c_int32 * LODS)()
        vox = (C.c_int32 * LODS)()
        L.ref_built_world_info(h, dims, nbytes, ccs, vox)
        if save_to is not None:   # WorldSaveFile.Serialize of exactly these worlds
            _check(L.ref_built_world_save(h, str(save_to).encode()), "ref_built_world_save")
        blobs = []
        for j in range(lods):
            b = np.zeros(nbytes[j], dtype=np.uint8)
            assert L.ref_built_world_copy_blob(h, j, b.ctypes.data_as(C.c_void_p), b.nbytes) == 0
            blobs.append(b)
        return tuple(dims), blobs, list(ccs[:lods]), list(vox[:lods])
    finally:
        L.ref_built_world_free(h)


def hardware_threads() -> int:
    return int(lib().ref_hardware_threads())
