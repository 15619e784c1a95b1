"""Generates tests/golden/golden_ref_v1.json with oracle/_ref — the REFERENCE'S OWN C# sources translated to C++ by
oracle/refbuild/cs2cpp.py and run here (see oracle/ref.py). Needs /root/reference (this container only); the vectors are
committed so that the GPU box, where the reference tree does not exist, can check against them.

Per case: the reference's RenderManager.DrawWorld end to end (vanishing point, segment setup, CameraData, the four jobs of
DrawSegmentRayJob, the blit) from a camera pose — segment ray counts, vanishing point, CRC32 of the two raybuffers.
World blobs: datasets/mill.obj through the reference's own voxelizer / RLE builder / DownSample; the synthetic worlds come
from this repository's generators (their blob CRCs are recorded so a drift is noticed).
    python tests/golden/make_golden_ref.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpuvox_b200 as cv  # noqa: E402  (host side only: synthetic world generators, pose helpers)
from conftest import MILL, POSES, crc, limited, pose_for, parse_obj  # noqa: E402
from oracle import ref  # noqa: E402

RESOLUTIONS = [(320, 180), (333, 217), (256, 400)]


def worlds():
    P, Cc = parse_obj(MILL)
    dims, blobs, ccs, vox = ref.build_world_from_mesh(P, Cc, np.arange(P.shape[0]), 256)
    yield "mill256", cv.World(dims, blobs, ccs, vox)
    yield "terrain256", cv.World.synthetic(0, (256, 256, 256), seed=1234)
    yield "structure512x128x256", cv.World.synthetic(1, (512, 128, 256), seed=7)


def main():
    assert ref.build(force=True), "needs /root/reference"
    out = {"version": 1, "generator": ref.describe(), "worlds": {}}
    for name, w in worlds():
        rw = ref.RefWorld(w.dims, w.blobs, w.column_counts)
        cases = []
        for spec in POSES:
            for (W, H) in RESOLUTIONS:
                pose = limited(cv, pose_for(cv, w, spec))
                lods = cv.setup_lods(w.max_dimension, W, H)
                td, lr, frame = ref.draw_world(rw, pose.position, pose.rotation, W, H, lods, far=pose.far_clip, threads=1)
                s = ref.frame_setup(pose.position, pose.rotation, W, H, lods, w.dims[1], far=pose.far_clip)
                cases.append({"pose": spec[0], "width": W, "height": H,
                              "position": [float(x) for x in pose.position], "rotation": [float(x) for x in pose.rotation],
                              "far_clip": float(pose.far_clip), "lod_distances": [float(x) for x in lods],
                              "ray_counts": [s.segments[k].ray_count for k in range(4)],
                              "vanishing_point": [float(s.vanishing_point_screen[0]), float(s.vanishing_point_screen[1])],
                              "setup_crc": crc(np.frombuffer(bytes(s), dtype=np.uint8)),
                              "td_crc": crc(td), "lr_crc": crc(lr)})
        out["worlds"][name] = {"dims": list(w.dims), "blob_crcs": [crc(b) for b in w.blobs], "cases": cases}
        print(name, len(cases), "cases")
    with open(os.path.join(ROOT, "tests", "golden", "golden_ref_v1.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
