"""Generates tests/golden/golden_v1.json with the CPU oracle (oracle/cpuvox_oracle.cpp).

The reference ships no golden vectors and cannot be run here (C#/Unity/Burst), so these
fixtures pin OUR restatement: counters, CRC32 of both raybuffers and of the final frame, the segment ray counts and
the vanishing point, per case. They guard the oracle, the world builders and the host setup against regressions and
give the GPU tests a second, file-based reference. Re-run only when a deliberate semantic change is made:
    python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpuvox_b200 as cv  # noqa: E402
from conftest import MILL, POSES, crc, setup_for  # noqa: E402
from oracle import oracle as orc  # noqa: E402

RESOLUTIONS = [(320, 180), (333, 217), (256, 400)]


def worlds():
    return {
        "terrain256": cv.World.synthetic(0, (256, 256, 256), seed=1234),
        "structure512x128x256": cv.World.synthetic(1, (512, 128, 256), seed=7),
        "mill256": cv.World.from_obj(MILL, 256),
    }


def case(world, ow, spec, W, H):
    s = setup_for(cv, world, spec, W, H)
    os_ = orc.copy_setup(s)
    td, lr, cn = orc.render_raybuffers(ow, os_, W, H, threads=1)
    frame = orc.blit(os_, W, H, td, lr, threads=1)
    return {
        "pose": spec[0], "width": W, "height": H,
        "ray_counts": [s.segments[k].ray_count for k in range(4)],
        "vanishing_point": [float(s.vanishing_point_screen[0]), float(s.vanishing_point_screen[1])],
        "counters": cn, "td_crc": crc(td), "lr_crc": crc(lr), "frame_crc": crc(frame),
    }


def main():
    out = {"version": 1, "worlds": {}}
    for name, w in worlds().items():
        ow = orc.OracleWorld(w.dims, w.blobs, w.column_counts)
        cases = [case(w, ow, spec, W, H) for spec in POSES for (W, H) in RESOLUTIONS]
        out["worlds"][name] = {"dims": list(w.dims), "blob_crcs": [crc(b) for b in w.blobs], "voxel_counts": list(w.voxel_counts), "cases": cases}
        print(name, len(cases), "cases")
    with open(os.path.join(ROOT, "tests", "golden", "golden_v1.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
