#!/usr/bin/env python3
"""Probe for a .NET SDK and, when there is one, build oracle/dotnet (the reference's own C# linked by path + Unity stand-ins),
render the golden cases with it and compare with tests/golden/golden_ref_v1.json (made by oracle/_ref, the same sources
translated to C++).  TEST INFRASTRUCTURE ONLY.

    python tools/dotnet_ref.py [--reference /path/to/cpuvox] [--probe-only]

Exit 0 with {"dotnet": null} when no SDK is found (this image and the GPU box: profiles/r02_probe.md)."""
from __future__ import annotations

import argparse
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def probe():
    found = {}
    for exe in ("dotnet", "mono", "mcs", "csc", "msbuild"):
        p = shutil.which(exe) or next((c for c in (os.path.expanduser("~/.dotnet/" + exe), "/usr/share/dotnet/" + exe, "/usr/lib/dotnet/" + exe) if os.path.exists(c)), None)
        if p:
            found[exe] = p
    return found


def write_case(path, W, H, world, case):
    with open(path, "wb") as f:
        f.write(struct.pack("<6i", W, H, *world.dims, len(world.blobs)))
        for b, cc in zip(world.blobs, world.column_counts):
            b = np.ascontiguousarray(b)
            f.write(struct.pack("<iq", int(cc), b.nbytes))
            f.write(b.tobytes())
        f.write(struct.pack("<3f4f3f6f", *case["position"], *case["rotation"], 85.0, 0.05, case["far_clip"], *case["lod_distances"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--probe-only", action="store_true")
    args = ap.parse_args()
    found = probe()
    if "dotnet" not in found or args.probe_only:
        print(json.dumps({"dotnet": found.get("dotnet"), "found": found, "note": "no .NET SDK: oracle/dotnet not built; oracle/_ref (C++ translation) is the runnable pin"}))
        return 0
    if not os.path.isdir(os.path.join(args.reference, "Assets", "Code")):
        print(json.dumps({"dotnet": found["dotnet"], "error": "reference checkout not found at " + args.reference}))
        return 1
    out = tempfile.mkdtemp(prefix="cpuvox_dotnet_")
    subprocess.check_call([found["dotnet"], "build", os.path.join(ROOT, "oracle", "dotnet"), "-c", "Release", "-o", out, "-p:CpuvoxReference=" + args.reference])
    exe = os.path.join(out, "cpuvox_ref_dotnet")
    import cpuvox_b200 as cv
    from conftest import MILL
    with open(os.path.join(ROOT, "tests", "golden", "golden_ref_v1.json")) as f:
        golden = json.load(f)["worlds"]
    worlds = {"terrain256": cv.World.synthetic(0, (256, 256, 256), seed=1234), "structure512x128x256": cv.World.synthetic(1, (512, 128, 256), seed=7),
              "mill256": cv.World.from_obj(MILL, 256)}
    ok = bad = 0
    for name, g in golden.items():
        for c in g["cases"]:
            W, H = c["width"], c["height"]
            cp, op = os.path.join(out, "case.bin"), os.path.join(out, "out.bin")
            write_case(cp, W, H, worlds[name], c)
            subprocess.check_call([exe, "render", cp, op])
            raw = np.fromfile(op, dtype=np.uint32)
            td, lr = raw[:(W + 2 * H) * H], raw[(W + 2 * H) * H:]
            same = (zlib.crc32(td.tobytes()), zlib.crc32(lr.tobytes())) == (c["td_crc"], c["lr_crc"])
            ok += same
            bad += not same
    print(json.dumps({"dotnet": found["dotnet"], "cases_equal": ok, "cases_different": bad}))
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
