"""Developer tool (GPU box): world production on the device vs on the host for datasets/mill.obj (times, blob equality)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv

maxdim = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rm = cv.RenderManager(0)
mill = os.path.join(ROOT, "tests", "data", "mill.obj")
for i in range(reps):
    t0 = time.perf_counter(); host = cv.World.from_obj(mill, maxdim); t1 = time.perf_counter()
    dev = rm.build_world_from_obj(mill, maxdim); t2 = time.perf_counter()
    same = all(np.array_equal(a, b) for a, b in zip(dev.blobs, host.blobs))
    print(f"mill {maxdim}^3 rep {i}: host {1000 * (t1 - t0):.1f} ms, device {1000 * (t2 - t1):.1f} ms, identical blobs: {same}, voxels {host.voxel_counts}", flush=True)
