"""Developer tool (GPU box): throughput with K frames in flight (K contexts round-robin, each with its own stream and buffers)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvox_b200 as cv  # noqa: E402


def main():
    res = sys.argv[1] if len(sys.argv) > 1 else "3840x2160"
    W, H = [int(x) for x in res.split("x")]
    world = cv.World.from_obj(os.path.join(ROOT, "tests", "data", "mill.obj"), 1024)
    poses = cv.benchmark_path(world.dims, 60, far_clip=2.0 * world.max_dimension)
    for K in (1, 2, 3, 4, 6):
        rms = []
        for _ in range(K):
            rm = cv.RenderManager(0)
            rm.upload_world(world)
            rm.set_resolution(W, H)
            rms.append(rm)
        setups = [rms[0].make_setup(p) for p in poses]
        for rep in range(2):
            for i, s in enumerate(setups):
                rms[i % K].draw_setup(s)
        for rm in rms:
            rm.sync()
        t0 = time.perf_counter()
        reps = 5
        for rep in range(reps):
            for i, s in enumerate(setups):
                rms[i % K].draw_setup(s)
        for rm in rms:
            rm.sync()
        dt = time.perf_counter() - t0
        print(f"{res} K={K}: {reps * len(setups) / dt:.1f} frames/s", flush=True)
        for rm in rms:
            rm.destroy()


main()
