mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc; numactl -H 2>/dev/null | head -5; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $(basename $d) numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist); fi; done | head -12; free -g | head -2 ) > gpurun_out/topology.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python - <<'PY'
import time, os, sys
sys.path.insert(0, os.getcwd())
import cpuvox_b200 as cv
rm = cv.RenderManager(0)
for name, w in (("mill 1024", cv.World.from_obj("tests/data/mill.obj", 1024)), ("structures 4096x1024x4096", cv.World.synthetic(1, (4096, 1024, 4096), seed=7))):
    for rep in range(2):
        t0 = time.perf_counter(); rm.upload_world(w); rm.sync(); t1 = time.perf_counter()
        print(f"upload {name}: {1000*(t1-t0):.1f} ms ({sum(b.nbytes for b in w.blobs)/1e6:.0f} MB of blobs)", flush=True)
PY
