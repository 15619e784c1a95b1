# 2-GPU check after a change to the batch / sharded paths: multi-GPU parity tests + the driver's bench command at N = 2
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1; tail -6 gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n2.json")); print("N=2 value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), d["e2e"].get("d2h_ceiling", {}).get("frames_per_s_ceiling"), d.get("rays_sharded"))
PY
