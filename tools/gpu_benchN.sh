# the driver's SCALE command at N GPUs: tools/gpu_benchN.sh N
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json")); print("N=$N value %.1f e2e %.1f ceiling %.1f" % (d["value"], d["e2e"]["value"], d["e2e"]["d2h_ceiling"]["frames_per_s_ceiling"]), d.get("rays_sharded", {}).get("value"))
PY
