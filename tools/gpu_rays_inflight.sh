# config 4 (8K, rays of every view sharded, frame ring) at N GPUs with more views in flight: tools/gpu_rays_inflight.sh N
N=${1:-4}
mkdir -p gpurun_out
run() { if [ $N -gt 1 ]; then python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; else python bench.py "$@"; fi; }
for kr in "6 8" "12 16" "16 24"; do set -- $kr
run --config 4 --steps 24 --warmup 2 --no-cpu-baseline --inflight $1 --ring-slots $2 2> gpurun_out/c4_n${N}_k$1.err | tail -1 > gpurun_out/c4_n${N}_k$1.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c4_n${N}_k$1.json")); print("N=$N inflight $1 ring $2: value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("N=$N inflight $1 FAILED", e); print(open("gpurun_out/c4_n${N}_k$1.err").read()[-1500:])
PY
done
