# 8-GPU round: the driver's SCALE command at N = 8 and config 4 (rays sharded; 96 views so that the pipeline's ramp does not dominate); config 5 with a third argument
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
( time run --steps 10 --warmup 3 ) 2> gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
( time run --config 4 --steps 24 --warmup 2 --no-cpu-baseline ) 2> gpurun_out/config4_n$N.err | tail -1 > gpurun_out/config4_n$N.json
[ -n "$2" ] && ( time run --config 5 --steps 4 --warmup 3 --no-cpu-baseline ) 2> gpurun_out/config5_n$N.err | tail -1 > gpurun_out/config5_n$N.json
python - <<PY
import json, os
for f in ("bench_n$N", "config4_n$N", "config5_n$N"):
    if not os.path.exists("gpurun_out/%s.json" % f): continue
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d["config"]["resolution"], "value %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"], d["scaling"], d.get("rays_sharded"), (d["e2e"].get("d2h_ceiling") or {}).get("job_gbs"), d.get("at_inflight", {}).get("1"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
