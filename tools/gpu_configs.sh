# bench lines of BASELINE configs 2-5 (and config 1 in rays mode) on the GPU box; N = number of GPUs (torchrun when > 1)
N=${1:-1}; shift
mkdir -p gpurun_out
run() {
  if [ "$N" = 1 ]; then python bench.py "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N "$@"; fi
}
for c in "$@"; do
  ( time run --config $c --steps 3 --warmup 3 --no-cpu-baseline ) 2> gpurun_out/config${c}_n$N.err | tail -1 > gpurun_out/config${c}_n$N.json
  python -c "
import json
try:
    d=json.load(open('gpurun_out/config${c}_n$N.json'))
    print('config $c N=$N', d['config']['resolution'], 'value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], d['scaling'], 'roofline %.4f' % d['roofline']['frac'], d.get('at_inflight',{}).get('1'), d.get('rays_sharded',{}).get('value'))
except Exception as e:
    print('config $c FAILED', e); print(open('gpurun_out/config${c}_n$N.err').read()[-1500:])
"
done
