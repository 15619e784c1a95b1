mkdir -p gpurun_out
for k in 6 8 12 16; do
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-1080p --no-extras --inflight $k --inflight-e2e $k 2> gpurun_out/inflight_$k.err | tee gpurun_out/inflight_$k.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('inflight $k value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))" || tail -5 gpurun_out/inflight_$k.err
done
