# e2e (frames to pinned host memory) vs views in flight, on the GPU box
for k in 2 4 6 8; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-1080p --no-extras --inflight-e2e $k 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight_e2e $k', 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
done
