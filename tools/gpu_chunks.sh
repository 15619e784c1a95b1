# --mode rays at N GPUs: views in flight / ring slots sweep (chunk 512) on config 1 (60 poses, 4K) and config 4 (8K)
N=$1
for k in "6 8" "8 16"; do
set -- $k
for cfg in 1 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus $N --config $cfg --mode rays --steps 3 --warmup 2 --no-cpu-baseline --shard-chunk 512 --inflight $1 --ring-slots $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N config $cfg inflight $1 ring $2 value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
done; done
