# group width x views in flight sweep (bench lines, no profiler), then BASELINE configs 2-5 at full size
mkdir -p gpurun_out
for g in 32 16 8; do for k in 4 8; do
python bench.py --no-cpu-baseline --steps 6 --warmup 3 --group $g --inflight $k 2>>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('group $g inflight $k', 'fps4k %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'fps1080 %.1f' % d['at_1080p']['value'], 'p1excl %.3f' % d['ms_per_frame']['exclusive_one_view_in_flight']['phase1_kernel'])" | tee -a gpurun_out/sweep.log
done; done
rm -f gpurun_out/configs.jsonl
timeout 900 python tools/configs_check.py > gpurun_out/configs.log 2>&1; tail -12 gpurun_out/configs.log
