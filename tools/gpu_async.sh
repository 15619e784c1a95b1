mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "asynchronous or batch or ring" ) > gpurun_out/pytest_sel.log 2>&1; tail -6 gpurun_out/pytest_sel.log
for f in "" "--e2e-sync"; do
python bench.py --no-cpu-baseline --no-extras --steps 10 --warmup 3 $f 2> gpurun_out/bench_async.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('e2e mode [$f]', 'value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], '1080p %.1f e2e %.1f' % (d['at_1080p']['value'], d['at_1080p']['e2e']))" || tail -5 gpurun_out/bench_async.err
done
