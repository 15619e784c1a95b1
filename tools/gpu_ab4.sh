mkdir -p gpurun_out
bash tools/ab.sh none noblind nomark noside 2>&1 | tee gpurun_out/ab_micro.log
timeout 300 python tools/configs_check.py --configs 2 --passes 2 --check 1 2>&1 | grep -E "^\{" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('config', d['config'], 'fps %.1f' % d['frames_per_s'], d['parity_vs_oracle'])"
