# A/B kernel variants on the GPU box: tools/ab.sh <variant names...>  (default library first)
mkdir -p gpurun_out
run() { python bench.py --no-cpu-baseline --steps 6 --warmup 3 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'fps4k %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'fps1080 %.1f' % d['at_1080p']['value'], 'p1excl %.3f' % d['ms_per_frame']['exclusive_one_view_in_flight']['phase1_kernel'], 'p2excl %.4f' % d['ms_per_frame']['exclusive_one_view_in_flight']['phase2_kernel'], 'p2excl1080 %.4f' % d['at_1080p']['phase2_kernel_ms_exclusive'])"; }
run default
for v in "$@"; do CPUVOX_B200_LIB=$PWD/cpuvox_b200/variants/lib_$v.so run $v; done
