mkdir -p gpurun_out
M="smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,sm__icc_request_hit_rate.pct,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,launch__registers_per_thread"
for v in default minb5 minb6 minb8; do
  if [ $v = default ]; then unset CPUVOX_B200_LIB; else export CPUVOX_B200_LIB=$PWD/cpuvox_b200/variants/lib_$v.so; fi
  ncu --metrics $M --clock-control none -k regex:phase1 -s 1 -c 1 --csv --log-file gpurun_out/var_$v.csv python tools/one_frame.py --res 3840x2160 --poses 59 --reps 2 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/var_$v.csv')) if len(r)>10 and r[0]!='ID']
print('$v', ' '.join('%s=%s'%(r[12].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','').replace('.avg.pct_of_peak_sustained_active',''),r[14]) for r in rows))
PY
done | tee gpurun_out/var_summary.log
