mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_misc_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio --clock-control none -k regex:phase1 --csv --log-file gpurun_out/group_stalls.csv python tools/group_probe.py ${RES:-1920x1080} > /dev/null
python - <<PY
import csv
rows=[l for l in open("gpurun_out/group_stalls.csv") if l.startswith('"')]
cur={}
for d in csv.DictReader(rows):
    cur.setdefault(d["ID"],{"k":d["Kernel Name"][23:34]})[d["Metric Name"].replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio","")]=d["Metric Value"]
seen=set()
for k,v in cur.items():
    if v["k"] in seen: continue
    seen.add(v["k"]); print(v)
PY
