# A/B of Phase-2 variants under ncu (kernel time, instructions, issue utilisation), poses 30 and 59 at 4K
for v in default "$@"; do
  if [ "$v" = default ]; then unset CPUVOX_B200_LIB; else export CPUVOX_B200_LIB=$PWD/cpuvox_b200/variants/lib_$v.so; fi
  echo "== $v"
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:phase2 -c 4 python tools/one_frame.py --res 3840x2160 --poses 30,59 --reps 2 2>&1 | grep -E "gpu__time|inst_executed|issue_active" | awk '{printf "%s ", $NF} END {print ""}'
done
