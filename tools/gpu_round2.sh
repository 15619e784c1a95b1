mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/configs.jsonl
timeout 900 python tools/configs_check.py > gpurun_out/configs.log 2>&1; tail -8 gpurun_out/configs.log
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:phase1 -c 190 --csv --log-file gpurun_out/inst.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-1080p > gpurun_out/bench_under_ncu2.log 2>&1
tail -2 gpurun_out/inst.csv
