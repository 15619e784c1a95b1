# end-of-session GPU round: parity tests, both bench arms, ncu launch list, instruction counts, full captures of Phase 1 and Phase 2
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-1080p --no-extras > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:phase1 -c 190 --csv --log-file gpurun_out/inst.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-1080p --no-extras > gpurun_out/bench_under_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase1 -s 2 -c 2 -f -o gpurun_out/prof_phase1_4k python tools/one_frame.py --res 3840x2160 --poses 30,59 --reps 2 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phase2 -s 2 -c 2 -f -o gpurun_out/prof_phase2_4k python tools/one_frame.py --res 3840x2160 --poses 30,59 --reps 2 > gpurun_out/ncu_full_p2.log 2>&1
tail -3 gpurun_out/ncu_full.log
