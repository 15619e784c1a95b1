# quick GPU round: selected parity tests + bench line (no profiler)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "batch or present or debug or product" ) > gpurun_out/pytest_sel.log 2>&1; tail -15 gpurun_out/pytest_sel.log
python bench.py --no-cpu-baseline "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
