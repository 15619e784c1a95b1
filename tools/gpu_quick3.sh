mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "ring or batch or ranges or product or golden_fixtures" ) > gpurun_out/pytest_sel.log 2>&1; tail -15 gpurun_out/pytest_sel.log
python tools/shard_probe.py --config 4 --chunk 512 --ranks 1,4,8 2>&1 | tee gpurun_out/shard_probe_c4b.log | tail -14
