mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -s -k "gpu_world_builder or debug_views" ) > gpurun_out/pytest_f2.log 2>&1; tail -25 gpurun_out/pytest_f2.log
