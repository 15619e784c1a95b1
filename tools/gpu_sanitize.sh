# compute-sanitizer over subsets of the GPU tests: memcheck (all kernels incl. the world builder, transcode, debug views, fuzz), racecheck (shared-memory protocol of Phase 1)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck.log python -m pytest tests -m gpu -x -q -k "gpu_world_builder or debug_views or draw_world_batch or hand_made or error_paths or fuzz or golden or tall_columns" > gpurun_out/sanitizer_pytest.log 2>&1
echo "memcheck exit $?"; tail -2 gpurun_out/sanitizer_pytest.log; tail -2 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitizer_racecheck.log python -m pytest tests -m gpu -x -q -k "golden or tall_columns or hand_made" > gpurun_out/sanitizer_pytest2.log 2>&1
echo "racecheck exit $?"; tail -2 gpurun_out/sanitizer_pytest2.log; tail -4 gpurun_out/sanitizer_racecheck.log
