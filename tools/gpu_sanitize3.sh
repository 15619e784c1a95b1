# memcheck over the final kernels: Phase 1 / Phase 2 (golden, hand-made, tall columns, ray ranges, owned blits), batches (pool, asynchronous), frame ring, presentation
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer3_memcheck.log python -m pytest tests -m gpu -x -q -k "golden or hand_made or tall_columns or ranges or asynchronous or draw_batch_equals or frame_ring or present or debug_views or error_paths or fewer_lods" > gpurun_out/sanitizer3_pytest.log 2>&1
echo "memcheck exit $?"; tail -2 gpurun_out/sanitizer3_pytest.log; tail -2 gpurun_out/sanitizer3_memcheck.log
