# compute-sanitizer memcheck over the paths added in the second half of round 2: framebuffer pool + asynchronous batches, frame ring on one
# device (sharded Phase 1 / owned Phase 2, flag kernels), packed RGB8 / JPEG presentation
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer2_memcheck.log python -m pytest tests -m gpu -x -q -k "asynchronous or draw_batch_equals or frame_ring or present or debug_views or ranges" > gpurun_out/sanitizer2_pytest.log 2>&1
echo "memcheck exit $?"; tail -2 gpurun_out/sanitizer2_pytest.log; tail -3 gpurun_out/sanitizer2_memcheck.log
