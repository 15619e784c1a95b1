mkdir -p gpurun_out
python tools/ray_timing.py --res 3840x2160 --poses 30,59 > gpurun_out/ray_timing_mill.log 2>&1
python tools/ray_timing.py --config2 --poses 0 > gpurun_out/ray_timing_c2.log 2>&1
cat gpurun_out/ray_timing_mill.log gpurun_out/ray_timing_c2.log
bash tools/ab.sh "$@" 2>&1 | tee gpurun_out/ab_last.log
