bash tools/ab.sh minb5 minb6 minb8 2>&1 | tee gpurun_out/ab_minb.log
python tools/ray_timing.py --res 3840x2160 --poses 30,59 > gpurun_out/ray_timing_mill.log 2>&1
python tools/ray_timing.py --config2 --poses 0 > gpurun_out/ray_timing_c2.log 2>&1
tail -12 gpurun_out/ray_timing_mill.log gpurun_out/ray_timing_c2.log
