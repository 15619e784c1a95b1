mkdir -p gpurun_out
bash tools/ab.sh minb5 minb6 2>&1 | tee gpurun_out/ab_minb2.log
bash tools/ab.sh minb6 2>&1 | tee -a gpurun_out/ab_minb2.log
