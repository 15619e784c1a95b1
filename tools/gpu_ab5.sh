mkdir -p gpurun_out
bash tools/ab.sh nohot m20 m24 2>&1 | tee gpurun_out/ab_feat.log
