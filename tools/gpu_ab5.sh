mkdir -p gpurun_out
bash tools/ab.sh nofns 2>&1 | tee gpurun_out/ab_feat.log
bash tools/ab.sh nofns 2>&1 | tee -a gpurun_out/ab_feat.log
