mkdir -p gpurun_out
bash tools/ab.sh r88 r72 r96 2>&1 | tee gpurun_out/ab_feat.log
