bash tools/ab.sh t32_16 t32_20 t32_24 t64_10 2>&1 | tee gpurun_out/ab_cta.log
