# A/B of kernel variants with a parity check of each variant first: tools/gpu_ab_parity.sh <variant names...>
mkdir -p gpurun_out
for v in "$@"; do
  CPUVOX_B200_LIB=$PWD/cpuvox_b200/variants/lib_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle_all_poses or hand_made or tall_columns" 2>&1 | tail -2
done
bash tools/ab.sh "$@"
