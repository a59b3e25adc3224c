#!/bin/bash
# DRAM traffic and duration of the C2 GEMM shapes under different rasterisation / cache-hint settings
mkdir -p gpurun_out
for cfg in "8 nn" "8 ln" "8 ll" "16 nn" "16 ln" "4 ln" "8 lf"; do
  set -- $cfg
  export MMADA_GEMM_GROUP_M=$1 MMADA_GEMM_HINTS=$2
  echo "=== group_m=$1 hints=$2"
  python scripts/bench_kernels.py --what gemm 2>&1 | grep "gemm_" | sed 's/{"ms": //; s/, "tflops": / ms  /; s/}//'
done
