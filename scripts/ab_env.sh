#!/bin/bash
# A/B of attention kernel selections by environment variable (each is read once per process):
#   usage: ab_env.sh "MMADA_ATT_PAIRX=0" "MMADA_ATT_PAIRX=1" ...   -> attention parity tests + isolated timing per setting
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  env $v timeout 300 python -m pytest -x -q -m gpu tests/test_kernels_gpu.py -k "test_attention and not split" 2>&1 | tail -2
  for rep in 1 2; do env $v timeout 120 python scripts/bench_kernels.py --what attn 2>&1 | grep "^attention"; done
done
