#!/bin/bash
# A/B of attention kernel variants on a GPU box: needs the EXPERIMENTS=1 build (make -C mmada_b200/csrc EXPERIMENTS=1 ->
# mmada_b200/libmmada_b200_exp.so, which reads the MMADA_ATT_* switches).  usage: ab_attention.sh "MMADA_ATT_POLY=0" "MMADA_ATT_POLY=2" ...
export MMADA_B200_LIB=$PWD/mmada_b200/libmmada_b200_exp.so
for setting in "$@"; do
  echo "== $setting"
  env $setting timeout 120 python scripts/bench_kernels.py --what attn | grep attention
done
