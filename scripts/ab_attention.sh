#!/bin/bash
# A/B of attention kernel builds: every variants/lib_*.so (built by hand from other commits / -D flags) against the
# in-tree library, interleaved and repeated so that clock drift shows up as noise rather than as a difference.
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in "" variants/lib_*.so; do
    if [ -n "$lib" ]; then export MMADA_B200_LIB=$PWD/$lib; else unset MMADA_B200_LIB; fi
    printf "%-28s " "${lib:-in-tree}"
    timeout 120 python scripts/bench_kernels.py --what attn 2>&1 | grep '^attention'
  done
done
