#!/bin/bash
# Round profile set (run under gpurun): launch list of a short bench, ncu --set full of the GEMM family and of the
# attention kernel.  Every profiled command is run once WITHOUT ncu first and must exit 0.  usage: profile_round.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
set -x
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/${tag}_bench_s2w3.json 2> gpurun_out/${tag}_bench_s2w3.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 300 python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/${tag}_bench_l2.json 2> gpurun_out/${tag}_bench_l2.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 24 -c 5 -f -o gpurun_out/${tag}_gemm \
  python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/${tag}_ncu_gemm.log 2>&1
timeout 120 python scripts/bench_kernels.py --what attn > gpurun_out/${tag}_attn_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_duo -s 3 -c 1 -f -o gpurun_out/${tag}_attn \
  python scripts/bench_kernels.py --what attn > gpurun_out/${tag}_ncu_attn.log 2>&1
ls -la gpurun_out
