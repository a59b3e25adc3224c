mkdir -p gpurun_out
timeout 300 python -m pytest -x -q -m gpu tests/test_kernels_gpu.py -k "attention" > gpurun_out/t1.log 2>&1; echo "attn tests exit $?"; tail -3 gpurun_out/t1.log
for v in "MMADA_ATT_SPLIT_TAIL=0" "MMADA_ATT_SPLIT_TAIL=1"; do
  echo "== $v"; env $v timeout 120 python scripts/bench_kernels.py --what attn 2>&1 | grep attention
done
