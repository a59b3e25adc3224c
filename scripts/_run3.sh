mkdir -p gpurun_out
timeout 400 python -m pytest -x -q -m gpu tests/test_kernels_gpu.py -k "attention" > gpurun_out/t1.log 2>&1; echo "attn tests exit $?"; tail -5 gpurun_out/t1.log
for m in 0 1 2 0 1 2; do printf "tail mode $m: "; MMADA_ATT_TAIL=$m timeout 120 python scripts/bench_kernels.py --what attn 2>&1 | grep "^attention"; done
