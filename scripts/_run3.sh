mkdir -p gpurun_out
timeout 300 python -m pytest -x -q -m gpu tests/test_kernels_gpu.py -k "attention" > gpurun_out/t1.log 2>&1; echo "attn tests exit $?"; tail -3 gpurun_out/t1.log
timeout 120 python scripts/bench_kernels.py --what attn 2>&1 | grep -v "^$"
MMADA_ATT_SPLIT_TAIL=0 timeout 120 python scripts/attn_trace.py 2>&1 | tee gpurun_out/attn_trace_new.txt | sed -n 18,30p | cut -c1-200
