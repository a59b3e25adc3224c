mkdir -p gpurun_out
for v in 10 11 12 13 7 10 12; do printf "variant $v: "; MMADA_SAMPLE_VARIANT=$v timeout 120 python scripts/bench_kernels.py --what sample 2>&1 | grep t2i_sample; done
for v in 10 12; do MMADA_SAMPLE_VARIANT=$v timeout 300 python -m pytest -x -q -m gpu tests/test_kernels_gpu.py tests/test_model_gpu.py -k "sample or t2i or t2m" 2>&1 | tail -1; done
