#!/bin/bash
# Debug build of the attention kernels with clock64 tracing of CTA 0 (see scripts/attn_trace.py)
set -e
cd "$(dirname "$0")/../mmada_b200/csrc"
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --compiler-options -fPIC -DMMADA_ATT_TRACE ${ATT_DEFS} \
     -shared -o build/libattn_trace.so attention.cu attention_pair.cu attention_duo.cu attention_quad.cu attention_duo64.cu attention_pair64.cu -cudart static
