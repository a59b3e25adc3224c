// Masked cross-entropy rows for the training-time forward (MMadaModelLM.forward_process).
//
// Replaces F.cross_entropy(logits_rows, labels, ignore_index, reduction='none') at
// /root/reference/models/modeling_mmada.py:240-243,253-256,264-267: per row
//   nll[r] = logsumexp(logits[r, :]) - logits[r, label[r]],   0 where label[r] == ignore_index.
// HBM-bound: one pass over the row (V = 134 656 fp32 = 539 KB), one CTA per row, 128-bit loads, every thread carries an
// online (maximum, sum of exponentials) pair over its strided share and the pairs are merged by warp shuffles — the
// reference's eager chain materialises log_softmax (reads + writes the whole [R, V] tensor again).
#include <math.h>

#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {
namespace {

constexpr int CE_THREADS = 256;

struct MS {
    float m, s;
};
__device__ __forceinline__ MS ms_merge(MS a, MS b) {
    const float m = fmaxf(a.m, b.m);
    if (m == -INFINITY) return MS{m, 0.f};
    return MS{m, a.s * __expf(a.m - m) + b.s * __expf(b.m - m)};
}

__global__ void __launch_bounds__(CE_THREADS) cross_entropy_rows_kernel(const float* __restrict__ logits, int64_t ld,
                                                                         const int64_t* __restrict__ labels,
                                                                         int64_t ignore_index, float* __restrict__ nll,
                                                                         int V) {
    const int r = blockIdx.x, tid = threadIdx.x;
    const int64_t label = labels[r];
    if (label == ignore_index || label < 0 || label >= V) {        // (an out-of-range label is the caller's error: 0, not a fault)
        if (tid == 0) nll[r] = (label == ignore_index) ? 0.f : NAN;
        return;
    }
    const float* row = logits + (int64_t)r * ld;
    MS acc{-INFINITY, 0.f};
    auto add4 = [&](float4 v) {
        const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        const float m = fmaxf(acc.m, m4);
        if (m == -INFINITY) return;
        acc.s = acc.s * __expf(acc.m - m) + (__expf(v.x - m) + __expf(v.y - m)) + (__expf(v.z - m) + __expf(v.w - m));
        acc.m = m;
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    const int V4 = vec ? V / 4 : 0;
    const float4* row4 = reinterpret_cast<const float4*>(row);
    int i = tid;
    // four 128-bit loads in flight per thread
    for (; i + 3 * CE_THREADS < V4; i += 4 * CE_THREADS) {
        const float4 a = __ldcs(row4 + i), b = __ldcs(row4 + i + CE_THREADS), c = __ldcs(row4 + i + 2 * CE_THREADS),
                     d = __ldcs(row4 + i + 3 * CE_THREADS);
        add4(a); add4(b); add4(c); add4(d);
    }
    for (; i < V4; i += CE_THREADS) add4(__ldcs(row4 + i));
    for (int j = 4 * V4 + tid; j < V; j += CE_THREADS) {
        const float x = row[j];
        acc = ms_merge(acc, MS{x, 1.f});
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MS other{__shfl_xor_sync(0xffffffffu, acc.m, o), __shfl_xor_sync(0xffffffffu, acc.s, o)};
        acc = ms_merge(acc, other);
    }
    __shared__ float sm[CE_THREADS / 32], ss[CE_THREADS / 32];
    if ((tid & 31) == 0) { sm[tid >> 5] = acc.m; ss[tid >> 5] = acc.s; }
    __syncthreads();
    if (tid < 32) {
        MS w = tid < CE_THREADS / 32 ? MS{sm[tid], ss[tid]} : MS{-INFINITY, 0.f};
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            MS other{__shfl_xor_sync(0xffffffffu, w.m, o), __shfl_xor_sync(0xffffffffu, w.s, o)};
            w = ms_merge(w, other);
        }
        if (tid == 0) nll[r] = (w.m + logf(w.s)) - row[label];
    }
}

}  // namespace
}  // namespace mmada

using namespace mmada;

extern "C" int mmada_cross_entropy_rows_f32(const float* logits, int64_t ld, const int64_t* labels, int64_t ignore_index,
                                            float* nll_out, int R, int V, void* stream) {
    if (!logits || !labels || !nll_out || R < 0 || V <= 0 || ld < V) return kBadArgument;
    if (R == 0) return kOk;
    cross_entropy_rows_kernel<<<R, CE_THREADS, 0, (cudaStream_t)stream>>>(logits, ld, labels, ignore_index, nll_out, V);
    return cuda_status(cudaGetLastError());
}
