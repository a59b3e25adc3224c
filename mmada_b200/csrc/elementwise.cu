// HBM-bound kernels of the LLaDA block: embedding gather, RMSNorm, rotary embedding.
//
//   embed    replaces  self.transformer.wte(input_ids)          models/modeling_llada.py:1222
//   rmsnorm  replaces  RMSLayerNorm.forward                     models/modeling_llada.py:315-329
//   rope     replaces  RotaryEmbedding.forward/apply_rotary...  models/modeling_llada.py:402-428
//
// The residual stream is kept in fp32 (the reference keeps it in the model dtype); RMSNorm reads it
// and writes the bf16 A-operand of the following GEMM, so every row is read once and written once
// at half width.  All accesses are 128-bit and coalesced along the feature dimension.
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- embedding gather: out fp32 [M, d] = bf16 table[ids[m], :] ---------------------------------
__global__ void __launch_bounds__(256) embed_kernel(const int64_t* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                                                    float* __restrict__ out, int M, int d, int64_t vocab) {
    const int chunks = d >> 3;   // 8 bf16 = 16 bytes per chunk
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)M * chunks;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / chunks), c = (int)(i % chunks);
        // an id outside the table (nn.Embedding raises a device assert there, reference modeling_llada.py:1222) yields a
        // row of NaNs: every logit of the sequence turns NaN instead of the bad id being clamped to a valid row
        const int64_t id = ids[m];
        const bool ok = id >= 0 && id < vocab;
        const uint4 w = ok ? __ldg(reinterpret_cast<const uint4*>(table + id * d) + c)
                           : make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
        float4 a, b;
        float2 t;
        t = __bfloat1622float2(h[0]); a.x = t.x; a.y = t.y;
        t = __bfloat1622float2(h[1]); a.z = t.x; a.w = t.y;
        t = __bfloat1622float2(h[2]); b.x = t.x; b.y = t.y;
        t = __bfloat1622float2(h[3]); b.z = t.x; b.w = t.y;
        float4* o = reinterpret_cast<float4*>(out + (int64_t)m * d) + 2 * c;
        o[0] = a;
        o[1] = b;
    }
}

// ---- row gather: dst[i, :] = src[rows[i], :], rows of row_bytes (a multiple of 16) bytes, 128-bit copies ----------
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ src, int64_t ld_src16,
                                                          const int32_t* __restrict__ rows, uint4* __restrict__ dst,
                                                          int n_rows, int chunks) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n_rows * chunks;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / chunks), c = (int)(i % chunks);
        dst[i] = __ldcs(src + (int64_t)rows[r] * ld_src16 + c);
    }
}

// ---- embedding gather for the folded-RMSNorm pipeline: also xb bf16 [M, d] (the table row itself = the A operand of
// the first q|k|v GEMM) and ssq[m] = sum of the row's squares.  One warp per token row.
__global__ void __launch_bounds__(256) embed_norm_kernel(const int64_t* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                                                         float* __restrict__ out, __nv_bfloat16* __restrict__ xb,
                                                         float* __restrict__ ssq, int M, int d, int64_t vocab) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const int64_t id = ids[m];
    const bool ok = id >= 0 && id < vocab;                     // out of range: a row of NaNs (see embed_kernel)
    const uint4* src = reinterpret_cast<const uint4*>(table + (ok ? id : 0) * d);
    uint4* xbr = reinterpret_cast<uint4*>(xb + (int64_t)m * d);
    float4* o = reinterpret_cast<float4*>(out + (int64_t)m * d);
    float ss = 0.f;
    for (int c = lane; c < (d >> 3); c += 32) {
        const uint4 w = ok ? __ldg(src + c) : make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
        float4 a, b;
        float2 t;
        t = __bfloat1622float2(h[0]); a.x = t.x; a.y = t.y;
        t = __bfloat1622float2(h[1]); a.z = t.x; a.w = t.y;
        t = __bfloat1622float2(h[2]); b.x = t.x; b.y = t.y;
        t = __bfloat1622float2(h[3]); b.z = t.x; b.w = t.y;
        ss += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w) + (b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w);
        o[2 * c] = a;
        o[2 * c + 1] = b;
        xbr[c] = w;
    }
    ss = warp_sum(ss);
    if (lane == 0) ssq[m] = ss;
}

// ---- RMSNorm: out bf16 [Mo, d] = (x * rsqrt(mean(x^2) + eps)) * w, one warp per row ------------
// rows != nullptr gathers: output row i is computed from input row rows[i].
template <int VEC>   // float4 loads per lane; d == 128 * VEC
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      __nv_bfloat16* __restrict__ out, const int32_t* __restrict__ rows,
                                                      int Mo, int d, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= Mo) return;
    const int64_t src = rows ? rows[warp] : warp;
    const float4* xr = reinterpret_cast<const float4*>(x + src * d);
    float4 v[VEC];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        v[i] = xr[i * 32 + lane];
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)d + eps);
    const float4* wr = reinterpret_cast<const float4*>(w);
    uint2* o = reinterpret_cast<uint2*>(out + (int64_t)warp * d);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float4 g = __ldg(wr + i * 32 + lane);
        o[i * 32 + lane] = make_uint2(pack_bf16(v[i].x * rstd * g.x, v[i].y * rstd * g.y),
                                      pack_bf16(v[i].z * rstd * g.z, v[i].w * rstd * g.w));
    }
}

// generic fallback (any d % 4 == 0): one warp per row, two passes over the row
__global__ void __launch_bounds__(256) rmsnorm_generic_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              __nv_bfloat16* __restrict__ out,
                                                              const int32_t* __restrict__ rows, int Mo, int d, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= Mo) return;
    const int64_t src = rows ? rows[warp] : warp;
    const float4* xr = reinterpret_cast<const float4*>(x + src * d);
    float ss = 0.f;
    for (int i = lane; i < d / 4; i += 32) {
        const float4 v = xr[i];
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)d + eps);
    const float4* wr = reinterpret_cast<const float4*>(w);
    uint2* o = reinterpret_cast<uint2*>(out + (int64_t)warp * d);
    for (int i = lane; i < d / 4; i += 32) {
        const float4 v = xr[i];
        const float4 g = __ldg(wr + i);
        o[i] = make_uint2(pack_bf16(v.x * rstd * g.x, v.y * rstd * g.y), pack_bf16(v.z * rstd * g.z, v.w * rstd * g.w));
    }
}

// ---- rotary embedding, in place on the q and k thirds of the fused qkv activations ----------------
// qkv bf16 [M, ld]; q at column 0, k at column d; row m has position m % seq_len.
// sin/cos fp32 [>= seq_len, hd/2] are the reference's tables (the two halves of `positions` are
// identical, modeling_llada.py:393).  out[i] = t[i]*cos + rot[i]*sin, rot = (-t[i+hd/2], t[i-hd/2]),
// evaluated in fp32 as two products and a sum like the eager reference.
__global__ void __launch_bounds__(256) rope_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ sin_t,
                                                   const float* __restrict__ cos_t, int M, int ld, int d, int hd,
                                                   int seq_len) {
    const int half = hd >> 1;
    const int pairs_per_row = 2 * (d / hd) * (half >> 2);    // q and k, 4 pairs (8 elements) per thread
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)M * pairs_per_row;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / pairs_per_row);
        int r = (int)(i % pairs_per_row);
        const int per_head = half >> 2;
        const int j4 = r % per_head;            // which group of 4 within the half
        r /= per_head;
        const int head = r % (d / hd);
        const int which = r / (d / hd);         // 0 = q, 1 = k
        const int pos = m % seq_len;
        __nv_bfloat16* base = qkv + (int64_t)m * ld + which * d + head * hd + j4 * 4;
        uint2 lo_raw = *reinterpret_cast<uint2*>(base);
        uint2 hi_raw = *reinterpret_cast<uint2*>(base + half);
        const float4 sn = __ldg(reinterpret_cast<const float4*>(sin_t + (int64_t)pos * half) + j4);
        const float4 cs = __ldg(reinterpret_cast<const float4*>(cos_t + (int64_t)pos * half) + j4);
        float lo[4], hi[4];
        {
            const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&lo_raw);
            const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&hi_raw);
            float2 t;
            t = __bfloat1622float2(a[0]); lo[0] = t.x; lo[1] = t.y;
            t = __bfloat1622float2(a[1]); lo[2] = t.x; lo[3] = t.y;
            t = __bfloat1622float2(b[0]); hi[0] = t.x; hi[1] = t.y;
            t = __bfloat1622float2(b[1]); hi[2] = t.x; hi[3] = t.y;
        }
        const float s[4] = {sn.x, sn.y, sn.z, sn.w}, c[4] = {cs.x, cs.y, cs.z, cs.w};
        float olo[4], ohi[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            olo[t] = __fadd_rn(__fmul_rn(lo[t], c[t]), __fmul_rn(-hi[t], s[t]));
            ohi[t] = __fadd_rn(__fmul_rn(hi[t], c[t]), __fmul_rn(lo[t], s[t]));
        }
        *reinterpret_cast<uint2*>(base) = make_uint2(pack_bf16(olo[0], olo[1]), pack_bf16(olo[2], olo[3]));
        *reinterpret_cast<uint2*>(base + half) = make_uint2(pack_bf16(ohi[0], ohi[1]), pack_bf16(ohi[2], ohi[3]));
    }
}

}  // namespace mmada

using namespace mmada;

extern "C" int mmada_embed_f32(const int64_t* ids, const void* table_bf16, float* out, int M, int d, int64_t vocab,
                               void* stream) {
    if (!ids || !table_bf16 || !out || M <= 0 || d <= 0) return kBadArgument;
    if (d % 8) return kUnsupportedShape;
    const int64_t work = (int64_t)M * (d / 8);
    int blocks = (int)((work + 255) / 256);
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    embed_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ids, (const __nv_bfloat16*)table_bf16, out, M, d, vocab);
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_gather_rows(const void* src, int64_t ld_src_bytes, const int32_t* rows, void* dst, int n_rows,
                                 int row_bytes, void* stream) {
    if (!src || !rows || !dst || n_rows <= 0 || row_bytes <= 0) return kBadArgument;
    if ((row_bytes % 16) || (ld_src_bytes % 16)) return kUnsupportedShape;
    if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) return kBadArgument;
    const int chunks = row_bytes / 16;
    const int64_t work = (int64_t)n_rows * chunks;
    int blocks = (int)((work + 255) / 256);
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, ld_src_bytes / 16, rows, (uint4*)dst,
                                                                 n_rows, chunks);
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_embed_norm_f32(const int64_t* ids, const void* table_bf16, float* out, void* xb_bf16, float* ssq_out,
                                    int M, int d, int64_t vocab, void* stream) {
    if (!ids || !table_bf16 || !out || !xb_bf16 || !ssq_out || M <= 0 || d <= 0) return kBadArgument;
    if (d % 8) return kUnsupportedShape;
    embed_norm_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(ids, (const __nv_bfloat16*)table_bf16, out,
                                                                     (__nv_bfloat16*)xb_bf16, ssq_out, M, d, vocab);
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_rmsnorm_bf16(const float* x, const float* weight, void* out_bf16, const int32_t* rows, int M_out,
                                  int d, float eps, void* stream) {
    if (!x || !weight || !out_bf16 || M_out <= 0 || d <= 0) return kBadArgument;
    if (d % 4) return kUnsupportedShape;
    const int blocks = (M_out + 7) / 8;   // 8 warps per block, one row per warp
    cudaStream_t s = (cudaStream_t)stream;
    __nv_bfloat16* o = (__nv_bfloat16*)out_bf16;
    switch (d) {
        case 256: rmsnorm_kernel<2><<<blocks, 256, 0, s>>>(x, weight, o, rows, M_out, d, eps); break;
        case 512: rmsnorm_kernel<4><<<blocks, 256, 0, s>>>(x, weight, o, rows, M_out, d, eps); break;
        case 1024: rmsnorm_kernel<8><<<blocks, 256, 0, s>>>(x, weight, o, rows, M_out, d, eps); break;
        case 4096: rmsnorm_kernel<32><<<blocks, 256, 0, s>>>(x, weight, o, rows, M_out, d, eps); break;
        default: rmsnorm_generic_kernel<<<blocks, 256, 0, s>>>(x, weight, o, rows, M_out, d, eps); break;
    }
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_rope_inplace_bf16(void* qkv_bf16, int64_t ld, const float* sin_table, const float* cos_table, int M,
                                       int d_model, int head_dim, int seq_len, void* stream) {
    if (!qkv_bf16 || !sin_table || !cos_table || M <= 0 || seq_len <= 0) return kBadArgument;
    if (head_dim % 16 || d_model % head_dim || ld % 4) return kUnsupportedShape;
    const int64_t work = (int64_t)M * 2 * (d_model / head_dim) * (head_dim / 8);
    int blocks = (int)((work + 255) / 256);
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    rope_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)qkv_bf16, sin_table, cos_table, M, (int)ld,
                                                          d_model, head_dim, seq_len);
    return cuda_status(cudaGetLastError());
}
