// MAGVIT-v2 token -> pixel path: look-up-free quantiser maps and the HBM-bound kernels of the VQGAN
// decoder.  The convolutions themselves run as implicit GEMMs on tcgen05 (gemm.cu,
// mmada_conv_nhwc_bf16); activations are NHWC so that channels are the contiguous (K) dimension.
//
//   lfq_*            LFQuantizer.get_codebook_entry / get_indices   models/modeling_magvitv2.py:186-221
//   groupnorm_*      Normalize = GroupNorm(32, eps 1e-6) + swish     models/common_modules.py:16-24
//   upsample2x       F.interpolate(scale 2, nearest)                 models/common_modules.py:37
//   softmax_rows     AttnBlock softmax over keys                     models/common_modules.py:203
//   image_to_uint8   clamp((x+1)/2,0,1)*255 -> uint8                 inference_t2i.py:123-125
#include <math.h>

#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

constexpr int CODE_BITS = 13;

// indices [B*N] -> bf16 NHWC [B*N, 64]: channels 0..12 = post_quant_conv(bits) (1x1 conv 13->13), rest 0
__global__ void lfq_decode_kernel(const int64_t* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ b,
                                  __nv_bfloat16* __restrict__ out, int total, int64_t max_code) {
    __shared__ float sw[CODE_BITS * CODE_BITS + CODE_BITS];
    for (int i = threadIdx.x; i < CODE_BITS * CODE_BITS + CODE_BITS; i += blockDim.x)
        sw[i] = i < CODE_BITS * CODE_BITS ? w[i] : b[i - CODE_BITS * CODE_BITS];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    int64_t code = idx[p];
    code = code < 0 ? 0 : (code > max_code ? max_code : code);
    float z[CODE_BITS];
#pragma unroll
    for (int k = 0; k < CODE_BITS; ++k) z[k] = ((code >> (CODE_BITS - 1 - k)) & 1) ? 1.f : -1.f;   // MSB first
    uint32_t packed[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) packed[i] = 0;
    float o[CODE_BITS + 1];
    o[CODE_BITS] = 0.f;
#pragma unroll
    for (int co = 0; co < CODE_BITS; ++co) {
        float acc = 0.f;
#pragma unroll
        for (int ci = 0; ci < CODE_BITS; ++ci) acc = fmaf(sw[co * CODE_BITS + ci], z[ci], acc);
        o[co] = acc + sw[CODE_BITS * CODE_BITS + co];
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) packed[i] = pack_bf16(o[2 * i], o[2 * i + 1]);
    uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)p * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
}

// indices [B, N] -> fp32 NCHW [B, 13, h*w] of -1/+1
__global__ void lfq_bits_kernel(const int64_t* __restrict__ idx, float* __restrict__ out, int B, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N, n = i % N;
    const int64_t code = idx[i];
#pragma unroll
    for (int k = 0; k < CODE_BITS; ++k)
        out[((int64_t)b * CODE_BITS + k) * N + n] = ((code >> (CODE_BITS - 1 - k)) & 1) ? 1.f : -1.f;
}
// fp32 NCHW [B, 13, N] -> int64 [B, N]: sum 2^(12-k) [z_k > 0]
__global__ void lfq_index_kernel(const float* __restrict__ z, int64_t* __restrict__ out, int B, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N, n = i % N;
    int64_t code = 0;
#pragma unroll
    for (int k = 0; k < CODE_BITS; ++k)
        code |= (int64_t)(z[((int64_t)b * CODE_BITS + k) * N + n] > 0.f) << (CODE_BITS - 1 - k);
    out[i] = code;
}

// ---- GroupNorm(32): statistics -------------------------------------------------------------
// x fp32 NHWC [B, P, C]; sums double [B, 32, 2] (sum, sum of squares), zeroed by the launcher.
// Each thread keeps a fixed group of 4 channels and walks pixels, so loads are coalesced float4.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, int P, int C,
                                                       int pix_per_cta) {
    __shared__ double s_acc[32][2];
    const int b = blockIdx.y;
    const int c4n = C >> 2;                       // float4 per pixel
    const int c4 = threadIdx.x % c4n;
    const int prow = threadIdx.x / c4n;
    const int pstep = blockDim.x / c4n;
    if (threadIdx.x < 64) s_acc[threadIdx.x >> 1][threadIdx.x & 1] = 0.0;
    __syncthreads();
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(P, p0 + pix_per_cta);
    float s = 0.f, ss = 0.f;
    const float4* xb = reinterpret_cast<const float4*>(x + (int64_t)b * P * C);
    for (int p = p0 + prow; p < p1; p += pstep) {
        const float4 v = xb[(int64_t)p * c4n + c4];
        s += v.x + v.y + v.z + v.w;
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const int g = (c4 * 4) / (C / 32);
    atomicAdd(&s_acc[g][0], (double)s);
    atomicAdd(&s_acc[g][1], (double)ss);
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(&sums[((int64_t)b * 32 + (threadIdx.x >> 1)) * 2 + (threadIdx.x & 1)],
                                    s_acc[threadIdx.x >> 1][threadIdx.x & 1]);
}

// y = (x - mean) * rstd * gamma + beta, optional swish, -> bf16 NHWC
template <bool SWISH>
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       __nv_bfloat16* __restrict__ out, int P, int C, float eps, int64_t total4) {
    const int c4n = C >> 2;
    const double inv_n = 1.0 / ((double)P * (C / 32));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        const int64_t pix = i / c4n;
        const int b = (int)(pix / P);
        const int g = (c4 * 4) / (C / 32);
        const double* sm = sums + ((int64_t)b * 32 + g) * 2;
        const double mean = sm[0] * inv_n;
        const double var = fmax(sm[1] * inv_n - mean * mean, 0.0);
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float mu = (float)mean;
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        float y[4] = {(v.x - mu) * rstd * ga.x + be.x, (v.y - mu) * rstd * ga.y + be.y, (v.z - mu) * rstd * ga.z + be.z,
                      (v.w - mu) * rstd * ga.w + be.w};
        if (SWISH) {
#pragma unroll
            for (int t = 0; t < 4; ++t) y[t] = y[t] / (1.0f + __expf(-y[t]));
        }
        reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]));
    }
}

// nearest 2x: fp32 NHWC [B,H,W,C] -> bf16 NHWC [B,2H,2W,C]
__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int H,
                                                         int W, int C, int64_t total4) {
    const int c4n = C >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t p = i / c4n;                     // output pixel
        const int ox = (int)(p % (2 * W));
        p /= 2 * W;
        const int oy = (int)(p % (2 * H));
        const int64_t b = p / (2 * H);
        const float4 v = reinterpret_cast<const float4*>(x)[((b * H + (oy >> 1)) * W + (ox >> 1)) * c4n + c4];
        reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        int64_t total4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

// row softmax of scale * x: fp32 [R, n] -> bf16 [R, n]; one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int R,
                                                           int n, float scale) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= R) return;
    const float* xr = x + (int64_t)row * n;
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, xr[i] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) sum += expf(xr[i] * scale - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int i = lane; i < n; i += 32) out[(int64_t)row * n + i] = __float2bfloat16_rn(expf(xr[i] * scale - mx) * inv);
}

// fp32 NHWC [B,P,C] -> fp32 NCHW [B,C,P]  (C small: the 3 output channels)
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, int P, int C, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i % P;
        const int64_t bc = i / P;
        const int c = (int)(bc % C);
        const int64_t b = bc / C;
        out[i] = x[(b * P + p) * C + c];
    }
}
// fp32 NHWC in [-1,1] -> uint8 NHWC: trunc(clamp((x+1)/2, 0, 1) * 255)
__global__ void image_to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float v = __fdiv_rn(__fadd_rn(x[i], 1.0f), 2.0f);
        v = fminf(fmaxf(v, 0.f), 1.f);
        out[i] = (uint8_t)(__fmul_rn(v, 255.0f));
    }
}

static inline int grid_for(int64_t work, int threads = 256) {
    int64_t b = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * 32;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace mmada

using namespace mmada;

// ---- encoder side (MAGVITv2.get_code: models/modeling_magvitv2.py:143-169, 423-427) ----------------------------
// pixel_values fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64] (channels 3..63 zero): the input of conv_in as an
// implicit GEMM with one whole 64-channel k-block per tap.
__global__ void image_to_nhwc64_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int P, int64_t total) {
    // one thread per (pixel, 8-channel group): 8 groups of 16 bytes per pixel
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int grp = (int)(i & 7);
        const int64_t pix = i >> 3;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (grp == 0) {
            const int64_t b = pix / P, pp = pix - b * P;
            const float* src = x + b * 3 * (int64_t)P + pp;
            w.x = pack_bf16(src[0], src[P]);
            w.y = pack_bf16(src[2 * (int64_t)P], 0.f);
        }
        reinterpret_cast<uint4*>(out)[i] = w;
    }
}
// Downsample (models/common_modules.py:73-90: zero-pad right/bottom by one, 3x3 convolution with stride 2) runs as
// a stride-1 convolution over the SPACE-TO-DEPTH image: fp32 NHWC [B,H,W,C] -> bf16 NHWC [B,H/2,W/2,4C], channel
// (2 sy + sx) C + c = pixel (2y + sy, 2x + sx).  The host rearranges the 3x3 weights accordingly.
__global__ void space_to_depth2_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int H, int W, int C,
                                       int64_t total4) {
    const int C4 = C >> 2, Ho = H >> 1, Wo = W >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT in units of 4 channels
        const int c4 = (int)(i % C4);
        int64_t r = i / C4;
        const int s = (int)(r & 3); r >>= 2;
        const int xo = (int)(r % Wo); r /= Wo;
        const int yo = (int)(r % Ho);
        const int64_t b = r / Ho;
        const float4 v = *reinterpret_cast<const float4*>(x + (((b * H + 2 * yo + (s >> 1)) * W + 2 * xo + (s & 1)) * (int64_t)C + 4 * c4));
        reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

extern "C" int mmada_lfq_decode_nhwc(const int64_t* indices, const float* pq_weight, const float* pq_bias, void* out_bf16,
                                     int total_tokens, void* stream) {
    if (!indices || !pq_weight || !pq_bias || !out_bf16 || total_tokens <= 0) return kBadArgument;
    lfq_decode_kernel<<<(total_tokens + 127) / 128, 128, 0, (cudaStream_t)stream>>>(indices, pq_weight, pq_bias,
                                                                                   (__nv_bfloat16*)out_bf16, total_tokens, 8191);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_lfq_indices_to_bits(const int64_t* indices, float* out_nchw, int B, int N, void* stream) {
    if (!indices || !out_nchw || B <= 0 || N <= 0) return kBadArgument;
    lfq_bits_kernel<<<(B * N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(indices, out_nchw, B, N);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_lfq_bits_to_indices(const float* z_nchw, int64_t* out, int B, int N, void* stream) {
    if (!z_nchw || !out || B <= 0 || N <= 0) return kBadArgument;
    lfq_index_kernel<<<(B * N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(z_nchw, out, B, N);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_groupnorm_stats(const float* x, double* sums, int B, int P, int C, void* stream) {
    if (!x || !sums || B <= 0 || P <= 0) return kBadArgument;
    if (C % 128 || C > 1024) return kUnsupportedShape;        // 32 groups of a multiple of 4 channels, C/4 <= 256 threads
    cudaStream_t s = (cudaStream_t)stream;
    MMADA_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * B * 64, s));
    const int pix_per_cta = 1024;
    dim3 grid((P + pix_per_cta - 1) / pix_per_cta, B);
    gn_stats_kernel<<<grid, 256, 0, s>>>(x, sums, P, C, pix_per_cta);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_groupnorm_apply_bf16(const float* x, const double* sums, const float* gamma, const float* beta,
                                          void* out_bf16, int B, int P, int C, float eps, int swish, void* stream) {
    if (!x || !sums || !gamma || !beta || !out_bf16) return kBadArgument;
    if (C % 128) return kUnsupportedShape;
    const int64_t total4 = (int64_t)B * P * C / 4;
    cudaStream_t s = (cudaStream_t)stream;
    if (swish) gn_apply_kernel<true><<<grid_for(total4), 256, 0, s>>>(x, sums, gamma, beta, (__nv_bfloat16*)out_bf16, P, C, eps, total4);
    else gn_apply_kernel<false><<<grid_for(total4), 256, 0, s>>>(x, sums, gamma, beta, (__nv_bfloat16*)out_bf16, P, C, eps, total4);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_upsample2x_nhwc_bf16(const float* x, void* out_bf16, int B, int H, int W, int C, void* stream) {
    if (!x || !out_bf16 || C % 4) return kBadArgument;
    const int64_t total4 = (int64_t)B * 4 * H * W * C / 4;
    upsample2x_kernel<<<grid_for(total4), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, H, W, C, total4);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_cast_f32_bf16(const float* x, void* out_bf16, int64_t n, void* stream) {
    if (!x || !out_bf16 || n <= 0 || n % 4) return kBadArgument;
    cast_bf16_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, n / 4);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_softmax_rows_bf16(const float* x, void* out_bf16, int R, int n, float scale, void* stream) {
    if (!x || !out_bf16 || R <= 0 || n <= 0) return kBadArgument;
    softmax_rows_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, R, n, scale);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_nhwc_to_nchw_f32(const float* x, float* out, int B, int P, int C, void* stream) {
    if (!x || !out) return kBadArgument;
    const int64_t total = (int64_t)B * P * C;
    nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, out, P, C, total);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_image_to_uint8(const float* x, uint8_t* out, int64_t n, void* stream) {
    if (!x || !out || n <= 0) return kBadArgument;
    image_to_uint8_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, out, n);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_image_to_nhwc64_bf16(const float* pixels_nchw, void* out_bf16, int B, int H, int W, void* stream) {
    if (!pixels_nchw || !out_bf16 || B <= 0 || H <= 0 || W <= 0) return kBadArgument;
    const int64_t total = (int64_t)B * H * W * 8;
    image_to_nhwc64_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(pixels_nchw, (__nv_bfloat16*)out_bf16, H * W, total);
    return cuda_status(cudaGetLastError());
}
extern "C" int mmada_space_to_depth2_bf16(const float* x, void* out_bf16, int B, int H, int W, int C, void* stream) {
    if (!x || !out_bf16 || B <= 0 || H <= 0 || W <= 0 || C <= 0) return kBadArgument;
    if ((H & 1) || (W & 1) || (C & 3)) return kUnsupportedShape;
    const int64_t total4 = (int64_t)B * H * W * C / 4;
    space_to_depth2_kernel<<<grid_for(total4), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, H, W, C, total4);
    return cuda_status(cudaGetLastError());
}
