// Motion VQ-VAE decoder helpers (text-to-motion, BASELINE config 5): the 1-D convolutions of
// /root/reference/motion_vqvae/models/encdec.py:35-67 and resnet.py:12-81 run as tcgen05 GEMMs (gemm.cu) over a
// gathered operand: for every output frame the taps' input frames are laid side by side, [t - dil | t | t + dil]
// x C channels, bf16, zero outside the sequence.  The gather also applies the ReLU in front of the convolution
// (ResConv1DBlock: activation -> conv) and nn.Upsample(scale_factor=2, mode='nearest') (Decoder blocks), so neither
// is a pass of its own.  Sequences of a batch are independent: no tap crosses a sequence boundary.
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

// x fp32 [B, T_in, C] -> out bf16 [B, T_in * up, taps * C]
__global__ void conv1d_gather_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int T_in, int C, int taps,
                                     int dil, int up, int relu, int64_t total4) {
    const int C4 = C >> 2;
    const int T_out = T_in * up;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        int64_t r = i / C4;
        const int k = (int)(r % taps); r /= taps;
        const int t = (int)(r % T_out);
        const int64_t b = r / T_out;
        const int tu = t + (k - taps / 2) * dil;                 // position in the (upsampled) input sequence
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tu >= 0 && tu < T_out) {
            v = *reinterpret_cast<const float4*>(x + ((b * T_in + tu / up) * (int64_t)C + 4 * c4));
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        }
        reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

__global__ void relu_f32_kernel(float* __restrict__ x, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(x)[i];
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        reinterpret_cast<float4*>(x)[i] = v;
    }
}

static inline int grid_for(int64_t work, int threads = 256) {
    int64_t blocks = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace mmada

using namespace mmada;

extern "C" int mmada_conv1d_gather_bf16(const float* x, void* out_bf16, int B, int T_in, int C, int taps, int dilation,
                                        int upsample, int relu, void* stream) {
    if (!x || !out_bf16 || B <= 0 || T_in <= 0 || C <= 0) return kBadArgument;
    if ((taps != 1 && taps != 3) || dilation < 1 || (upsample != 1 && upsample != 2)) return kBadArgument;
    if (C & 3) return kUnsupportedShape;
    const int64_t total4 = (int64_t)B * T_in * upsample * taps * (C / 4);
    conv1d_gather_kernel<<<grid_for(total4), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, T_in, C, taps,
                                                                              dilation, upsample, relu, total4);
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_relu_f32(float* x, int64_t n, void* stream) {
    if (!x || n <= 0 || (n & 3)) return kBadArgument;
    relu_f32_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(x, n / 4);
    return cuda_status(cudaGetLastError());
}
