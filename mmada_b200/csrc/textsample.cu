// Low-confidence-remasking sampling for the semi-autoregressive text / MMU path.
//
// Replaces, per denoising step of generate() (/root/reference/generate.py:79-111, identical to
// MMadaModelLM.mmu_generate, models/modeling_mmada.py:425-478):
//   CFG mix  un + (cfg+1)*(l - un)                         generate.py:86
//   add_gumbel_noise: exp(l64) / (-log u64)^T, argmax      generate.py:8-19,90-91   (fp64, Q12)
//   softmax(l64)[x0]                                       generate.py:93-96        (fp64)
//   block-end / unmasked positions -> -inf, per-row top-k, write-back   generate.py:102-111 (Q13)
//   get_num_transfer_tokens                                generate.py:22-40
// The reference evaluates the fp64 Gumbel/softmax chain over the whole (B, L, V) tensor; only the
// still-masked positions of the current block can be selected (Appendix A, Q21), so the kernels run
// on the candidate rows only: one CTA per (sequence, block position), logits read with 128-bit
// coalesced loads, reductions by warp shuffles.
#include <math.h>

#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

// ---- Philox4x32-10 (counter-based; one call yields two uniform doubles) -------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u64_to_unit_double(uint32_t hi, uint32_t lo) {
    const unsigned long long x = ((unsigned long long)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0);   // (0, 1)
}

struct ArgBest {
    double v;
    int i;
};
__device__ __forceinline__ void best_merge(ArgBest& a, double v, int i) {
    if (v > a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; }
}

constexpr int TS_THREADS = 512;

// one CTA per candidate row
__global__ void __launch_bounds__(TS_THREADS)
text_sample_kernel(const float* __restrict__ logits, const float* __restrict__ un_logits, float cfg_plus1,
                   const double* __restrict__ u_noise, unsigned long long seed, float temperature, int V,
                   int64_t* __restrict__ x0_out, double* __restrict__ conf_out) {
    __shared__ double s_v[TS_THREADS / 32];
    __shared__ int s_i[TS_THREADS / 32];
    __shared__ float s_m[TS_THREADS / 32];
    __shared__ double s_bcast[2];
    __shared__ int s_bi;
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* lr = logits + (int64_t)row * V;
    const float* ur = un_logits ? un_logits + (int64_t)row * V : nullptr;
    const double* nr = u_noise ? u_noise + (int64_t)row * V : nullptr;
    const double T = (double)temperature;
    auto mixed = [&](float l, float un) { return __fadd_rn(un, __fmul_rn(cfg_plus1, __fsub_rn(l, un))); };

    // ---- pass 1: row max (exact in fp32) and the Gumbel-max token
    ArgBest best{-INFINITY, 0x7fffffff};
    float mx = -INFINITY;
    const int V4 = V >> 2;
    for (int c = tid; c < V4; c += TS_THREADS) {
        float4 l4 = __ldg(reinterpret_cast<const float4*>(lr) + c);
        if (ur) {
            const float4 n4 = __ldg(reinterpret_cast<const float4*>(ur) + c);
            l4.x = mixed(l4.x, n4.x); l4.y = mixed(l4.y, n4.y); l4.z = mixed(l4.z, n4.z); l4.w = mixed(l4.w, n4.w);
        }
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
        double uu[4];
        if (temperature != 0.f) {
            if (nr) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(nr) + 2 * c);
                const double2 b = __ldg(reinterpret_cast<const double2*>(nr) + 2 * c + 1);
                uu[0] = a.x; uu[1] = a.y; uu[2] = b.x; uu[3] = b.y;
            } else {
                uint32_t r0[4], r1[4];
                philox4x32_10((uint32_t)c, (uint32_t)row, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r0);
                philox4x32_10((uint32_t)c, (uint32_t)row, 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r1);
                uu[0] = u64_to_unit_double(r0[0], r0[1]); uu[1] = u64_to_unit_double(r0[2], r0[3]);
                uu[2] = u64_to_unit_double(r1[0], r1[1]); uu[3] = u64_to_unit_double(r1[2], r1[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mx = fmaxf(mx, lv[j]);
            double v;
            if (temperature != 0.f) v = exp((double)lv[j]) / pow(-log(uu[j]), T);
            else v = (double)lv[j];
            best_merge(best, v, 4 * c + j);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
        best_merge(best, ov, oi);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_v[warp] = best.v; s_i[warp] = best.i; s_m[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        best.v = lane < TS_THREADS / 32 ? s_v[lane] : -INFINITY;
        best.i = lane < TS_THREADS / 32 ? s_i[lane] : 0x7fffffff;
        mx = lane < TS_THREADS / 32 ? s_m[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best.v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
            best_merge(best, ov, oi);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) { s_bi = best.i; s_bcast[0] = (double)mx; }
    }
    __syncthreads();
    // a row with no comparable value (every logit NaN) leaves the sentinel: commit token 0 with a NaN confidence, like
    // torch.argmax over an all-NaN row, instead of indexing the row with 0x7fffffff
    const int x0 = s_bi == 0x7fffffff ? 0 : s_bi;
    const double m = s_bcast[0];
    // ---- pass 2: fp64 softmax denominator (the row is L2-resident from pass 1)
    double sum = 0.0;
    for (int c = tid; c < V4; c += TS_THREADS) {
        float4 l4 = __ldg(reinterpret_cast<const float4*>(lr) + c);
        if (ur) {
            const float4 n4 = __ldg(reinterpret_cast<const float4*>(ur) + c);
            l4.x = mixed(l4.x, n4.x); l4.y = mixed(l4.y, n4.y); l4.z = mixed(l4.z, n4.z); l4.w = mixed(l4.w, n4.w);
        }
        sum += exp((double)l4.x - m) + exp((double)l4.y - m) + exp((double)l4.z - m) + exp((double)l4.w - m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) s_v[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < TS_THREADS / 32; ++w) tot += s_v[w];
        float l0 = lr[x0];
        if (ur) l0 = mixed(l0, ur[x0]);
        x0_out[row] = x0;
        conf_out[row] = exp((double)l0 - m) / tot;
    }
}

// cnt[b] = number of masked positions of sequence b inside the block (start of a block)
__global__ void block_mask_count_kernel(const int64_t* __restrict__ x, int64_t ld, int lo, int block, int64_t mask_id,
                                        int32_t* __restrict__ cnt) {
    const int b = blockIdx.x;
    int c = 0;
    for (int p = threadIdx.x; p < block; p += blockDim.x) c += x[(int64_t)b * ld + lo + p] == mask_id;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) cnt[b] = s;
}

// per sequence: pick the k most confident masked positions of the block and commit their tokens.
// k = cnt/steps + (step < cnt%steps)  (get_num_transfer_tokens); ties -> lower position first.
constexpr int MAX_BLOCK = 2048;
__global__ void __launch_bounds__(256)
text_transfer_kernel(int64_t* __restrict__ x, int64_t ld, int lo, int block, const int64_t* __restrict__ x0,
                     const double* __restrict__ conf, const double* __restrict__ conf_override,
                     const int32_t* __restrict__ cnt, int steps, int step, int64_t mask_id,
                     uint8_t* __restrict__ transfer_out) {
    __shared__ double s_c[MAX_BLOCK];
    const int b = blockIdx.x;
    int64_t* xr = x + (int64_t)b * ld + lo;
    for (int p = threadIdx.x; p < block; p += blockDim.x) {
        const bool masked = xr[p] == mask_id;
        const double c = conf_override ? conf_override[(int64_t)b * block + p] : conf[(int64_t)b * block + p];
        s_c[p] = masked ? c : -INFINITY;
    }
    __syncthreads();
    const int n = cnt[b];
    const int k = n / steps + (step < n % steps ? 1 : 0);
    for (int p = threadIdx.x; p < block; p += blockDim.x) {
        const double c = s_c[p];
        bool sel = false;
        if (c > -INFINITY) {
            int rank = 0;
            for (int q = 0; q < block; ++q) {
                const double cq = s_c[q];
                rank += (cq > c) || (cq == c && q < p);
            }
            sel = rank < k;
        }
        if (transfer_out) transfer_out[(int64_t)b * block + p] = sel;
        if (sel) xr[p] = x0[(int64_t)b * block + p];
    }
}

}  // namespace mmada

using namespace mmada;

extern "C" int mmada_text_sample_rows(const float* logits, const float* un_logits, float cfg_scale_plus1,
                                      const double* u_noise, uint64_t seed, float temperature, int R, int V,
                                      int64_t* x0_out, double* conf_out, void* stream) {
    if (!logits || !x0_out || !conf_out || R <= 0 || V <= 0) return kBadArgument;
    if (V % 4) return kUnsupportedShape;
    if ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(un_logits) | reinterpret_cast<uintptr_t>(u_noise)) & 15)
        return kBadArgument;
    text_sample_kernel<<<R, TS_THREADS, 0, (cudaStream_t)stream>>>(logits, un_logits, cfg_scale_plus1, u_noise,
                                                                  (unsigned long long)seed, temperature, V, x0_out, conf_out);
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_block_mask_count(const int64_t* x, int64_t ld, int lo, int block, int B, int64_t mask_id,
                                      int32_t* cnt_out, void* stream) {
    if (!x || !cnt_out || B <= 0 || block <= 0) return kBadArgument;
    block_mask_count_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(x, ld, lo, block, mask_id, cnt_out);
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_text_transfer(int64_t* x, int64_t ld, int lo, int block, const int64_t* x0, const double* conf,
                                   const double* conf_override, const int32_t* cnt, int steps, int step, int B,
                                   int64_t mask_id, uint8_t* transfer_out, void* stream) {
    if (!x || !x0 || (!conf && !conf_override) || !cnt || B <= 0 || steps <= 0) return kBadArgument;
    if (block <= 0 || block > MAX_BLOCK) return kUnsupportedShape;
    text_transfer_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, ld, lo, block, x0, conf, conf_override, cnt, steps, step,
                                                              mask_id, transfer_out);
    return cuda_status(cudaGetLastError());
}
