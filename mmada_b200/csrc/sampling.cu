// Fused sampling step of MMadaModelLM.t2i_generate and the MaskGIT re-masking rule.
//
// One launch replaces the ~25 eager PyTorch kernels of /root/reference/models/modeling_mmada.py:164-209
// and /root/reference/models/sampling.py:31-36 for one denoising step:
//   CFG mix (1+g)*cond - g*uncond        :167   (mul, mul, sub — no FMA contraction, Appendix A Q5)
//   softmax over the codebook            :176
//   multinomial == argmax(p / q)         :179   (q ~ Exp(1) supplied by the caller, Q6)
//   known-token override, p[sampled]     :183-193
//   mask_len clamp                       :195-200
//   log p + T * gumbel(u), k-th smallest, strict '<'   sampling.py:31-36 (Q7)
//   write-back into input_ids / the running code ids   :206-209
//
// Layout: one CTA per (batch row, image position); logits are read once with 128-bit coalesced
// loads and held in registers across the max / sum / argmax passes; reductions are warp shuffles.
// Positions whose token is already known skip their 3 x 32 KiB of logits and noise entirely.  The
// last CTA to finish a batch row (atomic ticket) sorts that row's N confidences in shared memory
// (bitonic) and applies the cut-off, so the whole step is a single kernel.
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

constexpr int MAX_TOKENS = 4096;   // image positions per row supported by the in-smem sort

struct T2ISampleParams {
    const float* cond;      // [B*N, C]
    const float* uncond;    // [B*N, C] or nullptr
    const float* q;         // [B*N, C]  Exp(1) noise
    const float* u;         // [B, N]    U(0,1) noise
    int64_t* known;         // [B, N]    code id or mask_id (in/out)
    int64_t* input_ids;     // [B, ld_ids] (in/out), image tokens at columns [img_off, img_off+N)
    int64_t* sampled_out;   // [B, N]
    float* sel_out;         // [B, N]    selected probs (finfo.max at known positions)
    uint8_t* masking_out;   // [B, N]    optional
    int64_t* raw_out;       // [B, N]    optional: the raw argmax(p/q) at EVERY position (disables the known-row skip)
    int no_remask;          // 1: commit the merged tokens and skip the re-masking (t2m_generate's last step)
    int32_t* tickets;       // [B]       zero on entry, zero on exit
    const int32_t* slot;    // [B, N]    optional: row of cond / uncond that holds position (b, n)'s logits (compact logits:
                            //           only the still-masked positions were given to the output head); nullptr = b*N + n
    int64_t ld_ids;
    int64_t img_off;
    int64_t mask_id;
    int64_t text_vocab;
    int B, N, C;
    float one_plus_g, g;
    float mask_len_raw;
    float temperature;
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_add(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float confidence_of(float p, float u, float temperature) {
    // log(clamp(p)) + T * (-log(clamp(-log(clamp(u)))))   with eps = 1e-20, each op rounded to fp32
    const float eps = 1e-20f;
    const float lp = logf(fmaxf(p, eps));
    const float g = -logf(fmaxf(-logf(fmaxf(u, eps)), eps));
    return __fadd_rn(lp, __fmul_rn(temperature, g));
}

// ascending bitonic sort of s[0..n_pad) (n_pad a power of two), all threads of the block
__device__ void bitonic_sort(float* s, int n_pad) {
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = s[i], b = s[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s[i] = b; s[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// re-masking of one batch row by the whole block.  sel/u/known/... point at the row.
__device__ void remask_row(const float* sel, const float* u, int64_t* known, int64_t* ids_row, const int64_t* sampled,
                           uint8_t* masking_out, int N, float mask_len_raw, float temperature, int64_t mask_id,
                           int64_t text_vocab, float* s_conf, int* s_int) {
    int n_pad = 1;
    while (n_pad < N) n_pad <<= 1;
    int unknown = 0;
    for (int n = threadIdx.x; n < n_pad; n += blockDim.x) {
        float c = INFINITY;
        if (n < N) {
            c = confidence_of(__ldcg(sel + n), u[n], temperature);
            unknown += (known[n] == mask_id);
        }
        s_conf[n] = c;
    }
    unknown = __reduce_add_sync(0xffffffffu, unknown);
    if (threadIdx.x == 0) *s_int = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicAdd(s_int, unknown);
    __syncthreads();
    unknown = *s_int;
    bitonic_sort(s_conf, n_pad);
    // mask_len = max(1, min(unknown - 1, raw)) evaluated in fp32 like the (B,1) float tensor, then .long()
    float ml = fmaxf(1.0f, fminf((float)(unknown - 1), mask_len_raw));
    int k = (int)ml;
    k = k < 0 ? 0 : (k > N - 1 ? N - 1 : k);
    const float cut = s_conf[k];
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float c = confidence_of(__ldcg(sel + n), u[n], temperature);
        const bool m = c < cut;
        const int64_t tok = __ldcg(sampled + n);
        if (ids_row) ids_row[n] = m ? mask_id : tok + text_vocab;
        known[n] = m ? mask_id : tok;
        if (masking_out) masking_out[n] = m ? 1 : 0;
    }
}

// THREADS x VEC 128-bit loads cover a row of C logits; MINB resident CTAs per SM bound the registers, CH = loads per
// array in flight per thread.  C = 8192: 4 CTAs x 256 threads at 64 registers (0.169 ms for 8 x 1024 rows against 0.214 ms
// at 2 CTAs of 112 registers: the kernel alternates load and reduce phases, so it wants resident CTAs, not registers).
template <int THREADS, int VEC, int MINB = 1, int CH = VEC>
__global__ void __launch_bounds__(THREADS, MINB) t2i_sample_kernel(const T2ISampleParams p) {
    __shared__ float s_red[32];
    __shared__ float s_sum[32];          // (its own array: no barrier is needed between the two reductions' reads and writes)
    __shared__ int s_idx[32];
    __shared__ int s_flag;
    __shared__ int s_cnt;
    __shared__ float s_conf[MAX_TOKENS];
    const int row = blockIdx.x;
    const int b = row / p.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    // (the slot of the position's logits is read next to its token, not behind the branch on it: one dependent global
    // load less in front of the logit loads)
    const int32_t slot_of_row = p.slot ? __ldg(p.slot + row) : 0;
    const int64_t known = p.known[row];

    if (known != p.mask_id && p.raw_out == nullptr) {
        if (tid == 0) {
            p.sampled_out[row] = known;
            p.sel_out[row] = FLT_MAX;
        }
    } else {
        const int64_t lrow = p.slot ? max(slot_of_row, 0) : row;     // (a masked position always has a slot: cap bounds them)
        const float4* c4 = reinterpret_cast<const float4*>(p.cond + lrow * p.C);
        const float4* q4 = reinterpret_cast<const float4*>(p.q + (int64_t)row * p.C);
        float l[VEC][4];
        // ---- CFG mix, CH 128-bit loads per array in flight per thread (register budget: the resident CTAs hide the rest)
        if (p.uncond) {
            const float4* u4 = reinterpret_cast<const float4*>(p.uncond + lrow * p.C);
#pragma unroll
            for (int i0 = 0; i0 < VEC; i0 += CH) {
                float4 a[CH], bb[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) a[i] = __ldcs(c4 + (i0 + i) * THREADS + tid);
#pragma unroll
                for (int i = 0; i < CH; ++i) bb[i] = __ldcs(u4 + (i0 + i) * THREADS + tid);
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    l[i0 + i][0] = __fsub_rn(__fmul_rn(p.one_plus_g, a[i].x), __fmul_rn(p.g, bb[i].x));
                    l[i0 + i][1] = __fsub_rn(__fmul_rn(p.one_plus_g, a[i].y), __fmul_rn(p.g, bb[i].y));
                    l[i0 + i][2] = __fsub_rn(__fmul_rn(p.one_plus_g, a[i].z), __fmul_rn(p.g, bb[i].z));
                    l[i0 + i][3] = __fsub_rn(__fmul_rn(p.one_plus_g, a[i].w), __fmul_rn(p.g, bb[i].w));
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float4 a = __ldcs(c4 + i * THREADS + tid);
                l[i][0] = a.x; l[i][1] = a.y; l[i][2] = a.z; l[i][3] = a.w;
            }
        }
        // ---- row max
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) mx = fmaxf(mx, l[i][j]);
        mx = warp_max(mx);
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        mx = s_red[lane < NW ? lane : 0];
        mx = warp_max(mx);
        // ---- exp and sum
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                l[i][j] = expf(__fsub_rn(l[i][j], mx));
                sum += l[i][j];
            }
        sum = warp_add(sum);
        if (lane == 0) s_sum[warp] = sum;
        __syncthreads();                 // (also orders every thread's read of s_red above before the argmax pass rewrites it)
        sum = lane < NW ? s_sum[lane] : 0.f;
        sum = warp_add(sum);
        // ---- argmax of p / q, first index wins ties.  The value that decides is r = fl(fl(e / sum) / q): two IEEE divisions
        // (reciprocal, Newton steps, range check with an out-of-line slow path — 40 % of the kernel's instructions when every
        // logit takes them).  Here a logit only takes them when an UPPER bound of its r, three instructions, does not stay
        // below a value some logit of this warp has already reached (rlb, refreshed across the warp whenever a lane
        // improves): the expected number of such steps per warp is the harmonic number H_32 = 4 of 32.  Nothing changes in
        // the result: a skipped logit has r < rlb <= the row's maximum, strictly, so it can neither win nor tie.
        //   ub >= r:  inv_s >= (1 / sum)(1 + 2^-19) (reciprocal and product rounded up), rcp.approx is within 2^-23 of 1 / q,
        //   both products are rounded up; r itself carries two roundings of 2^-24.  NaN compares as a candidate.
        float best = -INFINITY, best_p = 0.f;
        int best_i = 0x7fffffff;
        float rlb = -INFINITY;
        const float inv_s = __fmul_ru(__frcp_ru(sum), 1.0000019073486328f);
        // the noise is read CH 128-bit loads at a time, only now: the other resident CTAs cover the latency
#pragma unroll
        for (int i0 = 0; i0 < VEC; i0 += CH) {
            float4 qq[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) qq[i] = __ldcs(q4 + (i0 + i) * THREADS + tid);
#pragma unroll
            for (int ii = 0; ii < CH; ++ii) {
                const int i = i0 + ii;
                const float qv[4] = {qq[ii].x, qq[ii].y, qq[ii].z, qq[ii].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float rq;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(qv[j]));
                    const float ub = __fmul_ru(__fmul_ru(l[i][j], inv_s), rq);
                    const bool cand = !(ub < rlb);
                    if (__any_sync(0xffffffffu, cand)) {
                        if (cand) {
                            const float pr = __fdiv_rn(l[i][j], sum);
                            const float r = __fdiv_rn(pr, qv[j]);
                            if (r > best) { best = r; best_p = pr; best_i = (i * THREADS + tid) * 4 + j; }
                        }
                        float m = best;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                        rlb = m;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const float op = __shfl_xor_sync(0xffffffffu, best_p, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_p = op; best_i = oi; }
        }
        if (lane == 0) { s_red[warp] = best; s_idx[warp] = best_i; s_conf[warp] = best_p; }
        __syncthreads();
        if (warp == 0) {
            best = lane < NW ? s_red[lane] : -INFINITY;
            best_i = lane < NW ? s_idx[lane] : 0x7fffffff;
            best_p = lane < NW ? s_conf[lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const float op = __shfl_xor_sync(0xffffffffu, best_p, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                if (ob > best || (ob == best && oi < best_i)) { best = ob; best_p = op; best_i = oi; }
            }
            if (lane == 0) {
                const int64_t tok = best_i == 0x7fffffff ? 0 : best_i;
                if (p.raw_out) p.raw_out[row] = tok;
                p.sampled_out[row] = known != p.mask_id ? known : tok;
                p.sel_out[row] = known != p.mask_id ? FLT_MAX : best_p;
            }
        }
    }
    if (p.no_remask) {
        __syncthreads();
        if (tid == 0) {
            const int64_t tok = p.sampled_out[row];
            if (p.input_ids) p.input_ids[(int64_t)b * p.ld_ids + p.img_off + (row - b * p.N)] = tok + p.text_vocab;
            p.known[row] = tok;
            if (p.masking_out) p.masking_out[row] = 0;
        }
        return;
    }
    // ---- ticket: the last CTA of batch row b re-masks it
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const int old = atomicAdd(p.tickets + b, 1);
        s_flag = (old == p.N - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    remask_row(p.sel_out + (int64_t)b * p.N, p.u + (int64_t)b * p.N, p.known + (int64_t)b * p.N,
               p.input_ids ? p.input_ids + (int64_t)b * p.ld_ids + p.img_off : nullptr,
               p.sampled_out + (int64_t)b * p.N, p.masking_out ? p.masking_out + (int64_t)b * p.N : nullptr, p.N,
               p.mask_len_raw, p.temperature, p.mask_id, p.text_vocab, s_conf, &s_cnt);
    if (tid == 0) p.tickets[b] = 0;
}

// Rows of the still-masked image positions, for the output head restricted to them (north_star; the reference
// computes and discards the logits of known positions, modeling_mmada.py:183-184).  One CTA per batch row b:
//   rows[(r*B + b)*cap + j] = (r*B + b)*L + img_off + n_j   for the j-th masked position n_j of row b, r = 0 (cond) and,
//                             with `branches` = 2, r = 1 (uncond); slots j >= the number of masked positions repeat the
//                             row's first image position (computed and never read)
//   slot[b*N + n]           = b*cap + j for masked positions, -1 for known ones
// `cap` is the caller's upper bound on the masked positions per row (the previous step's mask_len).
__global__ void __launch_bounds__(256) compact_masked_kernel(const int64_t* __restrict__ known, int32_t* __restrict__ rows,
                                                             int32_t* __restrict__ slot, int B, int N, int L, int img_off,
                                                             int cap, int branches, int64_t mask_id) {
    __shared__ int s_cnt[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int per = (N + 255) / 256;
    const int n0 = tid * per, n1 = min(N, n0 + per);
    int c = 0;
    for (int n = n0; n < n1; ++n) c += known[(int64_t)b * N + n] == mask_id;
    s_cnt[tid] = c;
    __syncthreads();
    if (tid == 0) {                 // exclusive scan of 256 counts
        int acc = 0;
        for (int i = 0; i < 256; ++i) { const int v = s_cnt[i]; s_cnt[i] = acc; acc += v; }
    }
    __syncthreads();
    int j = s_cnt[tid];
    for (int n = n0; n < n1; ++n) {
        const bool m = known[(int64_t)b * N + n] == mask_id;
        int sl = -1;
        if (m && j < cap) {
            sl = b * cap + j;
            for (int r = 0; r < branches; ++r) rows[((int64_t)r * B + b) * cap + j] = (r * B + b) * L + img_off + n;
        }
        slot[(int64_t)b * N + n] = sl;
        j += m;
    }
    __syncthreads();
    // total masked = scan value of the last thread + its count; pad the unused slots
    __shared__ int s_total;
    if (tid == 255) s_total = j;
    __syncthreads();
    for (int k = s_total + tid; k < cap; k += 256)
        for (int r = 0; r < branches; ++r) rows[((int64_t)r * B + b) * cap + k] = (r * B + b) * L + img_off;
}

// standalone mask_by_random_topk: masking[b, n] = conf[b, n] < sorted(conf[b])[mask_len[b]]
__global__ void __launch_bounds__(256) random_topk_kernel(const float* __restrict__ probs, const float* __restrict__ u,
                                                          const int64_t* __restrict__ mask_len, uint8_t* __restrict__ out,
                                                          int N, float temperature) {
    __shared__ float s_conf[MAX_TOKENS];
    const int b = blockIdx.x;
    const float* pr = probs + (int64_t)b * N;
    const float* ur = u + (int64_t)b * N;
    int n_pad = 1;
    while (n_pad < N) n_pad <<= 1;
    for (int n = threadIdx.x; n < n_pad; n += blockDim.x)
        s_conf[n] = n < N ? confidence_of(pr[n], ur[n], temperature) : INFINITY;
    __syncthreads();
    bitonic_sort(s_conf, n_pad);
    int64_t k = mask_len[b];
    k = k < 0 ? k + N : k;                       // torch.gather accepts no negatives; kept for safety
    k = k < 0 ? 0 : (k > N - 1 ? N - 1 : k);
    const float cut = s_conf[k];
    for (int n = threadIdx.x; n < N; n += blockDim.x)
        out[(int64_t)b * N + n] = confidence_of(pr[n], ur[n], temperature) < cut ? 1 : 0;
}

}  // namespace mmada

using namespace mmada;

extern "C" int mmada_compact_masked_rows(const int64_t* known_ids, int32_t* rows_out, int32_t* slot_out, int B, int N,
                                         int L, int img_off, int cap, int branches, int64_t mask_id, void* stream) {
    if (!known_ids || !rows_out || !slot_out) return kBadArgument;
    if (B <= 0 || N <= 0 || cap <= 0 || cap > N || img_off < 0 || img_off + N > L || branches < 1 || branches > 2)
        return kBadArgument;
    if ((int64_t)branches * B * L > 0x7fffffffLL) return kUnsupportedShape;
    compact_masked_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(known_ids, rows_out, slot_out, B, N, L, img_off, cap,
                                                               branches, mask_id);
    return cuda_status(cudaGetLastError());
}

static int t2i_sample_step_impl(const float* cond_logits, const float* uncond_logits, const float* q_noise,
                                const float* u_noise, int64_t* known_ids, int64_t* input_ids, int64_t ld_ids,
                                int64_t img_off, int64_t* sampled_out, float* sel_out, uint8_t* masking_out,
                                int64_t* raw_out, int no_remask, int32_t* tickets, int B, int N, int C,
                                float one_plus_g, float g, float mask_len_raw, float temperature,
                                int64_t mask_id, int64_t text_vocab, const int32_t* logit_slot, void* stream) {
    if (logit_slot && raw_out) return kBadArgument;        // raw samples need the logits of every position
    if (!cond_logits || !q_noise || !u_noise || !known_ids || !sampled_out || !sel_out || !tickets) return kBadArgument;
    if (B <= 0 || N <= 0 || N > MAX_TOKENS) return kUnsupportedShape;
    if ((reinterpret_cast<uintptr_t>(cond_logits) | reinterpret_cast<uintptr_t>(uncond_logits) |
         reinterpret_cast<uintptr_t>(q_noise)) & 15)
        return kBadArgument;
    T2ISampleParams p;
    p.cond = cond_logits; p.uncond = uncond_logits; p.q = q_noise; p.u = u_noise;
    p.known = known_ids; p.input_ids = input_ids; p.sampled_out = sampled_out; p.sel_out = sel_out;
    p.masking_out = masking_out; p.raw_out = raw_out; p.no_remask = no_remask; p.tickets = tickets; p.ld_ids = ld_ids; p.img_off = img_off;
    p.slot = logit_slot;
    p.mask_id = mask_id; p.text_vocab = text_vocab; p.B = B; p.N = N; p.C = C;
    p.one_plus_g = one_plus_g; p.g = g; p.mask_len_raw = mask_len_raw; p.temperature = temperature;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = B * N;
    switch (C) {
        case 8192: t2i_sample_kernel<256, 8, 4, 4><<<grid, 256, 0, s>>>(p); break;
        case 4096: t2i_sample_kernel<256, 4, 4, 4><<<grid, 256, 0, s>>>(p); break;
        case 2048: t2i_sample_kernel<256, 2><<<grid, 256, 0, s>>>(p); break;
        case 1024: t2i_sample_kernel<256, 1><<<grid, 256, 0, s>>>(p); break;
        case 512: t2i_sample_kernel<128, 1><<<grid, 128, 0, s>>>(p); break;
        default: return kUnsupportedShape;
    }
    return cuda_status(cudaGetLastError());
}

extern "C" int mmada_t2i_sample_step(const float* cond_logits, const float* uncond_logits, const float* q_noise,
                                     const float* u_noise, int64_t* known_ids, int64_t* input_ids, int64_t ld_ids,
                                     int64_t img_off, int64_t* sampled_out, float* sel_out, uint8_t* masking_out,
                                     int64_t* raw_out, int no_remask, int32_t* tickets, int B, int N, int C,
                                     float one_plus_g, float g, float mask_len_raw, float temperature,
                                     int64_t mask_id, int64_t text_vocab, void* stream) {
    return t2i_sample_step_impl(cond_logits, uncond_logits, q_noise, u_noise, known_ids, input_ids, ld_ids, img_off,
                                sampled_out, sel_out, masking_out, raw_out, no_remask, tickets, B, N, C, one_plus_g, g,
                                mask_len_raw, temperature, mask_id, text_vocab, nullptr, stream);
}

extern "C" int mmada_t2i_sample_step_compact(const float* cond_logits, const float* uncond_logits, const float* q_noise,
                                             const float* u_noise, int64_t* known_ids, int64_t* input_ids,
                                             int64_t ld_ids, int64_t img_off, int64_t* sampled_out, float* sel_out,
                                             uint8_t* masking_out, int no_remask, int32_t* tickets, int B, int N, int C,
                                             float one_plus_g, float g, float mask_len_raw, float temperature,
                                             int64_t mask_id, int64_t text_vocab, const int32_t* logit_slot,
                                             void* stream) {
    if (!logit_slot) return kBadArgument;
    return t2i_sample_step_impl(cond_logits, uncond_logits, q_noise, u_noise, known_ids, input_ids, ld_ids, img_off,
                                sampled_out, sel_out, masking_out, nullptr, no_remask, tickets, B, N, C, one_plus_g, g,
                                mask_len_raw, temperature, mask_id, text_vocab, logit_slot, stream);
}

extern "C" int mmada_mask_by_random_topk(const float* probs, const float* u_noise, const int64_t* mask_len,
                                         uint8_t* masking_out, int B, int N, float temperature, void* stream) {
    if (!probs || !u_noise || !mask_len || !masking_out) return kBadArgument;
    if (B <= 0 || N <= 0 || N > MAX_TOKENS) return kUnsupportedShape;
    random_topk_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(probs, u_noise, mask_len, masking_out, N, temperature);
    return cuda_status(cudaGetLastError());
}
