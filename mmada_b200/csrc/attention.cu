// Bidirectional (non-causal, unmasked, no KV cache) flash attention on tcgen05 / TMEM.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False) at
// /root/reference/models/modeling_llada.py:653-660 (called from :711-718; the attention_bias the
// reference builds is never applied — SURVEY.md Appendix A, Q1).
//
// q, k, v are read in place from the fused projection output [B*L, ld] (head h at columns h*hd),
// the output is written token-major [B*L, ldo] ready for the attn_out GEMM: no transposes.
//
// One CTA handles one (batch, head) and up to two 128-row query tiles.  Keys are consumed in SUB-TILES of
// 64: every query tile owns two 64-column score buffers in TMEM, so the tensor pipe computes the scores
// of sub-tile t+1 (and t+2) while the softmax threads still work on sub-tile t.  With 128-key score tiles
// (one buffer per query tile) the chain  S -> softmax -> P -> PV -> next S  was serial per query tile and
// the kernel sat at 40 % tensor-pipe (profiles/r01_attention_*); the double buffer makes it a throughput
// problem (MUFU / issue slots / tensor pipe) instead of a latency problem.
//   warps 0-3 / 4-7  softmax for query tile 0 / 1, ONE thread per query row (no cross-thread exchange):
//                    tcgen05.ld the 64 scores, online softmax in base 2 with lazy rescaling of the output
//                    accumulator (only when the running max grows by more than 2^8), P (bf16) written back
//                    over the first 32 columns of the score buffer.  A compile-time share of the
//                    exponentials runs as a polynomial on the FMA pipe (the MUFU pipe, 16 ex2/clk/SM, needs
//                    as long for a 128x128 tile as the tensor pipe needs for its two MMAs).
//   warp 8           TMA producer: Q once, then K/V tiles of 128 keys through 2-stage rings
//   warp 9           MMA issuer: S_i(t) = Q_i K_t^T (both operands K-major in shared memory, N = 64) and
//                    O_i += P_i(t) V_t (P from TMEM, V as an MN-major shared-memory operand)
// TMEM: S0a S0b | S1a S1b | O0 | O1  (4 x 64 + 2 x hd columns).  The last sub-tile is shortened to a
// multiple of 16 keys.
#include <math.h>
#include <stdlib.h>

#include "attn_math.cuh"
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

constexpr int ATT_THREADS = 320;   // 8 softmax warps + TMA warp + MMA warp
constexpr int QT = 128;    // query rows per tile
constexpr int KT = 128;    // keys per TMA tile
constexpr int KS = 64;     // keys per score sub-tile
constexpr int TMA_WARP = 8, MMA_WARP = 9;
constexpr int kAttnPolyDefault = 2;

struct AttnParams {
    __nv_bfloat16* out;
    int64_t ldo;
    int L, H, B;
    int q_begin;           // first query row this launch handles (rows before it belong to another launch)
    float scale_log2;      // softmax scale * log2(e)
#ifdef MMADA_ATT_TRACE
    long long* trace;      // debug build only: clock64 timeline of CTA 0 (scripts/attn_trace.py)
#endif
};

#ifdef MMADA_ATT_TRACE
// slot layout: [role 0..2][t 0..63][event 0..7]
#define ATT_TR(role, t, ev)                                                                   \
    do {                                                                                        \
        if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (t) < 64)                  \
            p.trace[((role) * 64 + (t)) * 8 + (ev)] = clock64();                                \
    } while (0)
#else
#define ATT_TR(role, t, ev) do {} while (0)
#endif

template <int HD>
struct AttnCfg {
    static constexpr int TILE_BYTES = QT * HD * 2;          // one Q / K / V tile
    static constexpr int BOX_BYTES = QT * 64 * 2;           // one 64-column TMA box (16 KiB)
    static constexpr int NBOX = HD / 64;
    static constexpr int Q_OFF = 0;                         // 2 query tiles
    static constexpr int K_OFF = 2 * TILE_BYTES;            // 2 stages
    static constexpr int V_OFF = 4 * TILE_BYTES;            // 2 stages
    static constexpr int BAR_OFF = 6 * TILE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
    static constexpr int TM_S = 0;                          // S_i buffer b at TM_S + 128 i + 64 b
    static constexpr int TM_O = 256;                        // O_i at TM_O + HD i
};

// POLY = how many of the 8 key pairs of every 16-key chunk take the polynomial instead of MUFU.EX2
template <int HD, int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
    using Cfg = AttnCfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + Cfg::BAR_OFF;
    // barriers: q_full | k_full[2] | k_empty[2] | v_full[2] | v_empty[2] | s_full[2][2] | p_full[2][2] |
    //           pv_done[2][2] | o_full[2] | tmem ptr
    const uint32_t q_full = bars;
    auto k_full = [&](int s) { return bars + 8 * (1 + s); };
    auto k_empty = [&](int s) { return bars + 8 * (3 + s); };
    auto v_full = [&](int s) { return bars + 8 * (5 + s); };
    auto v_empty = [&](int s) { return bars + 8 * (7 + s); };
    auto s_full = [&](int i, int b) { return bars + 8 * (9 + 2 * i + b); };
    auto p_full = [&](int i, int b) { return bars + 8 * (13 + 2 * i + b); };
    auto pv_done = [&](int i, int b) { return bars + 8 * (17 + 2 * i + b); };
    auto o_full = [&](int i) { return bars + 8 * (21 + i); };
    const uint32_t tmem_ptr_addr = bars + 8 * 23;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Cfg::BAR_OFF + 8 * 23);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_pairs = (p.L - p.q_begin + 2 * QT - 1) / (2 * QT);
    const int qp = blockIdx.x % q_pairs;
    const int bh = blockIdx.x / q_pairs;
    const int h = bh % p.H, b = bh / p.H;
    const int q0 = p.q_begin + qp * 2 * QT;
    const int n_qt = (q0 + QT < p.L) ? 2 : 1;               // second query tile entirely out of range?
    const int n_kv = (p.L + KT - 1) / KT;                   // TMA tiles
    const int T = (p.L + KS - 1) / KS;                      // score sub-tiles
    const int tail = p.L - (T - 1) * KS;                    // valid keys in the last sub-tile (1..64)
    const int tail16 = (tail + 15) & ~15;

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(v_empty(s), 1);
            mbar_init(o_full(s), 1);
            for (int bb = 0; bb < 2; ++bb) {
                mbar_init(s_full(s, bb), 1);
                mbar_init(p_full(s, bb), 128);
                mbar_init(pv_done(s, bb), 1);
            }
        }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) {
        tmem_alloc<1>(tmem_ptr_addr, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;

    if (warp == TMA_WARP) {
        // ======================================= TMA producer =======================================
        // the whole warp walks the loop (uniform addresses), one elected lane issues
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, n_qt * Cfg::TILE_BYTES);
            for (int i = 0; i < n_qt; ++i)
                for (int c = 0; c < Cfg::NBOX; ++c)
                    tma_load_3d(sbase + Cfg::Q_OFF + i * Cfg::TILE_BYTES + c * Cfg::BOX_BYTES, &map_q, q_full,
                                h * HD + c * 64, q0 + i * QT, b, kEvictFirst);
        }
        __syncwarp();
        for (int j = 0; j < n_kv; ++j) {
            const int s = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(k_empty(s), ph ^ 1, 10);
            if (elect_one()) {
                mbar_arrive_expect_tx(k_full(s), Cfg::TILE_BYTES);
                for (int c = 0; c < Cfg::NBOX; ++c)
                    tma_load_3d(sbase + Cfg::K_OFF + s * Cfg::TILE_BYTES + c * Cfg::BOX_BYTES, &map_k, k_full(s),
                                h * HD + c * 64, j * KT, b, kEvictLast);
            }
            __syncwarp();
            mbar_wait(v_empty(s), ph ^ 1, 11);
            if (elect_one()) {
                mbar_arrive_expect_tx(v_full(s), Cfg::TILE_BYTES);
                for (int c = 0; c < Cfg::NBOX; ++c)
                    tma_load_3d(sbase + Cfg::V_OFF + s * Cfg::TILE_BYTES + c * Cfg::BOX_BYTES, &map_v, v_full(s),
                                h * HD + c * 64, j * KT, b, kEvictLast);
            }
            __syncwarp();
        }
    } else if (warp == MMA_WARP) {
        // ======================================= MMA issuer =======================================
        // the whole warp walks the schedule and waits; one elected lane issues MMAs and commits
        const uint64_t kdesc_hi = umma_desc_kmajor_sw128(0);
        const uint64_t vdesc_hi = umma_desc_mnmajor_sw128(0, Cfg::BOX_BYTES);
        // S_i(t) = Q_i . K_t^T : M=128, N=keys of the sub-tile, K=HD, both operands K-major.  Sub-tile t is
        // rows [64 (t&1), +64) of TMA tile t>>1: 64 rows x 128 B = 8 swizzle atoms further on.
        auto issue_s = [&](int i, int t) {
            const int keys = (t == T - 1) ? tail16 : KS;
            const uint32_t idesc = umma_idesc_bf16(QT, keys);
            const uint32_t qa = (sbase + Cfg::Q_OFF + i * Cfg::TILE_BYTES) >> 4;
            const uint32_t ka = (sbase + Cfg::K_OFF + ((t >> 1) & 1) * Cfg::TILE_BYTES + (t & 1) * (KS * 128)) >> 4;
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
                const uint32_t off = ((k >> 2) * Cfg::BOX_BYTES + (k & 3) * 32) >> 4;
                umma_bf16_ss<1>(tmem + Cfg::TM_S + 128 * i + KS * (t & 1), kdesc_hi | (uint64_t)(qa + off),
                                kdesc_hi | (uint64_t)(ka + off), idesc, k != 0);
            }
        };
        // O_i += P_i(t) . V_t : M=128, N=HD, K=keys; A = P in TMEM (bf16 pairs), B = V MN-major
        auto issue_pv = [&](int i, int t) {
            const int keys = (t == T - 1) ? tail16 : KS;
            constexpr uint32_t idesc = umma_idesc_bf16(QT, HD, 0, 1);
            const uint32_t va = (sbase + Cfg::V_OFF + ((t >> 1) & 1) * Cfg::TILE_BYTES + (t & 1) * (KS * 128)) >> 4;
            for (int k = 0; k < keys / 16; ++k)
                umma_bf16_ts(tmem + Cfg::TM_O + HD * i, tmem + Cfg::TM_S + 128 * i + KS * (t & 1) + 8 * k,
                             vdesc_hi | (uint64_t)(va + k * (2048 >> 4)), idesc, (t | k) != 0);
        };
        // a K (V) stage is free once the last S (PV) of its second sub-tile — or of the very last sub-tile — retires
        auto last_of_tile = [&](int t) { return (t & 1) || t == T - 1; };
        mbar_wait(q_full, 0, 20);
        mbar_wait(k_full(0), 0, 21);
        tc_fence_after();
        if (elect_one()) {
            for (int t = 0; t < 2 && t < T; ++t) {
                for (int i = 0; i < n_qt; ++i) {
                    issue_s(i, t);
                    umma_commit(s_full(i, t));
                }
                if (last_of_tile(t)) umma_commit(k_empty(0));
            }
        }
        __syncwarp();
        for (int t = 0; t < T; ++t) {
            const int j = t >> 1, bb = t & 1;
            const bool more = t + 2 < T;
            if (bb == 0) {
                mbar_wait(v_full(j & 1), (j >> 1) & 1, 22);
                if (more) mbar_wait(k_full((j + 1) & 1), ((j + 1) >> 1) & 1, 23);
            }
            for (int i = 0; i < n_qt; ++i) {
                ATT_TR(2, t, 4 * i + 0);
                mbar_wait(p_full(i, bb), j & 1, 24);
                tc_fence_after();
                ATT_TR(2, t, 4 * i + 1);
                if (elect_one()) {
                    issue_pv(i, t);
                    umma_commit(pv_done(i, bb));
                    if (i == n_qt - 1 && last_of_tile(t)) umma_commit(v_empty(j & 1));
                    if (more) {
                        issue_s(i, t + 2);
                        umma_commit(s_full(i, bb));
                        if (i == n_qt - 1 && last_of_tile(t + 2)) umma_commit(k_empty((j + 1) & 1));
                    }
                    if (t == T - 1) umma_commit(o_full(i));
                }
                __syncwarp();
                ATT_TR(2, t, 4 * i + 2);
            }
        }
    } else {
        // ======================================= softmax =======================================
        const int i = warp >> 2;                        // query tile
        const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
        if (i < n_qt) {
            const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
            const uint32_t t_s = tmem + Cfg::TM_S + 128 * i + lane_off;
            const uint32_t t_o = tmem + Cfg::TM_O + HD * i + lane_off;
            const int qrow = q0 + i * QT + quarter * 32 + lane;
            float m_used = -INFINITY, l_sum = 0.f;
            for (int t = 0; t < T; ++t) {
                const int bb = t & 1;
                const int keys = (t == T - 1) ? tail : KS;              // valid keys
                const int keys16 = (keys + 15) & ~15;
                if (quarter == 0) ATT_TR(i, t, 0);
                mbar_wait(s_full(i, bb), (t >> 1) & 1, 30);
                tc_fence_after();
                if (quarter == 0) ATT_TR(i, t, 1);
                uint32_t sv[64];
                tmem_ld_32x32b_x32(t_s + KS * bb, &sv[0]);
                if (keys16 > 32) tmem_ld_32x32b_x32(t_s + KS * bb + 32, &sv[32]);
                tmem_ld_wait();
                if (quarter == 0) ATT_TR(i, t, 2);
                if (keys < KS) {
#pragma unroll
                    for (int c = 0; c < 64; ++c)
                        if (c >= keys) sv[c] = 0xff800000u;            // -inf: masked (or never written) key
                }
                float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // 4 independent chains
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        mxa[u] = fmaxf(mxa[u], fmaxf(__uint_as_float(sv[c + 2 * u]), __uint_as_float(sv[c + 2 * u + 1])));
                }
                const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
                // lazy rescale: keep the stale reference max unless it grows by more than 2^8
                const float m_new = fmaxf(m_used, mx);
                const bool grow = (m_new - m_used) * p.scale_log2 > 8.0f;
                if (t == 0) {
                    m_used = m_new;
                } else if (__any_sync(0xffffffffu, grow)) {
                    // O_i must be quiescent: PV_i(t-1) may still be in flight
                    mbar_wait(pv_done(i, bb ^ 1), ((t - 1) >> 1) & 1, 32);
                    tc_fence_after();
                    const float alpha = grow ? ex2_mufu((m_used - m_new) * p.scale_log2) : 1.0f;
                    if (grow) m_used = m_new;
                    l_sum *= alpha;
#pragma unroll 1
                    for (int c = 0; c < HD / 16; ++c) {
                        uint32_t ov[16];
                        tmem_ld_32x32b_x16(t_o + c * 16, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 16; ++u) ov[u] = __float_as_uint(__uint_as_float(ov[u]) * alpha);
                        tmem_st_32x32b_x16(t_o + c * 16, ov);
                    }
                    tmem_st_wait();
                }
                if (quarter == 0) ATT_TR(i, t, 3);
                const float mb = m_used * p.scale_log2;
                const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nmb2 = make_float2(-mb, -mb);
                float2 rs2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                // P column u holds the bf16 pair for keys (2u, 2u+1); written over the scores, 16 keys at a time
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c * 16 < keys16) {
                        uint32_t pw[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            // packed fp32x2: one FFMA2 scales-and-shifts two scores, one FADD2 adds two exponentials
                            const float2 x = ffma2(make_float2(__uint_as_float(sv[c * 16 + 2 * u]), __uint_as_float(sv[c * 16 + 2 * u + 1])),
                                                   sc2, nmb2);
                            // spread the polynomial pairs evenly through the chunk
                            const bool poly = POLY > 0 && ((u + 1) * POLY / 8 != u * POLY / 8);
                            const float2 e = poly ? ex2_poly2(x) : make_float2(ex2_mufu(x.x), ex2_mufu(x.y));
                            rs2[u & 3] = fadd2(rs2[u & 3], e);
                            pw[u] = pack_bf16(e.x, e.y);
                        }
                        tmem_st_32x32b_x8(t_s + KS * bb + c * 8, pw);
                    }
                }
                l_sum += (rs2[0].x + rs2[0].y) + (rs2[1].x + rs2[1].y) + (rs2[2].x + rs2[2].y) + (rs2[3].x + rs2[3].y);
                if (quarter == 0) ATT_TR(i, t, 4);
                tmem_st_wait();
                tc_fence_before();
                if (quarter == 0) ATT_TR(i, t, 5);
                mbar_arrive(p_full(i, bb));
                if (quarter == 0) ATT_TR(i, t, 6);
            }
            // ---- epilogue: O / l -> bf16, token-major
            mbar_wait(o_full(i), 0, 31);
            tc_fence_after();
            const float inv = 1.0f / l_sum;
            __nv_bfloat16* orow = p.out + ((int64_t)b * p.L + qrow) * p.ldo + h * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(t_o + c * 32, ov);
                tmem_ld_wait();
                if (qrow < p.L) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 w;
                        w.x = pack_bf16(__uint_as_float(ov[8 * u + 0]) * inv, __uint_as_float(ov[8 * u + 1]) * inv);
                        w.y = pack_bf16(__uint_as_float(ov[8 * u + 2]) * inv, __uint_as_float(ov[8 * u + 3]) * inv);
                        w.z = pack_bf16(__uint_as_float(ov[8 * u + 4]) * inv, __uint_as_float(ov[8 * u + 5]) * inv);
                        w.w = pack_bf16(__uint_as_float(ov[8 * u + 6]) * inv, __uint_as_float(ov[8 * u + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + 8 * u) = w;
                    }
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc<1>(tmem, 512);
    }
}

#ifdef MMADA_ATT_TRACE
static long long* g_attn_trace = nullptr;
#endif

template <int HD, int POLY>
static int launch_attention(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                            int H, float scale, int q_begin, cudaStream_t stream) {
    using Cfg = AttnCfg<HD>;
    CUtensorMap mq, mk, mv;
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
    const uint32_t box[3] = {64, 128, 1};
    int st;
    if ((st = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box))) return st;
    if ((st = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box))) return st;
    if ((st = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box))) return st;
    auto kern = attention_kernel<HD, POLY>;
    static bool configured[kMaxDevices] = {};
    MMADA_CUDA_TRY(ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, configured));
    AttnParams p;
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.L = L; p.H = H; p.B = B;
    p.q_begin = q_begin;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef MMADA_ATT_TRACE
    p.trace = g_attn_trace;
#endif
    const int q_pairs = (L - q_begin + 2 * QT - 1) / (2 * QT);
    kern<<<B * H * q_pairs, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(mq, mk, mv, p);
    return cuda_status(cudaGetLastError());
}

// share of the exponentials on the FMA pipe, in eighths (EXPERIMENTS builds: MMADA_ATT_POLY; -1 = not set: every kernel
// then uses the default its own measurements gave)
static int attention_poly_env() {
    static int v = -2;
    if (v == -2) {
        v = experiment_env("MMADA_ATT_POLY", -1);
        if (v != 0 && v != 2 && v != 3 && v != 4) v = -1;
    }
    return v;
}
static int attention_poly_eighths() { return attention_poly_env() < 0 ? kAttnPolyDefault : attention_poly_env(); }

template <int HD>
static int dispatch_attention(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                              int H, float scale, int q_begin, cudaStream_t stream) {
    switch (attention_poly_eighths()) {
        case 0: return launch_attention<HD, 0>(q, k, v, ld, out, ldo, B, L, H, scale, q_begin, stream);
        case 2: return launch_attention<HD, 2>(q, k, v, ld, out, ldo, B, L, H, scale, q_begin, stream);
        case 3: return launch_attention<HD, 3>(q, k, v, ld, out, ldo, B, L, H, scale, q_begin, stream);
        default: return launch_attention<HD, 4>(q, k, v, ld, out, ldo, B, L, H, scale, q_begin, stream);
    }
}

// attention_duo64.cu (the product's head_dim-128 kernel): persistent, two query tiles per CTA, one thread per query row,
// the scores of every tile double-buffered in 64-key sub-tiles, one MMA-issuing warp per tile
int launch_attention_duo64(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                           int Lq, int H, float scale, int poly, cudaStream_t stream);
#ifdef MMADA_EXPERIMENTS
// EXPERIMENTS builds only (A/B runs): attention_pair64.cu = attention_duo64.cu on CTA pairs (0.658 ms);
// attention_duo.cu = two query tiles in ping-pong with P over its own 128-key score tile (the product kernel before
// attention_duo64.cu); attention_quad.cu = CTA pairs with two
// 256-row query blocks in flight and P through shared memory; attention_pair.cu = round 1's pair kernel
int launch_attention_duo(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                         int Lq, int H, float scale, int poly, cudaStream_t stream);
int launch_attention_pair64(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                            int Lq, int H, float scale, int poly, cudaStream_t stream);
int launch_attention_quad(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                          int Lq, int H, float scale, int poly, cudaStream_t stream);
int launch_attention_pair(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                          int Lq, int H, float scale, int poly, cudaStream_t stream);
#endif

}  // namespace mmada

using namespace mmada;

#ifdef MMADA_ATT_TRACE
extern "C" void mmada_attention_set_trace(void* buf) { mmada::g_attn_trace = (long long*)buf; }
#endif

extern "C" int mmada_attention_bf16(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo,
                                    int B, int L, int H, int head_dim, float scale, void* stream) {
    if (!q || !k || !v || !out || B <= 0 || L <= 0 || H <= 0) return kBadArgument;
    if ((ld % 8) || (ldo % 8)) return kUnsupportedShape;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(out)) & 15)
        return kBadArgument;
    cudaStream_t s = (cudaStream_t)stream;
    if (head_dim == 128) {
        if (L > 128) {
            // The persistent kernel walks items of 256 query rows; a last item with at most one 128-row tile (L = 1539: 3 rows)
            // runs with one slot.  With attention_duo64.cu that costs less than handing those rows to the single-CTA kernel
            // in a second launch, which re-reads all of K and V from HBM (0.605 against 0.629 ms at config 2, 29.1 against
            // 29.9 ms per step): the split below is kept for A/B runs of the EXPERIMENTS build only.
            auto main_kernel = [&](int Lq) {
#ifdef MMADA_EXPERIMENTS
                // MMADA_ATT_KERNEL = 0 round-1 pair kernel, 1 attention_duo.cu, 2 attention_quad.cu, 3 attention_duo64.cu (the product's), 4 attention_pair64.cu
                static const int which = experiment_env("MMADA_ATT_KERNEL", 3);
                if (which == 4) return launch_attention_pair64(q, k, v, ld, out, ldo, B, L, Lq, H, scale, attention_poly_env(), s);
                if (which == 1) return launch_attention_duo(q, k, v, ld, out, ldo, B, L, Lq, H, scale, attention_poly_env(), s);
                if (which == 2) return launch_attention_quad(q, k, v, ld, out, ldo, B, L, Lq, H, scale, attention_poly_env(), s);
                if (which == 0) return launch_attention_pair(q, k, v, ld, out, ldo, B, L, Lq, H, scale, attention_poly_env(), s);
#endif
                return launch_attention_duo64(q, k, v, ld, out, ldo, B, L, Lq, H, scale, attention_poly_env(), s);
            };
            static const int split_tail = experiment_env("MMADA_ATT_SPLIT_TAIL", 0);
            const int rem = L % 256;
            if (split_tail && L > 256 && rem > 0 && rem <= 128 && B * H >= num_sms()) {
                const int st = main_kernel(L - rem);
                if (st) return st;
                return dispatch_attention<128>(q, k, v, ld, out, ldo, B, L, H, scale, L - rem, s);
            }
            return main_kernel(L);
        }
        return dispatch_attention<128>(q, k, v, ld, out, ldo, B, L, H, scale, 0, s);
    }
    if (head_dim == 64) return dispatch_attention<64>(q, k, v, ld, out, ldo, B, L, H, scale, 0, s);
    return kUnsupportedShape;
}
