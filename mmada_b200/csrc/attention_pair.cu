// Bidirectional flash attention, head_dim 128: persistent CTA PAIRS (cta_group::2), tcgen05 / TMEM / TMA.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False) at
// /root/reference/models/modeling_llada.py:653-660 (no mask is ever applied).
//
// Why pairs.  The score MMA wants N >= 128 (a 128x64 UMMA was measured to cost as much as a 128x128 one) and
// the chain  S -> softmax -> P -> PV -> next S  must not be serial per query tile (the single-CTA kernel in
// attention.cu sits at 40 % tensor pipe because of it), so every query tile needs TWO 128-column score buffers
// next to its output accumulator: 384 of the 512 TMEM columns for one 128-row tile.  One tile per SM would
// double the K/V traffic from L2 (64 B/clk/SM, above the L2 limit), unless the two SMs of a pair share K and V:
// a 256 x 128 x 16 UMMA (cta_group::2) takes its A rows from both CTAs and HALF of its B operand from each, so
// every SM loads half of each K tile (64 keys) and half of each V tile (64 of the 128 head columns).
//
// A pair walks a list of work items (batch, head, 256 query rows), persistent, as ONE continuous stream of key
// tiles g = 0,1,2,...: score buffer g&1, probability buffer g&1 (P does NOT alias S: tile g+2's scores only need
// the softmax threads to have READ tile g's, not PV(g) to have retired).  The MMA warp of the leader CTA issues
//     S(0) S(1) | scores(0) read: S(2) | P(0) ready: PV(0) | scores(1) read: S(3) | P(1) ready: PV(1) | ...
// across item boundaries, so the softmax threads always find the next scores waiting and the tensor pipe always
// has the next score MMA queued behind the current PV.  The epilogue (separate warps) copies O to registers,
// releases the accumulator and then stores.
//   warps 0-7    softmax: 16 query rows per warp in the m16n8 fragment layout (a row = one quad of threads);
//                exponentials against a running reference maximum that is only revisited when an exponential
//                overflows its 2^8 window (read off the partial row sums)
//   warps 8-11   epilogue: O / l -> bf16, token-major, one thread per query row
//   warp 12      TMA producer (each CTA loads its halves; full barriers live in the leader CTA)
//   warp 13      MMA issuer (leader CTA only)
// TMEM (per CTA): S buffer 0 | S buffer 1 (128 columns each) | P buffer 0 | P buffer 1 (64 each) | O (128).
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "attn_math.cuh"
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

namespace {

constexpr int P_THREADS = 448;
constexpr int P_EPI_WARP0 = 8, P_TMA_WARP = 12, P_MMA_WARP = 13;
constexpr int HD = 128;
#ifndef MMADA_ATT_SLEEP_NS
#define MMADA_ATT_SLEEP_NS 200
#endif
constexpr unsigned SLEEP_NS = MMADA_ATT_SLEEP_NS;   // poll interval of the warps off the critical path
constexpr int KST = 4, VST = 4;                     // K / V ring depths
constexpr int Q_BYTES = 128 * HD * 2;               // 128 query rows (two 64-column boxes of 16 KiB)
constexpr int K_BYTES = 64 * HD * 2;                // 64 keys (two 64-column boxes of 8 KiB)
constexpr int V_BYTES = 128 * 64 * 2;               // 128 keys x 64 head columns (one box)
constexpr int Q_OFF = 0;                            // 2 buffers
constexpr int K_OFF = 2 * Q_BYTES;
constexpr int V_OFF = K_OFF + KST * K_BYTES;
constexpr int BAR_OFF = V_OFF + VST * V_BYTES;
constexpr int XCHG_OFF = BAR_OFF + 512;             // (spare)
constexpr int LBUF_OFF = XCHG_OFF + 2048;           // float [2 O buffers][128 rows]: row sums for the epilogue
constexpr int P_SMEM_BYTES = LBUF_OFF + 4096 + 1024;
constexpr int TM_S = 0, TM_P = 256, TM_O = 384;     // S buffers 2 x 128 | P buffers 2 x 64 | O 128 columns

struct PairParams {
    __nv_bfloat16* out;
    int64_t ldo;
    int L, H, B;
    int q_pairs, items;
    float scale_log2;
#ifdef MMADA_ATT_TRACE
    long long* trace;
#endif
};

#ifdef MMADA_ATT_TRACE
#define PTR(role, g, ev)                                                                                   \
    do {                                                                                                   \
        if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g) < 64)                             \
            p.trace[((role) * 64 + (g)) * 8 + (ev)] = clock64();                                           \
    } while (0)
#else
#define PTR(role, g, ev) do {} while (0)
#endif

// barrier indices (8 bytes each)
enum : int {
    B_QFULL = 0, B_QEMPTY = 2, B_KFULL = 4, B_KEMPTY = 4 + KST, B_VFULL = 4 + 2 * KST, B_VEMPTY = 4 + 2 * KST + VST,
    B_SFULL = 4 + 2 * KST + 2 * VST, B_SFREE = B_SFULL + 2, B_PFULL = B_SFREE + 2, B_PVDONE = B_PFULL + 2,
    B_OFULL = B_PVDONE + 2, B_OEMPTY = B_OFULL + 2, B_LFULL = B_OEMPTY + 2, B_TMEMPTR = B_LFULL + 2
};
static_assert(B_TMEMPTR * 8 + 8 <= 512, "barrier block");

template <int POLY, bool MAXCHK>
__global__ void __launch_bounds__(P_THREADS, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const PairParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    auto bar = [&](int idx) { return sbase + BAR_OFF + 8 * idx; };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 8 * B_TMEMPTR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
    const int n_items = (p.items - cluster_id + num_clusters - 1) / num_clusters;
    const int T = (p.L + 127) / 128;                        // key tiles per item
    const int tail = p.L - (T - 1) * 128;                   // valid keys in the last tile (1..128)
    const int tail16 = (tail + 15) & ~15;
    const int G = n_items * T;                              // this pair's stream of key tiles

    if (warp == P_TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_QFULL + i), 1);
            mbar_init(bar(B_QEMPTY + i), 1);
            mbar_init(bar(B_SFULL + i), 1);
            mbar_init(bar(B_SFREE + i), 16);     // one arrival per softmax warp of both CTAs
            mbar_init(bar(B_PFULL + i), 16);     // one arrival per softmax warp of both CTAs
            mbar_init(bar(B_PVDONE + i), 1);
            mbar_init(bar(B_OFULL + i), 1);
            mbar_init(bar(B_OEMPTY + i), 8);     // one arrival per epilogue warp of both CTAs
            mbar_init(bar(B_LFULL + i), 8);      // one arrival per softmax warp of this CTA
        }
        for (int s = 0; s < KST; ++s) { mbar_init(bar(B_KFULL + s), 1); mbar_init(bar(B_KEMPTY + s), 1); }
        for (int s = 0; s < VST; ++s) { mbar_init(bar(B_VFULL + s), 1); mbar_init(bar(B_VEMPTY + s), 1); }
        fence_mbar_init();
    }
    if (warp == P_MMA_WARP) {
        tmem_alloc<2>(bar(B_TMEMPTR), 512);
        tmem_relinquish<2>();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;

    auto item_coords = [&](int n, int& b, int& h, int& q0) {
        const int id = cluster_id + n * num_clusters;
        const int qp = id % p.q_pairs, bh = id / p.q_pairs;
        h = bh % p.H;
        b = bh / p.H;
        q0 = qp * 256 + (int)rank * 128;
    };

    if (warp == P_TMA_WARP) {
        // ======================================= TMA producer =======================================
        const uint32_t lead0 = mapa_u32(bar(0), 0);               // the leader CTA's barrier block
        auto lbar = [&](int idx) { return lead0 + 8 * idx; };
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0;
            item_coords(n, b, h, q0);
            const int qb = n & 1;
            mbar_wait_backoff(bar(B_QEMPTY + qb), ((n >> 1) & 1) ^ 1, 10, SLEEP_NS);
            if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(bar(B_QFULL + qb), 2 * Q_BYTES);
                const int qrow = q0 < p.L ? q0 : 0;               // a tile entirely past the end: any rows, never stored
                for (int c = 0; c < 2; ++c)
                    tma_load_3d_2sm(sbase + Q_OFF + qb * Q_BYTES + c * (Q_BYTES / 2), &map_q, lbar(B_QFULL + qb),
                                    h * HD + c * 64, qrow, b, kEvictFirst);
            }
            __syncwarp();
            for (int j = 0; j < T; ++j) {
                const int g = n * T + j;
                const int ks = g % KST, vs = g % VST;
                mbar_wait_backoff(bar(B_KEMPTY + ks), ((g / KST) & 1) ^ 1, 11, SLEEP_NS);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(bar(B_KFULL + ks), 2 * K_BYTES);
                    // this CTA's half of the N keys the score MMA covers (N = 128, or tail16 in the last tile)
                    const int half_n = (j == T - 1 ? tail16 : 128) / 2;
                    for (int c = 0; c < 2; ++c)
                        tma_load_3d_2sm(sbase + K_OFF + ks * K_BYTES + c * (K_BYTES / 2), &map_k, lbar(B_KFULL + ks),
                                        h * HD + c * 64, j * 128 + (int)rank * half_n, b, kEvictLast);
                }
                __syncwarp();
                mbar_wait_backoff(bar(B_VEMPTY + vs), ((g / VST) & 1) ^ 1, 12, SLEEP_NS);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(bar(B_VFULL + vs), 2 * V_BYTES);
                    tma_load_3d_2sm(sbase + V_OFF + vs * V_BYTES, &map_v, lbar(B_VFULL + vs), h * HD + (int)rank * 64,
                                    j * 128, b, kEvictLast);
                }
                __syncwarp();
            }
        }
    } else if (warp == P_MMA_WARP) {
        // ======================================= MMA issuer (leader) =======================================
        if (leader && G > 0) {
            const uint64_t kdesc_hi = umma_desc_kmajor_sw128(0);
            const uint64_t vdesc_hi = umma_desc_mnmajor_sw128(0, V_BYTES);
            // S(g) = Q . K^T : M = 256 (128 rows per CTA), N = keys, K = 128; operands K-major
            auto issue_s = [&](int g, int n, int j) {
                const int qb = n & 1, ks = g % KST, buf = g & 1;
                PTR(3, g, 0);
                if (j == 0) mbar_wait(bar(B_QFULL + qb), (n >> 1) & 1, 20);
                mbar_wait(bar(B_KFULL + ks), (g / KST) & 1, 21);
                tc_fence_after();
                PTR(3, g, 1);
                if (elect_one()) {
                    const uint32_t idesc = umma_idesc_bf16(256, j == T - 1 ? tail16 : 128);
                    const uint32_t qa = (sbase + Q_OFF + qb * Q_BYTES) >> 4;
                    const uint32_t ka = (sbase + K_OFF + ks * K_BYTES) >> 4;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint32_t qoff = ((k >> 2) * (Q_BYTES / 2) + (k & 3) * 32) >> 4;
                        const uint32_t koff = ((k >> 2) * (K_BYTES / 2) + (k & 3) * 32) >> 4;
                        umma_bf16_ss<2>(tmem + TM_S + 128 * buf, kdesc_hi | (uint64_t)(qa + qoff),
                                        kdesc_hi | (uint64_t)(ka + koff), idesc, k != 0);
                    }
                    umma_commit_2sm(bar(B_KEMPTY + ks), 0x3);
                    umma_commit_2sm(bar(B_SFULL + buf), 0x3);
                    if (j == T - 1) umma_commit_2sm(bar(B_QEMPTY + qb), 0x3);
                }
                PTR(3, g, 2);
                __syncwarp();
                PTR(3, g, 3);
            };
            // O += P(g) . V : M = 256, N = 128 head columns (64 per CTA), K = keys; A = P in TMEM, B = V MN-major
            auto issue_pv = [&](int g, int n, int j) {
                const int vs = g % VST, buf = g & 1;
                mbar_wait(bar(B_VFULL + vs), (g / VST) & 1, 22);
                if (j == 0) mbar_wait(bar(B_OEMPTY), (n & 1) ^ 1, 23);      // epilogue of the previous item has read O
                tc_fence_after();
                if (elect_one()) {
                    constexpr uint32_t idesc = umma_idesc_bf16(256, HD, 0, 1);
                    const uint32_t va = (sbase + V_OFF + vs * V_BYTES) >> 4;
                    const int ksteps = (j == T - 1 ? tail16 : 128) / 16;
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16_ts_cg<2>(tmem + TM_O, tmem + TM_P + 64 * buf + 8 * k,
                                           vdesc_hi | (uint64_t)(va + k * (2048 >> 4)), idesc, (j | k) != 0);
                    umma_commit_2sm(bar(B_VEMPTY + vs), 0x3);
                    umma_commit_2sm(bar(B_PVDONE + buf), 0x3);
                    if (j == T - 1) umma_commit_2sm(bar(B_OFULL), 0x3);
                }
                __syncwarp();
            };
            // (item, tile) of the score stream (two tiles ahead) and of the PV stream
            int ns = 0, js = 0, np = 0, jp = 0;
            auto step = [&](int& n, int& j) { if (++j == T) { j = 0; ++n; } };
            issue_s(0, ns, js);
            step(ns, js);
            if (G > 1) { issue_s(1, ns, js); step(ns, js); }
            for (int g = 0; g < G; ++g) {
                // the scores of tile g are in registers: their buffer can take tile g+2 while the softmax runs
                PTR(2, g, 0);
                if (g + 2 < G) {
                    mbar_wait(bar(B_SFREE + (g & 1)), (g >> 1) & 1, 25);
                    tc_fence_after();
                    issue_s(g + 2, ns, js);
                    step(ns, js);
                }
                PTR(2, g, 1);
                mbar_wait(bar(B_PFULL + (g & 1)), (g >> 1) & 1, 24);
                tc_fence_after();
                PTR(2, g, 2);
                issue_pv(g, np, jp);
                step(np, jp);
                PTR(2, g, 3);
            }
        }
    } else if (warp >= P_EPI_WARP0) {
        // ======================================= epilogue =======================================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t oempty_lead = mapa_u32(bar(B_OEMPTY), 0);
        const float* lbuf = reinterpret_cast<const float*>(smem + LBUF_OFF);
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0;
            item_coords(n, b, h, q0);
            const int ob = n & 1;
            const int qrow = q0 + row;
            mbar_wait_backoff(bar(B_LFULL + ob), (n >> 1) & 1, 41, 4 * SLEEP_NS);
            const float inv = 1.0f / lbuf[ob * 128 + row];
            mbar_wait_backoff(bar(B_OFULL), n & 1, 40, SLEEP_NS);
            tc_fence_after();
            // O -> registers (bf16) first, so that the accumulator is free for the next item before the stores go out
            uint32_t ow[HD / 2];
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(tmem + TM_O + c * 32 + lane_off, ov);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    ow[c * 16 + u] = pack_bf16(__uint_as_float(ov[2 * u]) * inv, __uint_as_float(ov[2 * u + 1]) * inv);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(oempty_lead);
            if (qrow < p.L) {
                __nv_bfloat16* orow = p.out + ((int64_t)b * p.L + qrow) * p.ldo + h * HD;
#pragma unroll
                for (int u = 0; u < HD / 8; ++u)
                    *reinterpret_cast<uint4*>(orow + 8 * u) = make_uint4(ow[4 * u], ow[4 * u + 1], ow[4 * u + 2], ow[4 * u + 3]);
            }
        }
    } else {
        // ======================================= softmax =======================================
        // Warp (quarter, sub) owns 16 query rows (TMEM lanes 32 quarter + 16 sub ..+15) and ALL keys of every tile,
        // in the m16n8 fragment layout (tcgen05.ld.16x256b): a row lives in the 4 threads of a quad, 32 scores
        // each, so a row maximum costs two shuffles and nothing has to cross warps.  The two warps that share a
        // scheduler (sub 0 / 1 of one quarter) are independent.
        const int quarter = warp & 3, sub = warp >> 2;
        const int q4 = lane & 3;
        const int r0 = quarter * 32 + sub * 16 + (lane >> 2);      // rows r0 and r0 + 8
        const uint32_t lane_off = (uint32_t)(quarter * 32 + sub * 16) << 16;
        const uint32_t pfull_lead = mapa_u32(bar(B_PFULL), 0), sfree_lead = mapa_u32(bar(B_SFREE), 0);
        float* lbuf = reinterpret_cast<float*>(smem + LBUF_OFF);
        float m_used[2] = {-INFINITY, -INFINITY}, l_sum[2] = {0.f, 0.f};
        int n = 0, j = 0;
        for (int g = 0; g < G; ++g) {
            const int buf = g & 1, ob = n & 1;
            const int keys = (j == T - 1) ? tail : 128;             // valid keys of this tile
            const int keys16 = (keys + 15) & ~15;
            const uint32_t t_s = tmem + TM_S + 128 * buf + lane_off, t_p = tmem + TM_P + 64 * buf + lane_off;
            if (j == 0) { m_used[0] = m_used[1] = -INFINITY; l_sum[0] = l_sum[1] = 0.f; }
            if (warp == 0) PTR(0, g, 0);
            mbar_wait(bar(B_SFULL + buf), (g >> 1) & 1, 30);
            tc_fence_after();
            if (warp == 0) PTR(0, g, 1);
            // sv[4i + c] = row r0, key 8i + 2 q4 + c;  sv[4i + 2 + c] = row r0 + 8, same key
            uint32_t sv[64];
            tmem_ld_16x256b_x8(t_s, &sv[0]);
            if (64 < keys16) tmem_ld_16x256b_x8(t_s + 64, &sv[32]);
            tmem_ld_wait();
            // the scores are in registers: hand the buffer back (tile g+2 may overwrite it)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(sfree_lead + 8 * buf);
            if (warp == 0) PTR(0, g, 2);
            if (keys < 128) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if (8 * i + 2 * q4 + c >= keys) sv[4 * i + c] = sv[4 * i + 2 + c] = 0xff800000u;   // -inf
            }
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
            // row maximum of the tile: per-thread chains, then the quad
            auto row_max = [&](float (&mx)[2]) {
                float mxa[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    mxa[0][i & 1] = fmaxf(mxa[0][i & 1], fmaxf(__uint_as_float(sv[4 * i]), __uint_as_float(sv[4 * i + 1])));
                    mxa[1][i & 1] = fmaxf(mxa[1][i & 1], fmaxf(__uint_as_float(sv[4 * i + 2]), __uint_as_float(sv[4 * i + 3])));
                }
                mx[0] = fmaxf(mxa[0][0], mxa[0][1]);
                mx[1] = fmaxf(mxa[1][0], mxa[1][1]);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
                }
            };
            // exponentials against the reference maxima m_used; P column 4i + q4 holds the bf16 pair of keys
            // (8i + 2 q4, +1): tcgen05.st.16x128b, 64 keys per store.  Returns the thread's partial row sums.
            // NI = 8-key groups per 64-key half that hold valid keys (8 for whole tiles; the last tile of a
            // sequence is often a few keys only — L = 1539: 3 — and then skips the exponentials of the padding)
            float pxm = -INFINITY;      // largest argument handed to the polynomial 2^x (it wraps above 2^128)
            auto exp_tile_n = [&](float (&part)[2], auto ni_tag) {
                constexpr int NI = decltype(ni_tag)::value;
                const float2 nmb2[2] = {make_float2(-m_used[0] * p.scale_log2, -m_used[0] * p.scale_log2),
                                        make_float2(-m_used[1] * p.scale_log2, -m_used[1] * p.scale_log2)};
                float2 rs2[2][2] = {{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}};
                pxm = -INFINITY;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (hh * 64 < keys16 && (NI == 8 || hh == 0)) {
                        uint32_t pw[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                if (i < NI) {
                                    const int s0 = 4 * (hh * 8 + i) + 2 * r;
                                    const float2 x = ffma2(make_float2(__uint_as_float(sv[s0]), __uint_as_float(sv[s0 + 1])), sc2, nmb2[r]);
                                    const int u = 2 * i + r;        // 16 pairs per store; spread the polynomial ones evenly
                                    const bool poly = POLY > 0 && ((u + 1) * POLY / 16 != u * POLY / 16);
                                    const float2 e = poly ? ex2_poly2(x) : make_float2(ex2_mufu(x.x), ex2_mufu(x.y));
                                    if (poly && !MAXCHK) pxm = fmaxf(pxm, fmaxf(x.x, x.y));     // FMNMX3
                                    rs2[r][i & 1] = fadd2(rs2[r][i & 1], e);
                                    pw[2 * i + r] = pack_bf16(e.x, e.y);
                                } else {
                                    pw[2 * i + r] = 0u;
                                }
                            }
                        }
                        tmem_st_16x128b_x8(t_p + hh * 32, pw);
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) part[r] = (rs2[r][0].x + rs2[r][0].y) + (rs2[r][1].x + rs2[r][1].y);
            };
            auto exp_tile = [&](float (&part)[2]) {
                if (keys16 <= 16) exp_tile_n(part, std::integral_constant<int, 2>{});
                else if (keys16 <= 32) exp_tile_n(part, std::integral_constant<int, 4>{});
                else exp_tile_n(part, std::integral_constant<int, 8>{});
            };
            float mx[2], part[2];
            if (j == 0) {
                row_max(mx);                                        // first tile of an item: no reference yet
                m_used[0] = mx[0];
                m_used[1] = mx[1];
            }
            if (warp == 0) PTR(0, g, 3);
            // the P buffer of tile g-2 must have been consumed (it has, long ago, unless the tensor pipe is behind)
            if (g >= 2) mbar_wait(bar(B_PVDONE + buf), ((g >> 1) & 1) ^ 1, 33);
            // SPECULATE that the reference maxima still hold (lazy rescale: they do unless a row maximum grows by
            // more than 2^8): the exponentials start right away.  Whether they held is read off the exponentials
            // themselves — a score more than 2^8 above its reference gives an exponential > 256, hence a partial row
            // sum > 256 (+inf once it overflows) — so the hot path takes no row maximum at all: no FMNMX chain, no
            // shuffles (MAXCHK = the previous scheme: tile maxima taken alongside the exponentials, kept for A/B
            // runs).  The polynomial 2^x wraps for arguments above 128: those are range-checked on their own.
            exp_tile(part);
            if (j != 0) {
                float m_new[2];
                bool grow[2];
                if constexpr (MAXCHK) {
                    row_max(mx);
                    m_new[0] = fmaxf(m_used[0], mx[0]);
                    m_new[1] = fmaxf(m_used[1], mx[1]);
                    grow[0] = (m_new[0] - m_used[0]) * p.scale_log2 > 8.0f;
                    grow[1] = (m_new[1] - m_used[1]) * p.scale_log2 > 8.0f;
                } else {
                    grow[0] = grow[1] = !(part[0] <= 256.0f) || !(part[1] <= 256.0f) || pxm > 120.0f;
                }
                if (__any_sync(0xffffffffu, grow[0] || grow[1])) {
                    if constexpr (!MAXCHK) {            // now the maxima are needed: every row that grew moves its reference
                        row_max(mx);
                        m_new[0] = fmaxf(m_used[0], mx[0]);
                        m_new[1] = fmaxf(m_used[1], mx[1]);
                        grow[0] = m_new[0] > m_used[0];
                        grow[1] = m_new[1] > m_used[1];
                    }
                    // mis-speculated: rescale O and the row sums, redo this tile's exponentials.  The accumulator must
                    // be quiescent: PV(g-1) may still be in flight
                    mbar_wait(bar(B_PVDONE + (buf ^ 1)), ((g - 1) >> 1) & 1, 32);
                    tc_fence_after();
                    float alpha[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        alpha[r] = grow[r] ? ex2_mufu((m_used[r] - m_new[r]) * p.scale_log2) : 1.0f;
                        if (grow[r]) m_used[r] = m_new[r];
                        l_sum[r] *= alpha[r];
                    }
                    const uint32_t t_o = tmem + TM_O + lane_off;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t ov[32];
                        tmem_ld_16x256b_x8(t_o + c * 64, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 32; ++u) ov[u] = __float_as_uint(__uint_as_float(ov[u]) * alpha[(u >> 1) & 1]);
                        tmem_st_16x256b_x8(t_o + c * 64, ov);
                    }
                    tmem_st_wait();
                    exp_tile(part);
                }
            }
            l_sum[0] += part[0];
            l_sum[1] += part[1];
            if (warp == 0) PTR(0, g, 4);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(pfull_lead + 8 * buf);
            if (warp == 0) PTR(0, g, 5);
            if (j == T - 1) {
                // row sums: add up the quad, hand them to the epilogue warps
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float l = l_sum[r];
                    l += __shfl_xor_sync(0xffffffffu, l, 1);
                    l += __shfl_xor_sync(0xffffffffu, l, 2);
                    if (q4 == 0) lbuf[ob * 128 + r0 + 8 * r] = l;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_LFULL + ob));
                j = 0;
                ++n;
            } else {
                ++j;
            }
        }
    }
    // teardown: everyone done with TMEM, and the peer done with our shared memory / barriers
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == P_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc<2>(tmem, 512);
    }
}

#ifdef MMADA_ATT_TRACE
long long* g_pair_trace = nullptr;
#endif

template <int POLY, bool MAXCHK>
int launch_pair(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L, int Lq, int H,
                float scale, cudaStream_t stream) {
    CUtensorMap mq, mk, mv;
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
    const uint32_t box128[3] = {64, 128, 1}, box64[3] = {64, 64, 1};
    int st;
    if ((st = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box128))) return st;
    if ((st = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box64))) return st;
    if ((st = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box128))) return st;
    auto kern = attention_pair_kernel<POLY, MAXCHK>;
    static bool configured[kMaxDevices] = {};
    MMADA_CUDA_TRY(ensure_dynamic_smem(kern, P_SMEM_BYTES, configured));
    PairParams p = {};
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.L = L; p.H = H; p.B = B;
    p.q_pairs = (Lq + 255) / 256;           // query rows [0, Lq): Lq == L, or a multiple of 256 (the rest: another launch)
    p.items = B * H * p.q_pairs;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef MMADA_ATT_TRACE
    p.trace = g_pair_trace;
#endif
    int clusters = num_sms() / 2;
    if (clusters > p.items) clusters = p.items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = P_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MMADA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, p));
    return kOk;
}

}  // namespace

// head_dim 128 entry used by mmada_attention_bf16 (attention.cu); poly = MMADA_ATT_POLY (share of the exponentials on the
// FMA pipe), < 0 when not set
int launch_attention_pair(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                          int Lq, int H, float scale, int poly, cudaStream_t stream) {
    // EXPERIMENTS builds, MMADA_ATT_MAXCHK=1: take the row maxima of every tile (the previous scheme; A/B runs)
    static const int maxchk = experiment_env("MMADA_ATT_MAXCHK", 0);
    if (maxchk) {
        switch (poly) {
            case 0: return launch_pair<0, true>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
            default: return launch_pair<4, true>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        }
    }
    // default (poly < 0): every exponential on the MUFU pipe — without the row-maximum chain in the hot path the
    // polynomial's extra FMA-pipe instructions cost more than the MUFU slots they free (0.705 ms against 0.742)
    switch (poly) {
        case 3: return launch_pair<2, false>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);     // an eighth
        case 2: case 4: return launch_pair<4, false>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);     // a quarter
        default: return launch_pair<0, false>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
    }
}

}  // namespace mmada

#ifdef MMADA_ATT_TRACE
extern "C" void mmada_attention_pair_set_trace(void* buf) { mmada::g_pair_trace = (long long*)buf; }
#endif
