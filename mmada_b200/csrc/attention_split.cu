// Bidirectional flash attention, head_dim 128: persistent CTA PAIRS (cta_group::2) with the KEY TILES of every
// work item split between two softmax agents, each with its own accumulator (tcgen05 / TMEM / TMA).
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False) at
// /root/reference/models/modeling_llada.py:653-660 (no mask is ever applied).
//
// STATUS: experimental alternative to attention_pair.cu, selected with MMADA_ATT_SPLIT=1 and parity-tested in a child
// process (tests/test_kernels_gpu.py::test_attention_split_kernel_variant).  Measured at config 2 (16 x 32 heads,
// L = 1539): 0.743 ms against 0.686 ms for the pair kernel — the default stays attention_pair.cu (DESIGN.md section 4,
// lesson 7b has the clock64 breakdown: the agent's chain is exponentials ~2400 clk + hand-over / PV / next-S round trip
// ~1700 clk, no shorter than the in-phase loop it replaces).
//
// Why two agents.  attention_pair.cu gives all 8 softmax warps of a CTA the same score tile: the two warps that share
// a scheduler run in phase, so every TMEM round trip and every drain of the MUFU pipe (row sums -> overflow vote ->
// store -> hand-over) is exposed, and the MUFU and tensor pipes both sit at ~50 % (profiles/r01c_attention_pair_*:
// 2300 clk per 128 x 128 tile against 1024 clk of MUFU work and 1024 clk of MMA work).  Here the key tiles of an item
// alternate between agent 0 (warps 0-3, even tiles of the pair's stream) and agent 1 (warps 4-7, odd tiles); an agent
// is 128 threads, ONE QUERY ROW PER THREAD (tcgen05.ld.32x32b: no shuffles anywhere), and owns
//     S/P buffer a   (128 TMEM columns: the scores, then — aliased over their first 64 columns — the bf16 probabilities)
//     accumulator a  (128 TMEM columns) with its own running maximum and row sum
// so that nothing is shared between the agents and the chain of one agent
//     S(g) ready -> exponentials (32 keys at a time, speculating on the running maximum), P(g) -> PV(g) -> S(g+2) -> ...
// leaves its MUFU slots to the other agent exactly while its own MMAs run: the two warps of a scheduler are in
// anti-phase by construction.  The epilogue merges the two partial results like split-KV attention:
//     out = (O_0 2^(m_0 - m) + O_1 2^(m_1 - m)) / (l_0 2^(m_0 - m) + l_1 2^(m_1 - m)),  m = max(m_0, m_1).
// The tensor pipe executes  S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...  in issue order, which is also what makes the
// aliasing safe: S(g+2) is issued after PV(g), and P(g) is only written after S(g) completed, i.e. after PV(g-2).
//
// As in attention_pair.cu a 256 x N x 16 UMMA takes its A rows from both CTAs and half of its B operand from each, so
// every SM loads half of each K tile (64 keys) and half of each V tile (64 of the 128 head columns).
//   warps 0-3 / 4-7   softmax agents 0 / 1 (warp & 3 = TMEM lane quarter)
//   warps 8-11        epilogue: merge the two accumulators, bf16, token-major stores, one thread per query row
//   warp 12           TMA producer (each CTA loads its halves; full barriers live in the leader CTA)
//   warp 13           MMA issuer (leader CTA only)
// TMEM (per CTA): S/P buffer 0 | S/P buffer 1 | O_0 | O_1, 128 columns each.
#include <math.h>
#include <stdlib.h>

#include "attn_math.cuh"
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

namespace {

constexpr int P_THREADS = 448;
constexpr int P_EPI_WARP0 = 8, P_TMA_WARP = 12, P_MMA_WARP = 13;
constexpr int HD = 128;
constexpr unsigned SLEEP_NS = 200;                  // poll interval of the warps off the critical path
constexpr int KST = 4, VST = 4;                     // K / V ring depths
constexpr int Q_BYTES = 128 * HD * 2;               // 128 query rows (two 64-column boxes of 16 KiB)
constexpr int K_BYTES = 64 * HD * 2;                // 64 keys (two 64-column boxes of 8 KiB)
constexpr int V_BYTES = 128 * 64 * 2;               // 128 keys x 64 head columns (one box)
constexpr int Q_OFF = 0;                            // 2 buffers
constexpr int K_OFF = 2 * Q_BYTES;
constexpr int V_OFF = K_OFF + KST * K_BYTES;
constexpr int BAR_OFF = V_OFF + VST * V_BYTES;
constexpr int LBUF_OFF = BAR_OFF + 512;             // float [2 item parities][2 agents][m | l][128 rows]
constexpr int P_SMEM_BYTES = LBUF_OFF + 4096 + 1024;
constexpr int TM_S = 0, TM_O = 256;                 // S/P buffers 2 x 128 | accumulators 2 x 128 columns

struct SplitParams {
    __nv_bfloat16* out;
    int64_t ldo;
    int L, H, B;
    int q_pairs, items;
    float scale_log2;
#ifdef MMADA_ATT_TRACE
    long long* trace;
#endif
};

// clock64 timeline of CTA 0 (scripts/attn_trace.py): roles 0 / 1 = warps 0 / 4 (the two agents' warps on scheduler 0),
// 2 = MMA warp
#ifdef MMADA_ATT_TRACE
#define PTR(role, g, ev)                                                                                   \
    do {                                                                                                   \
        if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g) < 64)                             \
            p.trace[((role) * 64 + (g)) * 8 + (ev)] = clock64();                                           \
    } while (0)
#else
#define PTR(role, g, ev) do {} while (0)
#endif
#define SPTR(g, ev) do { if (quarter == 0) PTR(a, g, ev); } while (0)

// barrier indices (8 bytes each)
enum : int {
    B_QFULL = 0, B_QEMPTY = 2, B_KFULL = 4, B_KEMPTY = 4 + KST, B_VFULL = 4 + 2 * KST, B_VEMPTY = 4 + 2 * KST + VST,
    B_SFULL = 4 + 2 * KST + 2 * VST, B_PFULL = B_SFULL + 2, B_OFULL = B_PFULL + 2, B_OEMPTY = B_OFULL + 2,
    B_LFULL = B_OEMPTY + 2 /* [agent][item parity] */, B_TMEMPTR = B_LFULL + 4
};
static_assert(B_TMEMPTR * 8 + 8 <= 512, "barrier block");

__global__ void __launch_bounds__(P_THREADS, 1)
attention_split_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, const SplitParams p) {
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
    const int T = (p.L + 127) / 128;                        // key tiles per item (>= 2: the launcher checks)
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
            mbar_init(bar(B_PFULL + i), 8);      // one arrival per warp of the agent, both CTAs
            mbar_init(bar(B_OFULL + i), 1);
            mbar_init(bar(B_OEMPTY + i), 8);     // one arrival per epilogue warp of both CTAs
            mbar_init(bar(B_LFULL + 2 * i), 4);  // one arrival per warp of the agent (this CTA)
            mbar_init(bar(B_LFULL + 2 * i + 1), 4);
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
    // row maxima and row sums of the two agents for the epilogue: [item parity][agent][m | l][row]
    float* lbuf = reinterpret_cast<float*>(smem + LBUF_OFF);
    auto lidx = [](int ob, int a, int which, int row) { return ((ob * 2 + a) * 2 + which) * 128 + row; };

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
            // S(g) = Q . K^T : M = 256 (128 rows per CTA), N = keys, K = 128; operands K-major; into S/P buffer g & 1
            auto issue_s = [&](int g, int n, int j) {
                const int qb = n & 1, ks = g % KST, a = g & 1;
                if (j == 0) mbar_wait(bar(B_QFULL + qb), (n >> 1) & 1, 20);
                mbar_wait(bar(B_KFULL + ks), (g / KST) & 1, 21);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t idesc = umma_idesc_bf16(256, j == T - 1 ? tail16 : 128);
                    const uint32_t qa = (sbase + Q_OFF + qb * Q_BYTES) >> 4;
                    const uint32_t ka = (sbase + K_OFF + ks * K_BYTES) >> 4;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint32_t qoff = ((k >> 2) * (Q_BYTES / 2) + (k & 3) * 32) >> 4;
                        const uint32_t koff = ((k >> 2) * (K_BYTES / 2) + (k & 3) * 32) >> 4;
                        umma_bf16_ss<2>(tmem + TM_S + 128 * a, kdesc_hi | (uint64_t)(qa + qoff),
                                        kdesc_hi | (uint64_t)(ka + koff), idesc, k != 0);
                    }
                    umma_commit_2sm(bar(B_KEMPTY + ks), 0x3);
                    umma_commit_2sm(bar(B_SFULL + a), 0x3);
                    if (j == T - 1) umma_commit_2sm(bar(B_QEMPTY + qb), 0x3);
                }
                __syncwarp();
            };
            // O_a += P(g) . V : M = 256, N = 128 head columns (64 per CTA), K = keys; A = P in TMEM (over the scores),
            // B = V MN-major.  Tiles j = 0, 1 of an item are the first of their agents, T-2, T-1 the last.
            auto issue_pv = [&](int g, int n, int j) {
                const int vs = g % VST, a = g & 1;
                mbar_wait(bar(B_VFULL + vs), (g / VST) & 1, 22);
                if (j < 2) mbar_wait(bar(B_OEMPTY + a), (n & 1) ^ 1, 23);   // epilogue of the previous item has read O_a
                tc_fence_after();
                if (elect_one()) {
                    constexpr uint32_t idesc = umma_idesc_bf16(256, HD, 0, 1);
                    const uint32_t va = (sbase + V_OFF + vs * V_BYTES) >> 4;
                    const int ksteps = (j == T - 1 ? tail16 : 128) / 16;
                    const uint32_t acc0 = j >= 2;
                    const uint32_t d_o = tmem + TM_O + 128 * a, a_p = tmem + TM_S + 128 * a;
                    const uint64_t vd = vdesc_hi | (uint64_t)va;
                    if (ksteps == 8) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16_ts_cg<2>(d_o, a_p + 8 * k, vd + (uint64_t)(k * (2048 >> 4)), idesc, k ? 1u : acc0);
                    } else {
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16_ts_cg<2>(d_o, a_p + 8 * k, vd + (uint64_t)(k * (2048 >> 4)), idesc, k ? 1u : acc0);
                    }
                    umma_commit_2sm(bar(B_VEMPTY + vs), 0x3);
                    if (j >= T - 2) umma_commit_2sm(bar(B_OFULL + a), 0x3);
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
                PTR(2, g, 0);
                mbar_wait(bar(B_PFULL + (g & 1)), (g >> 1) & 1, 24);
                tc_fence_after();
                PTR(2, g, 1);
                issue_pv(g, np, jp);
                step(np, jp);
                PTR(2, g, 2);
                if (g + 2 < G) {                // the buffer of P(g) takes the scores of tile g+2, behind PV(g)
                    issue_s(g + 2, ns, js);
                    step(ns, js);
                }
                PTR(2, g, 3);
            }
        }
    } else if (warp >= P_EPI_WARP0) {
        // ======================================= epilogue =======================================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t oempty_lead = mapa_u32(bar(B_OEMPTY), 0);
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0;
            item_coords(n, b, h, q0);
            const int ob = n & 1;
            const int qrow = q0 + row;
            mbar_wait_backoff(bar(B_LFULL + ob), (n >> 1) & 1, 41, 4 * SLEEP_NS);
            mbar_wait_backoff(bar(B_LFULL + 2 + ob), (n >> 1) & 1, 42, 4 * SLEEP_NS);
            const float m0 = lbuf[lidx(ob, 0, 0, row)], l0 = lbuf[lidx(ob, 0, 1, row)];
            const float m1 = lbuf[lidx(ob, 1, 0, row)], l1 = lbuf[lidx(ob, 1, 1, row)];
            const float m = fmaxf(m0, m1);
            float a0 = ex2_mufu((m0 - m) * p.scale_log2), a1 = ex2_mufu((m1 - m) * p.scale_log2);
            const float inv = 1.0f / (l0 * a0 + l1 * a1);
            a0 *= inv;
            a1 *= inv;
            mbar_wait_backoff(bar(B_OFULL), n & 1, 40, SLEEP_NS);
            mbar_wait_backoff(bar(B_OFULL + 1), n & 1, 43, SLEEP_NS);
            tc_fence_after();
            // O -> registers (bf16) first, so that the accumulators are free for the next item before the stores go out
            uint32_t ow[HD / 2];
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                uint32_t u[16], w[16];
                tmem_ld_32x32b_x16(tmem + TM_O + c * 16 + lane_off, u);
                tmem_ld_32x32b_x16(tmem + TM_O + 128 + c * 16 + lane_off, w);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    ow[c * 8 + i] = pack_bf16(__uint_as_float(u[2 * i]) * a0 + __uint_as_float(w[2 * i]) * a1,
                                              __uint_as_float(u[2 * i + 1]) * a0 + __uint_as_float(w[2 * i + 1]) * a1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(oempty_lead);
                mbar_arrive_cluster(oempty_lead + 8);
            }
            if (qrow < p.L) {
                __nv_bfloat16* orow = p.out + ((int64_t)b * p.L + qrow) * p.ldo + h * HD;
#pragma unroll
                for (int u = 0; u < HD / 8; ++u)
                    *reinterpret_cast<uint4*>(orow + 8 * u) = make_uint4(ow[4 * u], ow[4 * u + 1], ow[4 * u + 2], ow[4 * u + 3]);
            }
        }
    } else {
        // ======================================= softmax agents =======================================
        // Agent a = warp / 4 takes the tiles g = a, a+2, ... of the stream; thread = query row (TMEM lane).
        const int a = warp >> 2, quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t t_s = tmem + TM_S + 128 * a + lane_off, t_o = tmem + TM_O + 128 * a + lane_off;
        const uint32_t pfull_lead = mapa_u32(bar(B_PFULL + a), 0);
        const float sl2 = p.scale_log2;
        float m = -INFINITY, l = 0.f;
        for (int g = a; g < G; g += 2) {
            const int n = g / T, j = g - n * T, ob = n & 1;
            const bool first = j < 2, last = j >= T - 2;
            const int keys = (j == T - 1) ? tail : 128;             // valid keys of this tile
            const int nch = (keys + 31) >> 5;                       // 32-key chunks that hold valid keys
            uint32_t va[32], vb[32];
            SPTR(g, 0);
            mbar_wait(bar(B_SFULL + a), (g >> 1) & 1, 30);
            tc_fence_after();
            SPTR(g, 1);
            // SPECULATE that the reference maximum m still holds (lazy rescale: it does unless the row maximum grows by
            // more than 2^8): 32 keys at a time with the next load in flight, exponentials straight away, and whether m
            // held is read off the exponentials themselves — a score more than 2^8 above m gives an exponential > 256,
            // hence a chunk sum > 256 (+inf once it overflows) — so the hot path takes no row maximum at all.  P (bf16
            // pairs) goes over score columns [0, 64): chunk c's columns [16c, 16c+16) held keys < 32(c+1), all read.
            const float2 sc2 = make_float2(sl2, sl2);
            tmem_ld_32x32b_x32(t_s, va);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < nch) {
                    uint32_t (&cur)[32] = (c & 1) ? vb : va;
                    uint32_t (&nxt)[32] = (c & 1) ? va : vb;
                    tmem_ld_wait_dep32(cur);
                    if (c + 1 < nch) tmem_ld_32x32b_x32(t_s + 32 * (c + 1), nxt);
                    if (32 * (c + 1) > keys) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (32 * c + i >= keys) cur[i] = 0xff800000u;       // -inf -> probability 0
                    }
                    auto chunk_max = [&]() {
                        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])));
                            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])));
                        }
                        return fmaxf(mx0, mx1);
                    };
                    uint32_t pw[16];
                    auto chunk_exp = [&]() {
                        const float nm = -m * sl2;
                        const float2 nm2 = make_float2(nm, nm);
                        float2 rs[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float2 x = ffma2(make_float2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), sc2, nm2);
                            const float2 e = make_float2(ex2_mufu(x.x), ex2_mufu(x.y));
                            rs[i & 1] = fadd2(rs[i & 1], e);
                            pw[i] = pack_bf16(e.x, e.y);
                        }
                        return (rs[0].x + rs[0].y) + (rs[1].x + rs[1].y);
                    };
                    const bool opening = first && c == 0;       // the agent's first keys of this item: no reference yet
                    if (opening) {
                        m = chunk_max();
                        l = 0.f;
                    }
                    float part = chunk_exp();
                    if (!opening) {
                        const bool ovf = !(part <= 256.0f);
                        if (__any_sync(0xffffffffu, ovf)) {
                            // mis-speculated: rows that overflowed move their reference to this chunk's maximum; their
                            // accumulator row (quiescent: S(g) was issued behind PV(g-2), the agent's previous tile), row
                            // sum and the probabilities already written for this tile are rescaled, the chunk is redone
                            const float mx = chunk_max();
                            const float alpha = ovf ? ex2_mufu((m - mx) * sl2) : 1.0f;
                            if (ovf) m = mx;
                            l *= alpha;
                            if (!first) {
#pragma unroll 1
                                for (int cc = 0; cc < HD / 16; ++cc) {
                                    uint32_t ov[16];
                                    tmem_ld_32x32b_x16(t_o + 16 * cc, ov);
                                    tmem_ld_wait();
#pragma unroll
                                    for (int u = 0; u < 16; ++u) ov[u] = __float_as_uint(__uint_as_float(ov[u]) * alpha);
                                    tmem_st_32x32b_x16(t_o + 16 * cc, ov);
                                }
                            }
#pragma unroll 1
                            for (int cc = 0; cc < c; ++cc) {
                                uint32_t w[16];
                                tmem_ld_32x32b_x16(t_s + 16 * cc, w);
                                tmem_ld_wait();
#pragma unroll
                                for (int u = 0; u < 16; ++u)
                                    w[u] = pack_bf16(__uint_as_float(w[u] << 16) * alpha, __uint_as_float(w[u] & 0xffff0000u) * alpha);
                                tmem_st_32x32b_x16(t_s + 16 * cc, w);
                            }
                            tmem_st_wait();
                            part = chunk_exp();
                        }
                    }
                    l += part;
                    tmem_st_32x32b_x16(t_s + 16 * c, pw);
                }
            }
            SPTR(g, 3);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(pfull_lead);
            SPTR(g, 4);
            if (last) {
                lbuf[lidx(ob, a, 0, row)] = m;
                lbuf[lidx(ob, a, 1, row)] = l;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_LFULL + 2 * a + ob));
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
long long* g_split_trace = nullptr;
#endif

}  // namespace

// head_dim 128, L > 128 entry used by mmada_attention_bf16 (attention.cu).  Query rows [0, Lq): Lq == L, or a multiple
// of 256 (the rest goes to another launch).
int launch_attention_split(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                           int Lq, int H, float scale, cudaStream_t stream) {
    if (L <= 128) return kUnsupportedShape;                   // both agents need a tile in every item
    CUtensorMap mq, mk, mv;
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
    const uint32_t box128[3] = {64, 128, 1}, box64[3] = {64, 64, 1};
    int st;
    if ((st = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box128))) return st;
    if ((st = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box64))) return st;
    if ((st = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box128))) return st;
    auto kern = attention_split_kernel;
    static bool configured = false;
    if (!configured) {
        MMADA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
        configured = true;
    }
    SplitParams p = {};
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.L = L; p.H = H; p.B = B;
    p.q_pairs = (Lq + 255) / 256;
    p.items = B * H * p.q_pairs;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef MMADA_ATT_TRACE
    p.trace = g_split_trace;
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

}  // namespace mmada

#ifdef MMADA_ATT_TRACE
extern "C" void mmada_attention_split_set_trace(void* buf) { mmada::g_split_trace = (long long*)buf; }
#endif
