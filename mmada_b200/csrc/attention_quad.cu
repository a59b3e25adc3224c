// Bidirectional flash attention, head_dim 128: persistent CTA PAIRS (cta_group::2), two 256-row query blocks per pair in
// flight, probabilities through SHARED memory (tcgen05 / TMEM / TMA).
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False) at
// /root/reference/models/modeling_llada.py:653-660 (no mask is ever applied).
//
// Why this layout (round 2; timelines and stall profiles in profiles/r02_attention_*):
//   * round 1's pair kernel (attention_pair.cu): one query tile per CTA, all eight softmax warps on the same key tile in
//     phase — 51 % tensor pipe, bound by the softmax warps' serial per-tile protocol.
//   * attention_duo.cu: two query tiles per CTA in ping-pong, P written back over S in TMEM.  P over S forces the order
//     PV(g) -> S(g+1) -> softmax(g+1) per tile: a serial chain of softmax (1900 clk) + hand-overs (350) + two MMAs (1300)
//     per slot, half of which the other slot fills: 0.67 ms, period 3600 clk against 2048 clk of tensor work.
//   * here P goes to SHARED memory (a K-major SW128 A operand, like Q), so the score buffer of a slot is free as soon as
//     the softmax threads have READ it: S(g+1) runs while softmax(g) is still computing, and the softmax warps of both
//     slots work back to back without ever waiting for their own MMAs.  What that costs is shared memory (64 KB of Q +
//     64 KB of P per CTA), affordable only because the two CTAs of a pair split every K and V tile (16 KB each per tile
//     and CTA: a 256-row UMMA takes half of its B operand from each CTA).
//   A pair walks work items (batch, head, 512 query rows): slot s of CTA r holds rows q0 + 256 s + 128 r.  TMEM per CTA:
//   S0 | S1 | O0 | O1, 128 columns each.  Per key tile g the leader's MMA warp issues
//       S_s(g+1) as soon as the 8 softmax warps of slot s (both CTAs) have loaded S_s(g)      [both operands in smem]
//       PV_s(g)  as soon as they have stored P_s(g)                                            [A = P in smem, B = V MN-major]
//   warps 0-7 / 8-15: softmax + epilogue of slot 0 / 1.  A query row belongs to TWO threads (same lane of warps w and w+4:
//   tcgen05.ld.32x32b, TMEM lane = row), each with 64 of the tile's 128 scores in registers (104 registers after
//   setmaxnreg): with both slots busy every scheduler has FOUR softmax warps to pick from, which is what it takes to cover
//   the fixed-latency dependencies of the exponential code (with one or two warps per scheduler the same instructions ran
//   at 0.36-0.45 IPC: stall_wait).  Speculative exponentials against the running reference maximum (lazy rescale: it holds
//   unless a score exceeds it by more than 2^8): no row-maximum pass in front; whether it held is read off the row sums
//   with one 64-thread named-barrier reduction (bar.red.or) per tile, and a mis-speculated tile is redone from the
//   registers with the row maximum exchanged through shared memory.  P as bf16 to shared memory, O / l with 256-bit stores
//   at the end of an item.
//   warp 16: TMA producer (each CTA loads its halves; "full" barriers live in the leader CTA), warp 17: MMA issuer (leader),
//   warps 18-19 idle (they complete the fifth warpgroup, which hands registers to the softmax warpgroups).
#include <math.h>

#include "attn_math.cuh"
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

#ifndef QUAD_SCALAR
#define QUAD_SCALAR 0
#endif

namespace mmada {

namespace {

constexpr int Q_THREADS = 640;                        // 16 softmax warps + one warpgroup with the TMA and MMA warps
constexpr int Q_TMA_WARP = 16, Q_MMA_WARP = 17;
// setmaxnreg moves registers inside the CTA's launch allocation (640 x 96): 128 x (96 - 64) released = 512 x (104 - 96) claimed
constexpr int REGS_SOFTMAX = 104, REGS_OTHER = 64;
constexpr int HD = 128;
constexpr int KST = 3, VST = 2;                       // K / V ring depths (K is consumed a period ahead of V)
constexpr int QT_BYTES = 128 * HD * 2;                // a 128-row Q or P tile: two 64-column boxes of 16 KiB
constexpr int BOX_BYTES = 128 * 64 * 2;
constexpr int K_BYTES = 64 * HD * 2;                  // this CTA's 64 keys of a tile: two 64-column boxes of 8 KiB
constexpr int V_BYTES = 128 * 64 * 2;                 // 128 keys x this CTA's 64 head columns: one box
constexpr int Q_OFF = 0;                              // 2 slots
constexpr int P_OFF = 2 * QT_BYTES;                   // 2 slots
constexpr int K_OFF = P_OFF + 2 * QT_BYTES;
constexpr int V_OFF = K_OFF + KST * K_BYTES;
constexpr int BAR_OFF = V_OFF + VST * V_BYTES;
constexpr int MXBUF_OFF = BAR_OFF + 512;              // float [2 slots][2 tile parities][2 halves][128 rows]: half-tile row maxima
constexpr int LBUF_OFF = MXBUF_OFF + 4096;            // float [2 slots][2 halves][128 rows]: partial row sums at the end of an item
constexpr int Q_SMEM_BYTES = LBUF_OFF + 2048 + 1024;
static_assert(Q_SMEM_BYTES <= 232448, "shared memory");
constexpr int TM_S = 0, TM_O = 256;                   // S_s at TM_S + 128 s, O_s at TM_O + 128 s

enum : int {
    B_QFULL = 0, B_QEMPTY = 2, B_KFULL = 4, B_KEMPTY = 4 + KST, B_VFULL = 4 + 2 * KST, B_VEMPTY = 4 + 2 * KST + VST,
    B_SFULL = 4 + 2 * KST + 2 * VST, B_SFREE = B_SFULL + 2, B_PFULL = B_SFREE + 2, B_PVDONE = B_PFULL + 2,
    B_OFULL = B_PVDONE + 2, B_TMEMPTR = B_OFULL + 2
};
static_assert(B_TMEMPTR * 8 + 8 <= 512, "barrier block");

struct QuadParams {
    __nv_bfloat16* out;
    int64_t ldo;
    int L, H, B;
    int Lq;                // query rows [0, Lq) are written here
    int q_blocks, items;   // 512-row blocks per (batch, head); work items
    float scale_log2;
#ifdef MMADA_ATT_TRACE
    long long* trace;
#endif
};

#ifdef MMADA_ATT_TRACE
#define QTR(role, g, ev)                                                                                   \
    do {                                                                                                   \
        if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g) < 64)                             \
            p.trace[((role) * 64 + (g)) * 8 + (ev)] = clock64();                                           \
    } while (0)
#else
#define QTR(role, g, ev) do {} while (0)
#endif

// wait on a barrier of this CTA that CTAs of the cluster arrive on (acquire at cluster scope), call-free
__device__ __forceinline__ void mbar_wait_cluster_nocall(uint32_t bar, uint32_t parity) {
    uint32_t n = 0, done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (!done && ++n > (1u << 22)) __trap();
    } while (!done);
}
// named barrier over 64 threads that also ORs a predicate across them
__device__ __forceinline__ bool bar64_red_or(int id, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %2, 0;\n\tbar.red.or.pred q, %1, 64, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(id), "r"((uint32_t)pred)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x64(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
          "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
          "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
          "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
          "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// POLY = how many of every 8 key pairs take the polynomial 2^x (FMA pipe) instead of MUFU.EX2
template <int POLY>
__global__ void __launch_bounds__(Q_THREADS, 1)
attention_quad_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const QuadParams p) {
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

    if (warp == Q_TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_QFULL + i), 1);
            mbar_init(bar(B_QEMPTY + i), 1);
            mbar_init(bar(B_SFULL + i), 1);
            mbar_init(bar(B_SFREE + i), 16);     // one arrival per softmax warp of the slot, both CTAs
            mbar_init(bar(B_PFULL + i), 16);
            mbar_init(bar(B_PVDONE + i), 1);
            mbar_init(bar(B_OFULL + i), 1);
        }
        for (int s = 0; s < KST; ++s) { mbar_init(bar(B_KFULL + s), 1); mbar_init(bar(B_KEMPTY + s), 1); }
        for (int s = 0; s < VST; ++s) { mbar_init(bar(B_VFULL + s), 1); mbar_init(bar(B_VEMPTY + s), 1); }
        fence_mbar_init();
    }
    if (warp == Q_MMA_WARP) {
        tmem_alloc<2>(bar(B_TMEMPTR), 512);
        tmem_relinquish<2>();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;

    // item n of this pair -> (batch, head, first query row of the 512-row block)
    auto item_coords = [&](int n, int& b, int& h, int& q0) {
        const int id = cluster_id + n * num_clusters;
        const int qb = id % p.q_blocks, bh = id / p.q_blocks;
        h = bh % p.H;
        b = bh / p.H;
        q0 = qb * 512;
    };

    // setmaxnreg sits at the head of every role's branch: ptxas allocates a region with the count of the setmaxnreg that
    // dominates it, and falls back to the kernel-wide cap where paths with different counts merge
    if (warp == Q_TMA_WARP) {
        // ======================================= TMA producer =======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
        const uint32_t lead0 = mapa_u32(bar(0), 0);               // the leader CTA's barrier block
        auto lbar = [&](int idx) { return lead0 + 8 * idx; };
        // order per item: K tile 0, the Q tiles, V tile 0, then K/V tiles 1.. — the next item's first K tile is in flight
        // before its Q tiles have to wait for the current item's last score MMAs
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0;
            item_coords(n, b, h, q0);
            for (int j = 0; j < T; ++j) {
                const int g = n * T + j;
                const int ks = g % KST, vs = g % VST;
                mbar_wait_backoff(bar(B_KEMPTY + ks), ((g / KST) & 1) ^ 1, 11, 100);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(bar(B_KFULL + ks), 2 * K_BYTES);
                    // this CTA's half of the N keys the score MMA covers (N = 128, or tail16 in the last tile)
                    const int half_n = (j == T - 1 ? tail16 : 128) / 2;
                    for (int c = 0; c < 2; ++c)
                        tma_load_3d_2sm(sbase + K_OFF + ks * K_BYTES + c * (K_BYTES / 2), &map_k, lbar(B_KFULL + ks),
                                        h * HD + c * 64, j * 128 + (int)rank * half_n, b, kEvictLast);
                }
                __syncwarp();
                if (j == 0) {
                    for (int s = 0; s < 2; ++s) {
                        mbar_wait_backoff(bar(B_QEMPTY + s), (n & 1) ^ 1, 10, 100);
                        if (elect_one()) {
                            if (leader) mbar_arrive_expect_tx(bar(B_QFULL + s), 2 * QT_BYTES);
                            int qrow = q0 + 256 * s + 128 * (int)rank;
                            if (qrow >= p.L) qrow = 0;            // a tile entirely past the end: any rows, never stored
                            for (int c = 0; c < 2; ++c)
                                tma_load_3d_2sm(sbase + Q_OFF + s * QT_BYTES + c * BOX_BYTES, &map_q, lbar(B_QFULL + s),
                                                h * HD + c * 64, qrow, b, kEvictFirst);
                        }
                        __syncwarp();
                    }
                }
                mbar_wait_backoff(bar(B_VEMPTY + vs), ((g / VST) & 1) ^ 1, 12, 100);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(bar(B_VFULL + vs), 2 * V_BYTES);
                    tma_load_3d_2sm(sbase + V_OFF + vs * V_BYTES, &map_v, lbar(B_VFULL + vs), h * HD + (int)rank * 64,
                                    j * 128, b, kEvictLast);
                }
                __syncwarp();
            }
        }
    } else if (warp == Q_MMA_WARP) {
        // ======================================= MMA issuer (leader) =======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
        if (leader && G > 0) {
            const uint64_t kdesc_hi = umma_desc_kmajor_sw128(0);
            const uint64_t vdesc_hi = umma_desc_mnmajor_sw128(0, V_BYTES);
            // S_s(g) = Q_s . K_g^T : M = 256 (128 rows per CTA), N = keys, K = 128; both operands K-major in shared memory
            auto issue_s = [&](int s, int g, int n, int j) {
                const int ks = g % KST;
                if (j == 0) mbar_wait(bar(B_QFULL + s), n & 1, 20);
                mbar_wait(bar(B_KFULL + ks), (g / KST) & 1, 21);
                tc_fence_after();
                QTR(2 + s, g, 0);
                if (elect_one()) {
                    const uint32_t idesc = umma_idesc_bf16(256, j == T - 1 ? tail16 : 128);
                    const uint32_t qa = (sbase + Q_OFF + s * QT_BYTES) >> 4;
                    const uint32_t ka = (sbase + K_OFF + ks * K_BYTES) >> 4;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint32_t qoff = ((k >> 2) * BOX_BYTES + (k & 3) * 32) >> 4;
                        const uint32_t koff = ((k >> 2) * (K_BYTES / 2) + (k & 3) * 32) >> 4;
                        umma_bf16_ss<2>(tmem + TM_S + 128 * s, kdesc_hi | (uint64_t)(qa + qoff), kdesc_hi | (uint64_t)(ka + koff),
                                        idesc, k != 0);
                    }
                    umma_commit_2sm(bar(B_SFULL + s), 0x3);
                    if (s == 1) umma_commit_2sm(bar(B_KEMPTY + ks), 0x3);
                    if (j == T - 1) umma_commit_2sm(bar(B_QEMPTY + s), 0x3);
                }
                __syncwarp();
                QTR(2 + s, g, 1);
            };
            // O_s (+)= P_s(g) . V_g : M = 256, N = 128 head columns (64 per CTA), K = keys; A = P (K-major, shared memory),
            // B = V MN-major
            auto issue_pv = [&](int s, int g, int j) {
                const int vs = g % VST;
                QTR(2 + s, g, 2);
                mbar_wait_cluster_nocall(bar(B_PFULL + s), g & 1);
                mbar_wait(bar(B_VFULL + vs), (g / VST) & 1, 22);
                tc_fence_after();
                QTR(2 + s, g, 3);
                if (elect_one()) {
                    constexpr uint32_t idesc = umma_idesc_bf16(256, HD, 0, 1);
                    const uint32_t pa = (sbase + P_OFF + s * QT_BYTES) >> 4;
                    const uint32_t va = (sbase + V_OFF + vs * V_BYTES) >> 4;
                    const int ksteps = (j == T - 1 ? tail16 : 128) / 16;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint32_t poff = ((k >> 2) * BOX_BYTES + (k & 3) * 32) >> 4;
                        umma_bf16_ss<2>(tmem + TM_O + 128 * s, kdesc_hi | (uint64_t)(pa + poff),
                                        vdesc_hi | (uint64_t)(va + k * (2048 >> 4)), idesc, (j | k) != 0);
                    }
                    umma_commit_2sm(bar(B_PVDONE + s), 0x3);
                    if (s == 1) umma_commit_2sm(bar(B_VEMPTY + vs), 0x3);
                    if (j == T - 1) umma_commit_2sm(bar(B_OFULL + s), 0x3);
                }
                __syncwarp();
                QTR(2 + s, g, 4);
            };
            issue_s(0, 0, 0, 0);
            issue_s(1, 0, 0, 0);
            int n = 0, j = 0;                                     // (item, tile) of g
            for (int g = 0; g < G; ++g) {
                int n1 = n, j1 = j + 1;                           // (item, tile) of g + 1
                if (j1 == T) { j1 = 0; ++n1; }
                if (g + 1 < G) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        // the softmax threads of slot s (both CTAs) hold tile g's scores in registers: the buffer is free
                        mbar_wait_cluster_nocall(bar(B_SFREE + s), g & 1);
                        tc_fence_after();
                        issue_s(s, g + 1, n1, j1);
                    }
                }
#pragma unroll
                for (int s = 0; s < 2; ++s) issue_pv(s, g, j);
                n = n1;
                j = j1;
            }
        }
    } else if (warp < 16) {
        // ======================================= softmax + epilogue =======================================
        // (no out-of-line call in this region: ptxas only honours setmaxnreg.inc for call-free code)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
        const int s = warp >> 3;                        // slot
        const int half = (warp >> 2) & 1;               // which 64 keys of every tile (and which 64 output columns)
        const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t t_s = tmem + TM_S + 128 * s + 64 * half + lane_off;      // this thread's 64 scores
        const uint32_t t_o = tmem + TM_O + 128 * s + 64 * half + lane_off;      // its 64 output columns
        const int row = quarter * 32 + lane;
        // this row's 64 keys of the P tile in shared memory = box `half` (K-major, 128-byte swizzle): 16-byte chunk c of
        // the box sits at chunk position c ^ (row & 7)
        const uint32_t p_row = sbase + P_OFF + s * QT_BYTES + half * BOX_BYTES + row * 128;
        const uint32_t p_xor = (uint32_t)(row & 7);
        const uint32_t sfree_lead = mapa_u32(bar(B_SFREE + s), 0), pfull_lead = mapa_u32(bar(B_PFULL + s), 0);
        const float sl2 = p.scale_log2;
        float* mxbuf = reinterpret_cast<float*>(smem + MXBUF_OFF) + s * 512;   // [tile parity][half][row]
        float* lbuf = reinterpret_cast<float*>(smem + LBUF_OFF) + s * 256;     // [half][row]
        const int pair_bar = 1 + s * 4 + quarter;       // named barrier of the two warps that share these 32 rows
        const bool tr = quarter == 0 && half == 0;
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0;
            item_coords(n, b, h, q0);
            float m_used = -INFINITY, l_sum = 0.f;
            for (int j = 0; j < T; ++j) {
                const int g = n * T + j;
                if (tr) QTR(s, g, 0);
                mbar_wait_nocall(bar(B_SFULL + s), g & 1);
                tc_fence_after();
                if (tr) QTR(s, g, 1);
                // the row maximum over both halves: through shared memory and the pair's named barrier (first tile of
                // an item, mis-speculated tile, last partial tile only)
                auto row_max_exchange = [&](float mx) {
                    float* mxb = mxbuf + (g & 1) * 256;
                    mxb[half * 128 + row] = mx;
                    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                    return fmaxf(mx, mxb[(half ^ 1) * 128 + row]);
                };
                // rescale of this thread's half of the accumulator row and of its partial row sum (rare)
                auto rescale = [&](float m_new, bool grow) {
                    const float alpha = grow ? ex2_mufu((m_used - m_new) * sl2) : 1.0f;
                    if (grow) m_used = m_new;
                    l_sum *= alpha;
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        uint32_t ov[16];
                        tmem_ld_32x32b_x16(t_o + cc * 16, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 16; ++u) ov[u] = __float_as_uint(__uint_as_float(ov[u]) * alpha);
                        tmem_st_32x32b_x16(t_o + cc * 16, ov);
                    }
                    tmem_st_wait();
                };
                // 16 keys of P (8 bf16 pairs) of this row -> shared memory; c = chunk inside this thread's 64 keys
                auto store_p = [&](int c, const uint32_t (&pw)[8]) {
                    const uint32_t c0 = (uint32_t)(2 * c);
                    st_shared_v4(p_row + ((c0 ^ p_xor) << 4), pw[0], pw[1], pw[2], pw[3]);
                    st_shared_v4(p_row + (((c0 + 1) ^ p_xor) << 4), pw[4], pw[5], pw[6], pw[7]);
                };
                if (j == T - 1 && tail < 128) {
                    // ---- the last, partial tile of the sequence (once per item): a compact two-pass loop over the 16-key
                    // chunks of this half that the MMAs cover, scores re-read from TMEM, padding keys masked, no speculation
                    const int keys_h = tail - 64 * half;                              // valid keys of this half (may be <= 0)
                    const int k16_h = tail16 - 64 * half;
                    const int nch = k16_h <= 0 ? 0 : (k16_h >= 64 ? 4 : (k16_h >> 4));
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int ch = 0; ch < nch; ++ch) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(t_s + 16 * ch, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (16 * ch + i < keys_h) mx = fmaxf(mx, __uint_as_float(v[i]));
                    }
                    // PV_s of the previous tile has read P_s (and, for a rescale, O_s is quiescent)
                    if (g > 0) mbar_wait_nocall(bar(B_PVDONE + s), (g - 1) & 1);
                    tc_fence_after();
                    const float m_new = fmaxf(m_used, row_max_exchange(mx));
                    if (j == 0) {
                        m_used = m_new;
                    } else {
                        const bool grow = (m_new - m_used) * sl2 > 8.0f;
                        if (__any_sync(0xffffffffu, grow)) rescale(m_new, grow);     // same rows: the partner warp agrees
                    }
                    const float mb = m_used * sl2;
                    float part = 0.f;
#pragma unroll 1
                    for (int ch = 0; ch < nch; ++ch) {
                        uint32_t v[16], pw[8];
                        tmem_ld_32x32b_x16(t_s + 16 * ch, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float e0 = (16 * ch + 2 * u < keys_h) ? ex2_mufu(__uint_as_float(v[2 * u]) * sl2 - mb) : 0.f;
                            const float e1 = (16 * ch + 2 * u + 1 < keys_h) ? ex2_mufu(__uint_as_float(v[2 * u + 1]) * sl2 - mb) : 0.f;
                            part += e0 + e1;
                            pw[u] = pack_bf16(e0, e1);
                        }
                        store_p(ch, pw);
                    }
                    l_sum += part;
                    // the scores have been read for the last time: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(sfree_lead);
                } else {
                    // ---- a whole tile: this thread's 64 scores in registers, one basic block of 32 key pairs
                    uint32_t sv[64];
                    tmem_ld_32x32b_x64(t_s, sv);
                    tmem_ld_wait();
                    // the scores are in registers: hand the buffer back (S of the next tile may overwrite it)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(sfree_lead);
                    if (tr) QTR(s, g, 2);
                    // PV_s of the previous tile has read P_s (and, for a rescale, O_s is quiescent): long done, normally
                    if (g > 0) mbar_wait_nocall(bar(B_PVDONE + s), (g - 1) & 1);
                    tc_fence_after();
                    bool need_max = (j == 0);                           // no reference maximum yet
#pragma unroll 1
                    for (;;) {
                        if (need_max) {
                            float mxa[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) mxa[u] = fmaxf(__uint_as_float(sv[2 * u]), __uint_as_float(sv[2 * u + 1]));
#pragma unroll
                            for (int i = 8; i < 64; i += 8)
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    mxa[u] = fmaxf(mxa[u], fmaxf(__uint_as_float(sv[i + 2 * u]), __uint_as_float(sv[i + 2 * u + 1])));
                            const float mx = row_max_exchange(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])));
                            const float m_new = fmaxf(m_used, mx);
                            if (j == 0) {
                                m_used = m_new;
                            } else {
                                const bool grow = m_new > m_used;
                                if (__any_sync(0xffffffffu, grow)) rescale(m_new, grow);
                            }
                        }
                        const float mb = m_used * sl2;
                        const float2 sc2 = make_float2(sl2, sl2), nmb2 = make_float2(-mb, -mb);
                        float2 rs2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            uint32_t pw[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int i0 = 16 * ch + 2 * u;
                                const bool poly = POLY > 0 && ((u + 1) * POLY / 8 != u * POLY / 8);
#if QUAD_SCALAR
                                float2 e;
                                if (poly) {
                                    e = ex2_poly2(ffma2(make_float2(__uint_as_float(sv[i0]), __uint_as_float(sv[i0 + 1])), sc2, nmb2));
                                } else {
                                    e.x = ex2_mufu(fmaf(__uint_as_float(sv[i0]), sl2, -mb));
                                    e.y = ex2_mufu(fmaf(__uint_as_float(sv[i0 + 1]), sl2, -mb));
                                }
                                rs2[u & 3].x += e.x;
                                rs2[u & 3].y += e.y;
#else
                                const float2 x = ffma2(make_float2(__uint_as_float(sv[i0]), __uint_as_float(sv[i0 + 1])), sc2, nmb2);
                                const float2 e = poly ? ex2_poly2(x) : make_float2(ex2_mufu(x.x), ex2_mufu(x.y));
                                rs2[u & 3] = fadd2(rs2[u & 3], e);
#endif
                                pw[u] = pack_bf16(e.x, e.y);
                            }
                            store_p(ch, pw);
                        }
                        const float part = (rs2[0].x + rs2[0].y) + (rs2[1].x + rs2[1].y) + (rs2[2].x + rs2[2].y) + (rs2[3].x + rs2[3].y);
                        if (need_max) {                     // exponentials against the true maximum: nothing to check
                            l_sum += part;
                            break;
                        }
                        // SPECULATION check: an exponential above 2^8 (or a saturated polynomial: its argument is clamped to
                        // 2^127) shows in the row sum.  One named-barrier reduction tells both warps of the rows whether either
                        // half overflowed anywhere.
                        if (!bar64_red_or(pair_bar, !(part <= 256.0f))) {
                            l_sum += part;
                            break;
                        }
                        need_max = true;
                    }
                }
                if (tr) QTR(s, g, 4);
                // P_s(g) is in shared memory: make it visible to the tensor core (async proxy), then tell the leader
                fence_proxy_async();
                __syncwarp();
                if (tr) QTR(s, g, 3);
                if (lane == 0) mbar_arrive_cluster(pfull_lead);
                if (tr) QTR(s, g, 5);
            }
            // ---- epilogue: O / l -> bf16, token-major; this thread writes its 64 columns of the row.  The next item's first
            // PV waits for this slot's next P, which these threads only produce after the loads below have completed.
            lbuf[half * 128 + row] = l_sum;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const float inv = 1.0f / (l_sum + lbuf[(half ^ 1) * 128 + row]);
            mbar_wait_nocall(bar(B_OFULL + s), n & 1);
            tc_fence_after();
            if (tr) QTR(s, n * T + T - 1, 6);
            const int qrow = q0 + 256 * s + 128 * (int)rank + row;
            __nv_bfloat16* orow = p.out + ((int64_t)b * p.L + qrow) * p.ldo + h * HD + 64 * half;
            {
                uint32_t ov[64];
                tmem_ld_32x32b_x64(t_o, ov);
                tmem_ld_wait();
                if (qrow < p.Lq) {
                    // 256-bit stores: every lane writes whole 32-byte sectors (a lane owns a row)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            w[e] = pack_bf16(__uint_as_float(ov[16 * u + 2 * e]) * inv, __uint_as_float(ov[16 * u + 2 * e + 1]) * inv);
                        st_global_v8(orow + 16 * u, w);
                    }
                }
            }
            // (lbuf is written again a whole item later, after the bar.sync of the next item's first tile)
            tc_fence_before();
            if (tr) QTR(s, n * T + T - 1, 7);
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));      // idle warps of the third warpgroup
    }
    // teardown: everyone done with TMEM, and the peer done with our shared memory / barriers
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == Q_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc<2>(tmem, 512);
    }
}

#ifdef MMADA_ATT_TRACE
long long* g_quad_trace = nullptr;
#endif

template <int POLY>
int launch_quad(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L, int Lq, int H,
                float scale, cudaStream_t stream) {
    CUtensorMap mq, mk, mv;
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
    const uint32_t box128[3] = {64, 128, 1}, box64[3] = {64, 64, 1};
    int st;
    if ((st = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box128))) return st;
    if ((st = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box64))) return st;
    if ((st = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box128))) return st;
    auto kern = attention_quad_kernel<POLY>;
    static bool configured[kMaxDevices] = {};
    MMADA_CUDA_TRY(ensure_dynamic_smem(kern, Q_SMEM_BYTES, configured));
    QuadParams p = {};
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.L = L; p.H = H; p.B = B;
    p.Lq = Lq;
    p.q_blocks = (Lq + 511) / 512;
    p.items = B * H * p.q_blocks;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef MMADA_ATT_TRACE
    p.trace = g_quad_trace;
#endif
    int clusters = num_sms() / 2;
    if (clusters > p.items) clusters = p.items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(Q_THREADS);
    cfg.dynamicSmemBytes = Q_SMEM_BYTES;
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

// head_dim 128 entry used by mmada_attention_bf16 (attention.cu): query rows [0, Lq) of every (batch, head)
int launch_attention_quad(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                          int Lq, int H, float scale, int poly, cudaStream_t stream) {
    switch (poly) {
        case 0: return launch_quad<0>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 3: return launch_quad<3>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 4: return launch_quad<4>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        default: return launch_quad<2>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
    }
}

}  // namespace mmada

#ifdef MMADA_ATT_TRACE
extern "C" void mmada_attention_quad_set_trace(void* buf) { mmada::g_quad_trace = (long long*)buf; }
#endif
