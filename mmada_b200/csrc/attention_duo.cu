// Bidirectional flash attention, head_dim 128: persistent single-CTA kernel, TWO 128-row query tiles per CTA in
// ping-pong, one thread per query row (tcgen05 / TMEM / TMA).
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False) at
// /root/reference/models/modeling_llada.py:653-660 (no mask is ever applied).
//
// Layout (round 2; the round-1 pair kernel, attention_pair.cu, sat at 51 % tensor pipe because all eight softmax warps
// of a CTA worked on the SAME tile in phase: per tile they paid the serial protocol — barrier wake-up, TMEM load,
// store, hand-over — with nothing else to run on their schedulers):
//   * a work item = (batch, head, 256 query rows) = two query tiles ("slots") that share every K / V tile in shared
//     memory.  TMEM: S0 | S1 (128 fp32 columns each; P, bf16, is written back over the first 64 columns of its own S)
//     | O0 | O1 (128 columns each) = all 512 columns.
//   * warps 0-3 own slot 0, warps 4-7 slot 1: one THREAD per query row (tcgen05.ld.32x32b, TMEM lane = row) with all 128
//     keys of the tile in its registers: exponentials, row sum and the rare rescale never leave the thread — no shuffles,
//     no votes, no shared memory.  The exponentials SPECULATE that the running reference maximum still holds (lazy
//     rescale: it does unless a score exceeds it by more than 2^8) and start as soon as the scores arrive; whether it
//     held is read off the row sum afterwards, and a mis-speculated tile is redone from the registers.  The two warps
//     that share a scheduler belong to different slots and run half a period apart.  (A variant with 16 softmax warps,
//     two threads per row, needed 96-register threads and lost to spills: profiles/r02_attention_duo_timelines.txt.)
//   * the MMA warp issues, per key tile g and slot s:  PV_s(g-1) ; S_s(g)  — the tensor pipe executes in order, so
//     S_s(g) may overwrite P_s(g-1) without a barrier; while slot 0 is in its softmax the pipe runs slot 1's pair.
//     Steady state: period = softmax latency + 1024 clk, against 2048 clk of tensor work per period.
//   * a compile-time share of the exponentials runs as a polynomial on the FMA pipe: the MUFU pipe (16 ex2/clk/SM)
//     needs as long for two 128x128 tiles as the tensor pipe needs for their four MMAs.
//   * persistent: the key-tile stream runs across item boundaries (the next item's first K tile and Q tiles are
//     loaded while the current item's last tiles are still in their softmax); the softmax threads also write their
//     row of O (scaled by 1/l) at the end of an item.
//   warps 0-7 softmax + epilogue (200 registers after setmaxnreg), warp 8 TMA producer, warp 9 MMA issuer, warps 10-11
//   idle (they complete the third warpgroup, which hands registers to the softmax warpgroups).
#include <math.h>

#include <type_traits>

#include "attn_math.cuh"
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

namespace {

constexpr int D_THREADS = 384;
constexpr int D_TMA_WARP = 8, D_MMA_WARP = 9;
constexpr int REGS_SOFTMAX = 200, REGS_OTHER = 104;  // setmaxnreg: 256 x 200 + 128 x 104 = 64512 <= 65536
constexpr int HD = 128;
constexpr int TILE_BYTES = 128 * HD * 2;            // one Q / K / V tile (two 64-column boxes of 16 KiB)
constexpr int BOX_BYTES = 128 * 64 * 2;
constexpr int KST = 2, VST = 2;                     // K / V ring depths
constexpr int Q_OFF = 0;                            // 2 slots
constexpr int K_OFF = 2 * TILE_BYTES;
constexpr int V_OFF = K_OFF + KST * TILE_BYTES;
constexpr int BAR_OFF = V_OFF + VST * TILE_BYTES;
constexpr int D_SMEM_BYTES = BAR_OFF + 512 + 1024;
constexpr int TM_S = 0, TM_O = 256;                 // S_s at TM_S + 128 s (P_s over its first 64 columns), O_s at TM_O + 128 s

enum : int {
    B_QFULL = 0, B_QEMPTY = 2, B_KFULL = 4, B_KEMPTY = 4 + KST, B_VFULL = 4 + 2 * KST, B_VEMPTY = 4 + 2 * KST + VST,
    B_SFULL = 4 + 2 * KST + 2 * VST, B_PFULL = B_SFULL + 2, B_PVDONE = B_PFULL + 2, B_OFULL = B_PVDONE + 2,
    B_TMEMPTR = B_OFULL + 2
};
static_assert(B_TMEMPTR * 8 + 8 <= 512, "barrier block");

struct DuoParams {
    __nv_bfloat16* out;
    int64_t ldo;
    int L, H, B;
    int Lq;                // query rows [0, Lq) are handled here
    int q_pairs, items;
    float scale_log2;
#ifdef MMADA_ATT_TRACE
    long long* trace;
#endif
};

#ifdef MMADA_ATT_TRACE
#define DTR(role, g, ev)                                                                                   \
    do {                                                                                                   \
        if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g) < 64)                             \
            p.trace[((role) * 64 + (g)) * 8 + (ev)] = clock64();                                           \
    } while (0)
#else
#define DTR(role, g, ev) do {} while (0)
#endif

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
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

// tcgen05.wait::ld that carries a data dependency on 16 registers an earlier tcgen05.ld fills: their uses cannot be
// scheduled above the wait when other code sits between the load and the wait (software-pipelined loads)
__device__ __forceinline__ void tmem_ld_wait_dep16(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}
// named barrier over `64` threads that also ORs a predicate across them
__device__ __forceinline__ bool bar64_red_or(int id, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %2, 0;\n\tbar.red.or.pred q, %1, 64, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(id), "r"((uint32_t)pred)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// POLY = how many of every 8 key pairs take the polynomial 2^x (FMA pipe) instead of MUFU.EX2
template <int POLY>
__global__ void __launch_bounds__(D_THREADS, 1)
attention_duo_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const DuoParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    auto bar = [&](int idx) { return sbase + BAR_OFF + 8 * idx; };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 8 * B_TMEMPTR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int T = (p.L + 127) / 128;                        // key tiles per item
    const int tail = p.L - (T - 1) * 128;                   // valid keys in the last tile (1..128)
    const int tail16 = (tail + 15) & ~15;

    if (warp == D_TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_QFULL + i), 1);
            mbar_init(bar(B_QEMPTY + i), 1);
            mbar_init(bar(B_SFULL + i), 1);
            mbar_init(bar(B_PFULL + i), 4);      // one arrival per softmax warp of the slot
            mbar_init(bar(B_PVDONE + i), 1);
            mbar_init(bar(B_OFULL + i), 1);
        }
        for (int s = 0; s < KST; ++s) { mbar_init(bar(B_KFULL + s), 1); mbar_init(bar(B_KEMPTY + s), 1); }
        for (int s = 0; s < VST; ++s) { mbar_init(bar(B_VFULL + s), 1); mbar_init(bar(B_VEMPTY + s), 1); }
        fence_mbar_init();
    }
    if (warp == D_MMA_WARP) {
        tmem_alloc<1>(bar(B_TMEMPTR), 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;

    // item n of this CTA -> (batch, head, first query row, number of query tiles with rows below Lq)
    auto item_coords = [&](int n, int& b, int& h, int& q0, int& nqt) {
        const int id = (int)blockIdx.x + n * (int)gridDim.x;
        const int qp = id % p.q_pairs, bh = id / p.q_pairs;
        h = bh % p.H;
        b = bh / p.H;
        q0 = qp * 256;
        nqt = (q0 + 128 < p.Lq) ? 2 : 1;
    };

    // setmaxnreg sits at the head of every role's branch: ptxas allocates a region with the count of the setmaxnreg that
    // dominates it, and falls back to the kernel-wide cap where paths with different counts merge
    if (warp == D_TMA_WARP) {
        // ======================================= TMA producer =======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
        // order per item: K tile 0, the Q tiles, V tile 0, then K/V tiles 1.. — the next item's first K tile is in flight
        // before its Q tiles have to wait for the current item's last score MMAs
        int qi[2] = {0, 0};                                   // items loaded per slot
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0, nqt;
            item_coords(n, b, h, q0, nqt);
            for (int j = 0; j < T; ++j) {
                const int g = n * T + j;
                const int ks = g % KST, vs = g % VST;
                mbar_wait_backoff(bar(B_KEMPTY + ks), ((g / KST) & 1) ^ 1, 11, 100);
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar(B_KFULL + ks), TILE_BYTES);
                    for (int c = 0; c < 2; ++c)
                        tma_load_3d(sbase + K_OFF + ks * TILE_BYTES + c * BOX_BYTES, &map_k, bar(B_KFULL + ks),
                                    h * HD + c * 64, j * 128, b, kEvictLast);
                }
                __syncwarp();
                if (j == 0) {
                    for (int s = 0; s < nqt; ++s) {
                        mbar_wait_backoff(bar(B_QEMPTY + s), (qi[s] & 1) ^ 1, 10, 100);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(bar(B_QFULL + s), TILE_BYTES);
                            for (int c = 0; c < 2; ++c)
                                tma_load_3d(sbase + Q_OFF + s * TILE_BYTES + c * BOX_BYTES, &map_q, bar(B_QFULL + s),
                                            h * HD + c * 64, q0 + s * 128, b, kEvictFirst);
                        }
                        __syncwarp();
                        ++qi[s];
                    }
                }
                mbar_wait_backoff(bar(B_VEMPTY + vs), ((g / VST) & 1) ^ 1, 12, 100);
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar(B_VFULL + vs), TILE_BYTES);
                    for (int c = 0; c < 2; ++c)
                        tma_load_3d(sbase + V_OFF + vs * TILE_BYTES + c * BOX_BYTES, &map_v, bar(B_VFULL + vs),
                                    h * HD + c * 64, j * 128, b, kEvictLast);
                }
                __syncwarp();
            }
        }
    } else if (warp == D_MMA_WARP) {
        // ======================================= MMA issuer =======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
        // the whole warp walks the schedule and waits; one elected lane issues MMAs and commits
        const uint64_t kdesc_hi = umma_desc_kmajor_sw128(0);
        const uint64_t vdesc_hi = umma_desc_mnmajor_sw128(0, BOX_BYTES);
        int cs[2] = {0, 0};             // score tiles issued per slot   (SFULL phases)
        int cp[2] = {0, 0};             // PV tiles issued per slot      (PFULL / PVDONE phases)
        int qi[2] = {0, 0};             // items started per slot        (QFULL phases)
        // the tile of slot s whose PV is still to be issued
        bool pend[2] = {false, false};
        int pend_g[2] = {0, 0}, pend_j[2] = {0, 0};
        bool pend_last_slot[2] = {false, false};

        // O_s (+)= P_s(g) . V_g : M = 128, N = 128 head columns, K = keys; A = P in TMEM (bf16 pairs), B = V MN-major
        auto issue_pv = [&](int s) {
            const int g = pend_g[s], j = pend_j[s], vs = g % VST;
            DTR(2 + s, g, 0);
            mbar_wait(bar(B_PFULL + s), cp[s] & 1, 24);
            mbar_wait(bar(B_VFULL + vs), (g / VST) & 1, 22);
            tc_fence_after();
            DTR(2 + s, g, 1);
            if (elect_one()) {
                constexpr uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
                const uint32_t va = (sbase + V_OFF + vs * TILE_BYTES) >> 4;
                const int ksteps = (j == T - 1 ? tail16 : 128) / 16;
                for (int k = 0; k < ksteps; ++k)
                    umma_bf16_ts(tmem + TM_O + 128 * s, tmem + TM_S + 128 * s + 8 * k,
                                 vdesc_hi | (uint64_t)(va + k * (2048 >> 4)), idesc, (j | k) != 0);
                umma_commit(bar(B_PVDONE + s));
                if (pend_last_slot[s]) umma_commit(bar(B_VEMPTY + vs));
                if (j == T - 1) umma_commit(bar(B_OFULL + s));
            }
            __syncwarp();
            ++cp[s];
            pend[s] = false;
            DTR(2 + s, g, 2);
        };
        // S_s(g) = Q_s . K_g^T : M = 128, N = keys, K = 128; both operands K-major in shared memory
        auto issue_s = [&](int s, int g, int j, bool last_slot) {
            const int ks = g % KST;
            if (j == 0) { mbar_wait(bar(B_QFULL + s), qi[s] & 1, 20); ++qi[s]; }
            mbar_wait(bar(B_KFULL + ks), (g / KST) & 1, 21);
            tc_fence_after();
            DTR(2 + s, g, 3);
            if (elect_one()) {
                const uint32_t idesc = umma_idesc_bf16(128, j == T - 1 ? tail16 : 128);
                const uint32_t qa = (sbase + Q_OFF + s * TILE_BYTES) >> 4;
                const uint32_t ka = (sbase + K_OFF + ks * TILE_BYTES) >> 4;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint32_t off = ((k >> 2) * BOX_BYTES + (k & 3) * 32) >> 4;
                    umma_bf16_ss<1>(tmem + TM_S + 128 * s, kdesc_hi | (uint64_t)(qa + off), kdesc_hi | (uint64_t)(ka + off),
                                    idesc, k != 0);
                }
                umma_commit(bar(B_SFULL + s));
                if (last_slot) umma_commit(bar(B_KEMPTY + ks));
                if (j == T - 1) umma_commit(bar(B_QEMPTY + s));
            }
            __syncwarp();
            ++cs[s];
            DTR(2 + s, g, 4);
        };
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0, nqt;
            item_coords(n, b, h, q0, nqt);
            for (int j = 0; j < T; ++j) {
                const int g = n * T + j;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (pend[s]) issue_pv(s);
                    if (s < nqt) {
                        issue_s(s, g, j, s == nqt - 1);
                        pend[s] = true;
                        pend_g[s] = g;
                        pend_j[s] = j;
                        pend_last_slot[s] = (s == nqt - 1);
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (pend[s]) issue_pv(s);
    } else if (warp < 8) {
        // ======================================= softmax + epilogue =======================================
        // (no out-of-line call in this region: ptxas only honours setmaxnreg.inc for call-free code)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
        const int s = warp >> 2;                        // slot
        const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t t_s = tmem + TM_S + 128 * s + lane_off;      // scores; probabilities (bf16 pairs) over their first half
        const uint32_t t_o = tmem + TM_O + 128 * s + lane_off;
        const int row = quarter * 32 + lane;
        const float sl2 = p.scale_log2;
        const bool tr = quarter == 0;
        int c = 0, it = 0;                              // tiles / items processed by this slot
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0, nqt;
            item_coords(n, b, h, q0, nqt);
            if (s >= nqt) continue;
            float m_used = -INFINITY, l_sum = 0.f;
            for (int j = 0; j < T; ++j, ++c) {
                const int g = n * T + j;
                if (tr) DTR(s, g, 0);
                mbar_wait_nocall(bar(B_SFULL + s), c & 1);
                tc_fence_after();
                if (tr) DTR(s, g, 1);
                // rescale of the accumulator row and of the row sum when the reference maximum moves (rare)
                auto rescale = [&](float m_new, bool grow) {
                    // O_s must be quiescent: PV_s of the previous tile may still be in flight
                    mbar_wait_nocall(bar(B_PVDONE + s), (c - 1) & 1);
                    tc_fence_after();
                    const float alpha = grow ? ex2_mufu((m_used - m_new) * sl2) : 1.0f;
                    if (grow) m_used = m_new;
                    l_sum *= alpha;
#pragma unroll 1
                    for (int cc = 0; cc < HD / 16; ++cc) {
                        uint32_t ov[16];
                        tmem_ld_32x32b_x16(t_o + cc * 16, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 16; ++u) ov[u] = __float_as_uint(__uint_as_float(ov[u]) * alpha);
                        tmem_st_32x32b_x16(t_o + cc * 16, ov);
                    }
                    tmem_st_wait();
                };
                if (j == T - 1 && tail < 128) {
                    // ---- the last, partial tile of the sequence (once per item): a compact two-pass loop over the 16-key
                    // chunks the MMAs cover, scores re-read from TMEM, padding keys masked, no speculation
                    const int nch = tail16 >> 4;
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int ch = 0; ch < nch; ++ch) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(t_s + 16 * ch, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (16 * ch + i < tail) mx = fmaxf(mx, __uint_as_float(v[i]));
                    }
                    const float m_new = fmaxf(m_used, mx);
                    if (j == 0) {
                        m_used = m_new;
                    } else {
                        const bool grow = (m_new - m_used) * sl2 > 8.0f;
                        if (__any_sync(0xffffffffu, grow)) rescale(m_new, grow);
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
                            const float e0 = (16 * ch + 2 * u < tail) ? ex2_mufu(__uint_as_float(v[2 * u]) * sl2 - mb) : 0.f;
                            const float e1 = (16 * ch + 2 * u + 1 < tail) ? ex2_mufu(__uint_as_float(v[2 * u + 1]) * sl2 - mb) : 0.f;
                            part += e0 + e1;
                            pw[u] = pack_bf16(e0, e1);
                        }
                        // P chunk ch (columns 8 ch ..) lies over score chunk ch / 2, which has been read
                        tmem_st_32x32b_x8(t_s + 8 * ch, pw);
                    }
                    l_sum += part;
                } else {
                    // ---- a whole tile: all 128 scores of the row in registers, one basic block of 64 key pairs
                    uint32_t sv[128];
                    tmem_ld_32x32b_x64(t_s, &sv[0]);
                    tmem_ld_32x32b_x64(t_s + 64, &sv[64]);
                    tmem_ld_wait();
                    if (tr) DTR(s, g, 2);
                    bool need_max = (j == 0);                           // no reference maximum yet
#pragma unroll 1
                    for (;;) {
                        if (need_max) {
                            // (z is an opaque zero defined inside this branch: OR-ing it into the operands keeps ptxas from
                            // hoisting the ~90 FMNMX of the maximum pass above the test, where every tile would pay for them)
                            uint32_t z;
                            asm volatile("mov.u32 %0, 0;" : "=r"(z));
                            float mxa[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) mxa[u] = fmaxf(__uint_as_float(sv[2 * u] | z), __uint_as_float(sv[2 * u + 1] | z));
#pragma unroll
                            for (int i = 16; i < 128; i += 16)
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    mxa[u] = fmaxf(mxa[u], fmaxf(__uint_as_float(sv[i + 2 * u] | z), __uint_as_float(sv[i + 2 * u + 1] | z)));
                            const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])),
                                                   fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
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
                        // P column u holds the bf16 pair of keys (2u, 2u+1): stored 16 keys at a time over the thread's own
                        // scores, which stay in its registers (a redo rewrites them)
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch) {
                            uint32_t pw[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int i0 = 16 * ch + 2 * u;
                                const float2 x = ffma2(make_float2(__uint_as_float(sv[i0]), __uint_as_float(sv[i0 + 1])), sc2, nmb2);
                                const bool poly = POLY > 0 && ((u + 1) * POLY / 8 != u * POLY / 8);
                                const float2 e = poly ? ex2_poly2(x) : make_float2(ex2_mufu(x.x), ex2_mufu(x.y));
                                rs2[u & 3] = fadd2(rs2[u & 3], e);
                                pw[u] = pack_bf16(e.x, e.y);
                            }
                            tmem_st_32x32b_x8(t_s + 8 * ch, pw);
                        }
                        const float part = (rs2[0].x + rs2[0].y) + (rs2[1].x + rs2[1].y) + (rs2[2].x + rs2[2].y) + (rs2[3].x + rs2[3].y);
                        // an exponential above 2^8 (or a saturated polynomial: its argument is clamped to 2^127) shows in the sum
                        if (need_max || !__any_sync(0xffffffffu, !(part <= 256.0f))) {
                            l_sum += part;
                            break;
                        }
                        need_max = true;
                    }
                }
                if (tr) DTR(s, g, 4);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_PFULL + s));
                if (tr) DTR(s, g, 5);
            }
            // ---- epilogue: O / l -> bf16, token-major.  The next item's scores may already be on their way; its first
            // PV waits for this slot's next P, which these threads only produce after the loads below have completed.
            mbar_wait_nocall(bar(B_OFULL + s), it & 1);
            tc_fence_after();
            if (tr) DTR(s, n * T + T - 1, 6);
            ++it;
            const float inv = 1.0f / l_sum;
            const int qrow = q0 + s * 128 + row;
            __nv_bfloat16* orow = p.out + ((int64_t)b * p.L + qrow) * p.ldo + h * HD;
#pragma unroll 1
            for (int cc = 0; cc < HD / 32; ++cc) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(t_o + cc * 32, ov);
                tmem_ld_wait();
                if (qrow < p.Lq) {
                    // 256-bit stores: every lane writes whole 32-byte sectors (a lane owns a row; 128-bit stores left half
                    // sectors to be merged in L2 and cost ~16000 clk per item)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        uint32_t w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            w[e] = pack_bf16(__uint_as_float(ov[16 * u + 2 * e]) * inv, __uint_as_float(ov[16 * u + 2 * e + 1]) * inv);
                        st_global_v8(orow + cc * 32 + 16 * u, w);
                    }
                }
            }
            tc_fence_before();
            if (tr) DTR(s, n * T + T - 1, 7);
        }
    }
    else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));      // idle warps of the third warpgroup
    }
    // teardown: everyone done with TMEM
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == D_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc<1>(tmem, 512);
    }
}

#ifdef MMADA_ATT_TRACE
long long* g_duo_trace = nullptr;
#endif

template <int POLY>
int launch_duo(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L, int Lq, int H,
               float scale, cudaStream_t stream) {
    CUtensorMap mq, mk, mv;
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
    const uint32_t box[3] = {64, 128, 1};
    int st;
    if ((st = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box))) return st;
    if ((st = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box))) return st;
    if ((st = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box))) return st;
    auto kern = attention_duo_kernel<POLY>;
    static bool configured[kMaxDevices] = {};
    MMADA_CUDA_TRY(ensure_dynamic_smem(kern, D_SMEM_BYTES, configured));
    DuoParams p = {};
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.L = L; p.H = H; p.B = B;
    p.Lq = Lq;
    p.q_pairs = (Lq + 255) / 256;
    p.items = B * H * p.q_pairs;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef MMADA_ATT_TRACE
    p.trace = g_duo_trace;
#endif
    int ctas = num_sms();
    if (ctas > p.items) ctas = p.items;
    kern<<<ctas, D_THREADS, D_SMEM_BYTES, stream>>>(mq, mk, mv, p);
    return cuda_status(cudaGetLastError());
}

}  // namespace

// head_dim 128 entry used by mmada_attention_bf16 (attention.cu): query rows [0, Lq) of every (batch, head)
int launch_attention_duo(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                         int Lq, int H, float scale, int poly, cudaStream_t stream) {
    switch (poly) {
        case 0: return launch_duo<0>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 1: return launch_duo<1>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 3: return launch_duo<3>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 4: return launch_duo<4>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        default: return launch_duo<2>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
    }
}

}  // namespace mmada

#ifdef MMADA_ATT_TRACE
extern "C" void mmada_attention_duo_set_trace(void* buf) { mmada::g_duo_trace = (long long*)buf; }
#endif
