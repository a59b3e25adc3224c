// Bidirectional flash attention, head_dim 128: attention_duo64.cu on CTA PAIRS (cta_group::2).
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False) at
// /root/reference/models/modeling_llada.py:653-660 (no mask is ever applied).
//
// attention_duo64.cu is bound by shared-memory operand reads: its score MMAs (128 x 64 x 16, both operands in shared
// memory) re-read Q for every 64 keys and run at 59 instead of 32 clk, and the K / V tiles arrive through the same
// port.  Here two CTAs share a work item of 512 query rows: a slot is one 256-row UMMA tile (128 rows per CTA), every
// CTA holds HALF of every K tile (32 of the 64 keys of a score sub-tile) and half of every V tile (64 of the 128 head
// columns) — per score MMA a CTA reads 4 + 1 KiB instead of 4 + 2, per PV MMA 2 instead of 4, and it receives half the
// K / V bytes; the 256 x 64 x 16 score MMA takes 43.6 clk for twice the rows (profiles/r01_ubench_mma_rate.txt).
// Everything else is attention_duo64.cu: two slots per CTA, two 64-column score buffers per slot (TMEM 2 x 2 x 64 + 2 x 128),
// thread per query row, speculative exponentials, one MMA-issuing warp per slot (in the leader CTA, for both CTAs),
// per-warp TMA-store epilogue, one-slot / few-row items at the end of a sequence.
// MEASURED (EXPERIMENTS builds only, MMADA_ATT_KERNEL=4; parity green): 0.658 ms against 0.599 ms for attention_duo64.cu at
// config 2.  The score MMAs do get cheaper (8 of them issue in 314 instead of 640 clk), but every PV now waits for the P of
// BOTH CTAs and every score buffer for a commit that crosses the pair: ~450 clk between the leader's own P and the last
// remote arrival, ~400 clk from the commit to the remote softmax warps — a round trip per sub-tile that eats the slack of
// the double buffer (the softmax warps wait ~400 clk per sub-tile again; period 1585 against 1540 clk), and the one-slot
// item at the end of a sequence occupies two SMs.  Kept for A/B runs.
//   slot s of an item = query rows [q0 + 256 s, +256): rank r owns rows q0 + 256 s + 128 r ..
//   barriers the leader's issuers wait on (Q / K / V full, P full) live in the leader's shared memory and are signalled
//   by both CTAs; everything the softmax / producer warps wait on is committed into BOTH CTAs (multicast commit).
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
constexpr int KST = 4, VST = 4;                     // K / V ring depths
constexpr int KV_BYTES = TILE_BYTES / 2;            // this CTA's half of a K tile (2 sub-tiles x 32 keys x 128 columns) or V
                                                    // tile (128 keys x 64 head columns)
constexpr int KBOX_BYTES = 32 * 64 * 2;             // one K box: 32 keys x 64 columns
constexpr int Q_OFF = 0;                            // 2 slots
constexpr int K_OFF = 2 * TILE_BYTES;
constexpr int V_OFF = K_OFF + KST * KV_BYTES;
constexpr int O_OFF = V_OFF + VST * KV_BYTES;      // output staging: 8 softmax warps x 32 rows x 64 columns
constexpr int STAGE_BYTES = 32 * 128;
constexpr int BAR_OFF = O_OFF + 8 * STAGE_BYTES;
constexpr int D_SMEM_BYTES = BAR_OFF + 512 + 1024;
constexpr int SUB = 64;                              // keys per score sub-tile
constexpr int SUB_BYTES = SUB * 128;                 // 64 key rows of the V tile (128-byte rows)
constexpr int KSUB_BYTES = 2 * KBOX_BYTES;           // this CTA's 32 keys of one score sub-tile: two boxes
// score buffer b of slot s at TM_S + 128 s + 64 b (its P over the first 32 columns), O_s at TM_O + 128 s
constexpr int TM_S = 0, TM_O = 256;

enum : int {
    B_QFULL = 0, B_QEMPTY = 2, B_KFULL = 4, B_KEMPTY = 4 + KST, B_VFULL = 4 + 2 * KST, B_VEMPTY = 4 + 2 * KST + VST,
    B_SFULL = 4 + 2 * KST + 2 * VST /* [slot][buffer] */, B_PFULL = B_SFULL + 4 /* [slot][buffer] */,
    B_PVDONE = B_PFULL + 4, B_OFULL = B_PVDONE + 2, B_TMEMPTR = B_OFULL + 2
};
static_assert(B_TMEMPTR * 8 + 8 <= 512, "barrier block");

struct Pair64Params {
    int L, H, B;
    int Lq;                // query rows [0, Lq) are handled here
    int q_quads, items;   // items of 512 query rows per (batch, head)
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
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// TMA store of one box from shared memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// POLY = how many of every 8 key pairs take the polynomial 2^x (FMA pipe) instead of MUFU.EX2
template <int POLY>
__global__ void __launch_bounds__(D_THREADS, 1)
attention_pair64_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                     const Pair64Params p) {
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
    const uint32_t lead0 = mapa_u32(bar(0), 0);                   // the leader CTA's barrier block (shared::cluster)
    auto lbar = [&](int idx) { return lead0 + 8 * idx; };
    const int T = (p.L + 127) / 128;                        // K / V tiles (128 keys) per item
    const int T2 = (p.L + SUB - 1) / SUB;                   // score sub-tiles (64 keys) per item
    const int tail = p.L - (T2 - 1) * SUB;                  // valid keys in the last sub-tile (1..64)
    const int tail16 = (tail + 15) & ~15;

    if (warp == D_TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        tma_prefetch_desc(&map_o);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_QFULL + i), 1);
            mbar_init(bar(B_QEMPTY + i), 1);
            for (int b = 0; b < 2; ++b) {
                mbar_init(bar(B_SFULL + 2 * i + b), 1);
                mbar_init(bar(B_PFULL + 2 * i + b), 8);      // one arrival per softmax warp of the slot, both CTAs
            }
            mbar_init(bar(B_PVDONE + i), 1);
            mbar_init(bar(B_OFULL + i), 1);
        }
        // a K / V stage is free when BOTH slots' MMA streams have passed it
        for (int s = 0; s < KST; ++s) { mbar_init(bar(B_KFULL + s), 1); mbar_init(bar(B_KEMPTY + s), 2); }
        for (int s = 0; s < VST; ++s) { mbar_init(bar(B_VFULL + s), 1); mbar_init(bar(B_VEMPTY + s), 2); }
        fence_mbar_init();
    }
    if (warp == D_MMA_WARP) {
        tmem_alloc<2>(bar(B_TMEMPTR), 512);
        tmem_relinquish<2>();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;

    // item n of this CTA -> (batch, head, first query row, number of query tiles with rows below Lq)
    // item n of this pair -> (batch, head, first query row of the item, number of slots with rows below Lq)
    auto item_coords = [&](int n, int& b, int& h, int& q0, int& nqt) {
        const int id = cluster_id + n * num_clusters;
        const int qq = id % p.q_quads, bh = id / p.q_quads;
        h = bh % p.H;
        b = bh / p.H;
        q0 = qq * 512;
        nqt = (q0 + 256 < p.Lq) ? 2 : 1;
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
                    if (leader) mbar_arrive_expect_tx(bar(B_KFULL + ks), 2 * KV_BYTES);
                    // this CTA's share of the keys of the tile's two score sub-tiles: 32 of 64, or half of the 16-key
                    // blocks the last sub-tile's MMAs cover (the box always holds 32 rows; the rest is not read)
                    for (int hh = 0; hh < 2; ++hh) {
                        const int u = 2 * j + hh;
                        const int half_n = (u == T2 - 1 ? tail16 : SUB) / 2;
                        const int key0 = u < T2 ? u * SUB + (int)rank * half_n : 0;
                        for (int c = 0; c < 2; ++c)
                            tma_load_3d_2sm(sbase + K_OFF + ks * KV_BYTES + hh * KSUB_BYTES + c * KBOX_BYTES, &map_k,
                                            lbar(B_KFULL + ks), h * HD + c * 64, key0, b, kEvictLast);
                    }
                }
                __syncwarp();
                if (j == 0) {
                    for (int s = 0; s < nqt; ++s) {
                        mbar_wait_backoff(bar(B_QEMPTY + s), (qi[s] & 1) ^ 1, 10, 100);
                        if (elect_one()) {
                            if (leader) mbar_arrive_expect_tx(bar(B_QFULL + s), 2 * TILE_BYTES);
                            const int qrow = q0 + s * 256 + (int)rank * 128;
                            for (int c = 0; c < 2; ++c)
                                tma_load_3d_2sm(sbase + Q_OFF + s * TILE_BYTES + c * BOX_BYTES, &map_q, lbar(B_QFULL + s),
                                                h * HD + c * 64, qrow < p.L ? qrow : 0, b, kEvictFirst);
                        }
                        __syncwarp();
                        ++qi[s];
                    }
                }
                mbar_wait_backoff(bar(B_VEMPTY + vs), ((g / VST) & 1) ^ 1, 12, 100);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(bar(B_VFULL + vs), 2 * KV_BYTES);
                    tma_load_3d_2sm(sbase + V_OFF + vs * KV_BYTES, &map_v, lbar(B_VFULL + vs), h * HD + (int)rank * 64,
                                    j * 128, b, kEvictLast);
                }
                __syncwarp();
            }
        }
    } else if (warp == D_MMA_WARP || warp == D_MMA_WARP + 1) {
        // ======================================= MMA issuers (one warp per slot) =======================================
        // The two slots' MMA streams are independent (own TMEM regions, K / V read-only): with one issuer per slot the
        // barrier waits and commits of one stream overlap the other's issue (a single issuer left the tensor pipe idle for
        // ~430 of every 1170 clk: gpurun_out/attn_trace_duo64.txt).  K / V stages are released by both issuers (count 2).
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
        const int s = warp - D_MMA_WARP;
        if (leader) {                                   // the leader CTA issues for the pair
        // the whole warp walks the schedule and waits; one elected lane issues MMAs and commits
        const uint64_t kdesc_hi = umma_desc_kmajor_sw128(0);
        const uint64_t vdesc_hi = umma_desc_mnmajor_sw128(0, KV_BYTES);
        const uint32_t t_s = tmem + TM_S + 128 * s, t_o = tmem + TM_O + 128 * s;
        int cs = 0;                     // score sub-tiles issued (buffer = count & 1, phase = count >> 1)
        int cp = 0;                     // PV sub-tiles issued
        int qi = 0;                     // items started (QFULL phases)
        // the (at most two) sub-tiles whose scores are issued and whose PV is still to come, oldest first:
        // global K/V tile index, flags
        enum : int { F_HALF = 1, F_FIRST = 2, F_LAST = 4, F_RELV = 8 };
        int qn = 0, e_gt0 = 0, e_gt1 = 0, e_fl0 = 0, e_fl1 = 0;

        // O_s (+)= P_s(u) . V(u) : M = 128, N = 128 head columns, K = 64 keys; A = P in TMEM (bf16 pairs), B = V MN-major
        auto issue_pv = [&]() {
            const int gt = e_gt0, fl = e_fl0;
            e_gt0 = e_gt1;
            e_fl0 = e_fl1;
            --qn;
            const int vs = gt % VST, b = cp & 1;
            DTR(2 + s, cp, 0);
            mbar_wait(bar(B_PFULL + 2 * s + b), (cp >> 1) & 1, 24);
            mbar_wait(bar(B_VFULL + vs), (gt / VST) & 1, 22);
            tc_fence_after();
            DTR(2 + s, cp, 1);
            if (elect_one()) {
                constexpr uint32_t idesc = umma_idesc_bf16(256, HD, 0, 1);
                const uint32_t va = (sbase + V_OFF + vs * KV_BYTES + ((fl & F_HALF) ? SUB_BYTES : 0)) >> 4;
                const int ksteps = ((fl & F_LAST) ? tail16 : SUB) / 16;
                for (int k = 0; k < ksteps; ++k)
                    umma_bf16_ts_cg<2>(t_o, t_s + 64 * b + 8 * k, vdesc_hi | (uint64_t)(va + k * (2048 >> 4)), idesc,
                                       !((fl & F_FIRST) && k == 0));
                umma_commit_2sm(bar(B_PVDONE + s), 0x3);
                if (fl & F_RELV) umma_commit_2sm(bar(B_VEMPTY + vs), 0x3);
                if (fl & F_LAST) umma_commit_2sm(bar(B_OFULL + s), 0x3);
            }
            __syncwarp();
            DTR(2 + s, cp, 2);
            ++cp;
        };
        // S_s(u) = Q_s . K(u)^T : M = 128, N = 64 keys, K = 128; both operands K-major in shared memory
        auto issue_s = [&](int gt, int u) {
            const int ks = gt % KST, b = cs & 1;
            const bool last = (u == T2 - 1), tile_done = last || (u & 1);
            if (u == 0) { mbar_wait(bar(B_QFULL + s), qi & 1, 20); ++qi; }
            mbar_wait(bar(B_KFULL + ks), (gt / KST) & 1, 21);
            tc_fence_after();
            DTR(2 + s, cs, 3);
            if (elect_one()) {
                const uint32_t idesc = umma_idesc_bf16(256, last ? tail16 : SUB);
                const uint32_t qa = (sbase + Q_OFF + s * TILE_BYTES) >> 4;
                const uint32_t ka = (sbase + K_OFF + ks * KV_BYTES + ((u & 1) ? KSUB_BYTES : 0)) >> 4;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint32_t qoff = ((k >> 2) * BOX_BYTES + (k & 3) * 32) >> 4;
                    const uint32_t koff = ((k >> 2) * KBOX_BYTES + (k & 3) * 32) >> 4;
                    umma_bf16_ss<2>(t_s + 64 * b, kdesc_hi | (uint64_t)(qa + qoff), kdesc_hi | (uint64_t)(ka + koff), idesc,
                                    k != 0);
                }
                umma_commit_2sm(bar(B_SFULL + 2 * s + b), 0x3);
                if (tile_done) umma_commit_2sm(bar(B_KEMPTY + ks), 0x3);
                if (last) umma_commit_2sm(bar(B_QEMPTY + s), 0x3);
            }
            __syncwarp();
            DTR(2 + s, cs, 4);
            ++cs;
            const int fl = ((u & 1) ? F_HALF : 0) | (u == 0 ? F_FIRST : 0) | (last ? F_LAST : 0) | (tile_done ? F_RELV : 0);
            if (qn == 0) { e_gt0 = gt; e_fl0 = fl; } else { e_gt1 = gt; e_fl1 = fl; }
            ++qn;
        };
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0, nqt;
            item_coords(n, b, h, q0, nqt);
            if (s < nqt) {
                for (int u = 0; u < T2; ++u) {
                    if (qn == 2) issue_pv();
                    issue_s(n * T + (u >> 1), u);
                }
                // finish the item before the next one's scores: its last PV completes O, which the softmax threads are
                // waiting for to write the rows out (left pending, it sat behind the next item's Q load: 2500 clk), while
                // the next item's first scores are not needed before that epilogue is through
                while (qn > 0) issue_pv();
            } else {
                // no query tile for this slot in this item: finish what is pending (its PVs release V stages), then pass
                // every K / V stage of the item on in step with the loads (never ahead of them: count-2 barriers)
                while (qn > 0) issue_pv();
                for (int j = 0; j < T; ++j) {
                    const int gt = n * T + j;
                    mbar_wait(bar(B_KFULL + gt % KST), (gt / KST) & 1, 25);
                    if (elect_one()) {
                        mbar_arrive(bar(B_KEMPTY + gt % KST));
                        mbar_arrive_cluster(mapa_u32(bar(B_KEMPTY + gt % KST), 1));
                    }
                    __syncwarp();
                    mbar_wait(bar(B_VFULL + gt % VST), (gt / VST) & 1, 26);
                    if (elect_one()) {
                        mbar_arrive(bar(B_VEMPTY + gt % VST));
                        mbar_arrive_cluster(mapa_u32(bar(B_VEMPTY + gt % VST), 1));
                    }
                    __syncwarp();
                }
            }
        }
        while (qn > 0) issue_pv();
        }
    } else if (warp < 8) {
        // ======================================= softmax + epilogue =======================================
        // (no out-of-line call in this region: ptxas only honours setmaxnreg.inc for call-free code)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
        const int s = warp >> 2;                        // slot
        const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t t_s0 = tmem + TM_S + 128 * s + lane_off;     // score buffer 0 of the slot (buffer 1: + 64)
        const uint32_t t_o = tmem + TM_O + 128 * s + lane_off;
        const float sl2 = p.scale_log2;
        const bool tr = quarter == 0;
        int c = 0, it = 0;                              // sub-tiles / items processed by this slot
        for (int n = 0; n < n_items; ++n) {
            int b, h, q0, nqt;
            item_coords(n, b, h, q0, nqt);
            if (s >= nqt) continue;
            if (q0 + s * 256 + (int)rank * 128 + quarter * 32 >= p.Lq) {
                // none of this warp's 32 query rows exists (the last item of a sequence: L = 1539 leaves 3 rows): keep the
                // protocol going and do no work.  Their P rows stay whatever the buffer holds — rows of an MMA are
                // independent and these rows of O are never stored.
                for (int u = 0; u < T2; ++u, ++c) {
                    mbar_wait_nocall(bar(B_SFULL + 2 * s + (c & 1)), (c >> 1) & 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lbar(B_PFULL + 2 * s + (c & 1)));
                }
                mbar_wait_nocall(bar(B_OFULL + s), it & 1);
                ++it;
                continue;
            }
            float m_used = -INFINITY, l_sum = 0.f;
            if (q0 + s * 256 + (int)rank * 128 + quarter * 32 + 8 >= p.Lq) {
                // ---- at most 8 of this warp's query rows exist (the last item of a sequence: L = 1539 leaves 3).  A thread
                // per row would spend a full warp's exponentials on them; here FOUR threads share a row (tcgen05.ld.16x256b:
                // thread t holds row t / 4, key columns 8 i + 2 (t % 4) + {0, 1}; P goes back with st.16x128b, whose
                // layout is exactly that of the packed pairs): 16 exponentials per thread and sub-tile instead of 64.  Exact
                // row maximum (two shuffles), lazy rescale as everywhere else; rows 8-31 of the warp are never stored.
                const int c4 = lane & 3;
                float lpart = 0.f;                       // this thread's share of its row's sum
                for (int u = 0; u < T2; ++u, ++c) {
                    const int buf = c & 1;
                    const uint32_t t_s = t_s0 + 64 * buf;
                    mbar_wait_nocall(bar(B_SFULL + 2 * s + buf), (c >> 1) & 1);
                    tc_fence_after();
                    uint32_t v[32];
                    tmem_ld_16x256b_x8(t_s, v);
                    tmem_ld_wait();
                    const int nk = (u == T2 - 1) ? tail : SUB;
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int col = 8 * i + 2 * c4;
                        if (col < nk) mx = fmaxf(mx, __uint_as_float(v[4 * i]));
                        if (col + 1 < nk) mx = fmaxf(mx, __uint_as_float(v[4 * i + 1]));
                    }
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    const float m_new = fmaxf(m_used, mx);
                    if (u == 0) {
                        m_used = m_new;
                    } else {
                        const bool grow = (m_new - m_used) * sl2 > 8.0f;
                        if (__any_sync(0xffffffffu, grow)) {
                            // O_s quiescent: see the thread-per-row path below
                            mbar_wait_nocall(bar(B_PVDONE + s), (c - 1) & 1);
                            tc_fence_after();
                            const float alpha = grow ? ex2_mufu((m_used - m_new) * sl2) : 1.0f;
                            if (grow) m_used = m_new;
                            lpart *= alpha;
                            float alpha_row = __shfl_sync(0xffffffffu, alpha, (4 * lane) & 31);    // lane r owns lane r of O
                            if (lane >= 8) alpha_row = 1.0f;
#pragma unroll 1
                            for (int cc = 0; cc < HD / 16; ++cc) {
                                uint32_t ov[16];
                                tmem_ld_32x32b_x16(t_o + cc * 16, ov);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha_row);
                                tmem_st_32x32b_x16(t_o + cc * 16, ov);
                            }
                            tmem_st_wait();
                        }
                    }
                    const float mb = m_used * sl2;
                    uint32_t pw[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int col = 8 * i + 2 * c4;
                        const float e0 = col < nk ? ex2_mufu(__uint_as_float(v[4 * i]) * sl2 - mb) : 0.f;
                        const float e1 = col + 1 < nk ? ex2_mufu(__uint_as_float(v[4 * i + 1]) * sl2 - mb) : 0.f;
                        lpart += e0 + e1;
                        pw[2 * i] = pack_bf16(e0, e1);
                        pw[2 * i + 1] = 0u;                  // rows 8-15
                    }
                    tmem_st_16x128b_x8(t_s, pw);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lbar(B_PFULL + 2 * s + buf));
                }
                lpart += __shfl_xor_sync(0xffffffffu, lpart, 1);
                lpart += __shfl_xor_sync(0xffffffffu, lpart, 2);
                l_sum = __shfl_sync(0xffffffffu, lpart, (4 * lane) & 31);               // back to one thread per row
                if (lane >= 8) l_sum = 1.0f;
            } else
            for (int u = 0; u < T2; ++u, ++c) {
                const int buf = c & 1;
                const uint32_t t_s = t_s0 + 64 * buf;   // scores; probabilities (bf16 pairs) over their first 32 columns
                if (tr) DTR(s, c, 0);
                mbar_wait_nocall(bar(B_SFULL + 2 * s + buf), (c >> 1) & 1);
                tc_fence_after();
                if (tr) DTR(s, c, 1);
                // rescale of the accumulator row and of the row sum when the reference maximum moves (rare)
                auto rescale = [&](float m_new, bool grow) {
                    // O_s must be quiescent.  S(c) complete means PV(c-2) complete (in-order pipe); PV(c-1) may be in
                    // flight or not yet issued, PV(c) cannot start before this thread hands P(c) over: the barrier has
                    // seen c - 1 or c completions
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
                        for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                        tmem_st_32x32b_x16(t_o + cc * 16, ov);
                    }
                    tmem_st_wait();
                };
                if (u == T2 - 1 && tail < SUB) {
                    // ---- the last, partial sub-tile of the sequence (once per item): a compact two-pass loop over the
                    // 16-key chunks the MMAs cover, scores re-read from TMEM, padding keys masked, no speculation
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
                    if (u == 0) {
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
                        for (int i = 0; i < 8; ++i) {
                            const float e0 = (16 * ch + 2 * i < tail) ? ex2_mufu(__uint_as_float(v[2 * i]) * sl2 - mb) : 0.f;
                            const float e1 = (16 * ch + 2 * i + 1 < tail) ? ex2_mufu(__uint_as_float(v[2 * i + 1]) * sl2 - mb) : 0.f;
                            part += e0 + e1;
                            pw[i] = pack_bf16(e0, e1);
                        }
                        // P chunk ch (columns 8 ch ..) lies over score chunk ch / 2, which has been read
                        tmem_st_32x32b_x8(t_s + 8 * ch, pw);
                    }
                    l_sum += part;
                } else {
                    // ---- a whole sub-tile: all 64 scores of the row in registers, one basic block of 32 key pairs
                    uint32_t sv[SUB];
                    tmem_ld_32x32b_x64(t_s, &sv[0]);
                    tmem_ld_wait();
                    if (tr) DTR(s, c, 2);
                    bool need_max = (u == 0);                           // no reference maximum yet
#pragma unroll 1
                    for (;;) {
                        if (need_max) {
                            // (z is an opaque zero defined inside this branch: OR-ing it into the operands keeps ptxas from
                            // hoisting the FMNMX of the maximum pass above the test, where every sub-tile would pay for them)
                            uint32_t z;
                            asm volatile("mov.u32 %0, 0;" : "=r"(z));
                            float mxa[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) mxa[i] = fmaxf(__uint_as_float(sv[2 * i] | z), __uint_as_float(sv[2 * i + 1] | z));
#pragma unroll
                            for (int i = 16; i < SUB; i += 16)
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    mxa[k] = fmaxf(mxa[k], fmaxf(__uint_as_float(sv[i + 2 * k] | z), __uint_as_float(sv[i + 2 * k + 1] | z)));
                            const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])),
                                                   fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
                            const float m_new = fmaxf(m_used, mx);
                            if (u == 0) {
                                m_used = m_new;
                            } else {
                                const bool grow = m_new > m_used;
                                if (__any_sync(0xffffffffu, grow)) rescale(m_new, grow);
                            }
                        }
                        const float mb = m_used * sl2;
                        const float2 sc2 = make_float2(sl2, sl2), nmb2 = make_float2(-mb, -mb);
                        float2 rs2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                        // P column i holds the bf16 pair of keys (2i, 2i+1): stored 16 keys at a time over the thread's own
                        // scores, which stay in its registers (a redo rewrites them)
#pragma unroll
                        for (int ch = 0; ch < SUB / 16; ++ch) {
                            uint32_t pw[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int i0 = 16 * ch + 2 * i;
                                const float2 x = ffma2(make_float2(__uint_as_float(sv[i0]), __uint_as_float(sv[i0 + 1])), sc2, nmb2);
                                const bool poly = POLY > 0 && ((i + 1) * POLY / 8 != i * POLY / 8);
                                const float2 e = poly ? ex2_poly2(x) : make_float2(ex2_mufu(x.x), ex2_mufu(x.y));
                                rs2[i & 3] = fadd2(rs2[i & 3], e);
                                pw[i] = pack_bf16(e.x, e.y);
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
                if (tr) DTR(s, c, 4);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(lbar(B_PFULL + 2 * s + buf));
                if (tr) DTR(s, c, 5);
            }
            // ---- epilogue: O / l -> bf16, token-major.  The next item's scores may already be on their way; its first
            // PV waits for this slot's next P, which these threads only produce after the loads below have completed.
            mbar_wait_nocall(bar(B_OFULL + s), it & 1);
            tc_fence_after();
            if (tr) DTR(s, c - 1, 6);
            ++it;
            const float inv = 1.0f / l_sum;
            // all four loads in flight before the first wait (the score registers are dead here)
            uint32_t ovv[HD];
#pragma unroll
            for (int cc = 0; cc < HD / 32; ++cc) tmem_ld_32x32b_x32(t_o + cc * 32, &ovv[cc * 32]);
            tmem_ld_wait();
            // Rows leave through shared memory and the TMA: a lane owns a row, so direct stores touch 32 different lines
            // per instruction (one L1 tag cycle each: ~2000 clk per item for the eight warps).  Every warp stages its own
            // 32 rows x 64 columns (4 KiB, 128-byte swizzle) and stores them as one box — no barrier across warps; the
            // second half reuses the buffer once the first store has read it.  Rows >= Lq are clipped by the tensor map.
            const uint32_t stage = sbase + O_OFF + warp * STAGE_BYTES;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (lane == 0) tma_store_wait_read();       // the buffer's previous store (other half / previous item)
                __syncwarp();
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {            // 16-byte chunk ch of the 128-byte row: columns 8 ch ..
                    const uint32_t* ov = &ovv[64 * half + 8 * ch];
                    const uint32_t w0 = pack_bf16(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
                    const uint32_t w1 = pack_bf16(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
                    const uint32_t w2 = pack_bf16(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
                    const uint32_t w3 = pack_bf16(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
                    st_shared_v4(stage + lane * 128 + ((ch ^ (lane & 7)) << 4), w0, w1, w2, w3);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&map_o, stage, h * HD + 64 * half, q0 + s * 256 + (int)rank * 128 + quarter * 32, b);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            if (tr) DTR(s, c - 1, 7);
        }
        if (lane == 0) tma_store_wait_all();
    }
    else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));      // idle warp of the third warpgroup
    }
    // teardown: everyone done with TMEM
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();                 // the peer is done with our shared memory / barriers as well
    if (warp == D_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc<2>(tmem, 512);
    }
}

#ifdef MMADA_ATT_TRACE
long long* g_pair64_trace = nullptr;
#endif

template <int POLY>
int launch_pair64(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L, int Lq, int H,
               float scale, cudaStream_t stream) {
    CUtensorMap mq, mk, mv, mo;
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
    const uint32_t box[3] = {64, 128, 1}, kbox[3] = {64, 32, 1};
    int st;
    if ((st = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box))) return st;
    if ((st = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, kbox))) return st;
    if ((st = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box))) return st;
    {
        // output rows [0, Lq) of every sequence, boxes of 32 rows x 64 columns (one softmax warp's rows)
        const uint64_t odims[3] = {(uint64_t)H * HD, (uint64_t)Lq, (uint64_t)B};
        const uint64_t ostrides[2] = {(uint64_t)ldo * 2, (uint64_t)L * ldo * 2};
        const uint32_t obox[3] = {64, 32, 1};
        if ((st = make_tmap(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, odims, ostrides, obox))) return st;
    }
    auto kern = attention_pair64_kernel<POLY>;
    static bool configured[kMaxDevices] = {};
    MMADA_CUDA_TRY(ensure_dynamic_smem(kern, D_SMEM_BYTES, configured));
    Pair64Params p = {};
    p.L = L; p.H = H; p.B = B;
    p.Lq = Lq;
    p.q_quads = (Lq + 511) / 512;
    p.items = B * H * p.q_quads;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef MMADA_ATT_TRACE
    p.trace = g_pair64_trace;
#endif
    int clusters = num_sms() / 2;
    if (clusters > p.items) clusters = p.items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(D_THREADS);
    cfg.dynamicSmemBytes = D_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MMADA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, mo, p));
    return kOk;
}

}  // namespace

// head_dim 128 entry used by mmada_attention_bf16 (attention.cu): query rows [0, Lq) of every (batch, head).
// poly = share of the exponentials on the FMA pipe in eighths; the product runs them all on the MUFU pipe (measured:
// 0.626 / 0.640 / 0.674 ms for 0 / 1 / 2 eighths at config 2 — with two softmax warps per scheduler in flight the FMA
// pipe's issue slots are worth more than the MUFU cycles saved); other shares exist in EXPERIMENTS builds only.
int launch_attention_pair64(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int B, int L,
                           int Lq, int H, float scale, int poly, cudaStream_t stream) {
    switch (poly) {
#ifdef MMADA_EXPERIMENTS
        case 1: return launch_pair64<1>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 2: return launch_pair64<2>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 3: return launch_pair64<3>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
        case 4: return launch_pair64<4>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
#endif
        default: return launch_pair64<0>(q, k, v, ld, out, ldo, B, L, Lq, H, scale, stream);
    }
}

}  // namespace mmada

#ifdef MMADA_ATT_TRACE
extern "C" void mmada_attention_pair64_set_trace(void* buf) { mmada::g_pair64_trace = (long long*)buf; }
#endif
