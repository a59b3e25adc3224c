// tcgen05 / TMEM / TMA GEMM for the LLaDA block projections, the restricted lm_head and (as an
// implicit GEMM) the 3x3 / 1x1 convolutions of the MAGVIT-v2 decoder.
//
//   D[M,N] = A[M,K] . B[N,K]^T      A, B bf16 (K contiguous), fp32 accumulation in TMEM
//
// replaces the nn.Linear calls of the reference's LLaDALlamaBlock
// (/root/reference/models/modeling_llada.py:901-903 q/k/v_proj, :724 attn_out, :924 ff_proj/up_proj,
//  :930 ff_out, :1362 transformer.ff_out), each a cuBLAS GEMM there, and the cuDNN convolutions of
// VQGANDecoder (/root/reference/models/modeling_magvitv2.py:365-399, models/common_modules.py).
//
// Structure (one persistent CTA, or CTA pair, per SM):
//   warp 0      TMA producer: A and B k-blocks (64 columns = one 128-byte swizzle row) into a ring of
//               shared-memory stages, signalled through mbarriers.  In convolution mode the A tile is
//               a 4-D box (64 channels x BW x BH pixels of one image, NHWC) fetched at the tap's
//               offset; out-of-image coordinates are zero-filled by the TMA unit = zero padding.
//   warp 1      allocates TMEM; one elected lane issues tcgen05.mma (UMMA 128xBNx16, or 256xBNx16 with
//               cta_group::2) into one of two BN-column accumulator stages; tcgen05.commit releases
//               smem stages and publishes finished accumulators
//   warps 2..5  epilogue: tcgen05.ld the accumulator (one warp per 32-lane quarter), apply the fused
//               epilogue, store to global; overlaps with the next tile's MMAs
// The producer and MMA warps run their loops with all 32 lanes (addresses stay warp-uniform) and
// elect one lane only around the TMA / MMA instructions: a single diverged lane executing the loop
// bookkeeping was measured to cap the tensor pipe at ~36 %.
#include <stdlib.h>

#include "common.cuh"
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

constexpr int BK = 64;          // k-block: 64 bf16 = 128 bytes = one swizzle row
constexpr int BM = 128;         // A rows per CTA
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct GemmParams {
    void* out;
    const void* aux;     // residual (fp32, ld = ldo)
    const float* bias;   // fp32 [N]
    int64_t ldo;
    int M, N, K;
    int num_m_tiles, num_n_tiles;
    // convolution mode (A = NHWC activations [B,H,W,C], K = taps * C)
    int conv_H, conv_W, conv_C, conv_taps;   // taps = 9 (3x3, pad 1) or 1
    const float* rope_sin;                   // MMADA_EPI_ROPE_BF16: fp32 tables [>= rope_L, rope_hd/2]
    const float* rope_cos;
    int rope_hd, rope_L, rope_cols;          // head dim, sequence length (position = row % L), columns [0, rope_cols) are rotated
    int group_m;                             // rasterisation: tiles walk group_m m-tiles before the next n-tile
    int serpentine;                          // odd groups sweep the n-tiles backwards: the B bands the previous group read
                                             // last are still in L2 when the next group starts
    int group_by_n;                          // 1: the group is group_m N-tiles (a band of B stays in L2) and the m-tiles sweep
    // RMSNorm folded into the GEMMs around it (models/modeling_llada.py:315-329).  Producer (EPI_RESID_NORM_F32):
    // besides x (fp32) the epilogue writes bf16(x) = the next GEMM's A operand and, per n-tile, the row's sum of
    // squares of the 256 new values.  Consumer (row_ssq != nullptr): accumulator rows are scaled by
    // rsqrt(sum(ssq[row, 0..ssq_tiles)) * ssq_inv_d + ssq_eps); the norm weight is folded into B's columns.
    __nv_bfloat16* xb_out;
    int64_t ld_xb;
    float* ssq_out;                          // [M, num_n_tiles]
    const float* row_ssq;                    // [M, ssq_tiles]
    int ssq_tiles;
    float ssq_inv_d, ssq_eps;
    uint64_t hint_a, hint_b;                 // L2 eviction-priority hints of the operand loads
};

template <int CG, int BN>
struct GemmCfg {
    static constexpr int LOAD_N = BN / CG;                     // B rows loaded per CTA
    static constexpr int A_BYTES = BM * BK * 2;                // 16 KiB
    static constexpr int B_BYTES = LOAD_N * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (196 * 1024 / STAGE_BYTES) > 8 ? 8 : (196 * 1024 / STAGE_BYTES);
    static constexpr int STAGING_OFF = STAGES * STAGE_BYTES + 256;       // after the barriers
    static constexpr int STAGING_BYTES = 4 * 32 * 144;                   // 4 epilogue warps x 32 rows x (128 + 16) B
    static constexpr int SMEM_BYTES = STAGING_OFF + STAGING_BYTES + 1024 /*align*/;
};

__device__ __forceinline__ void tile_coords(int idx, int num_m_tiles, int num_n_tiles, int GROUP_M, int serpentine,
                                            int by_n, int& mt, int& nt) {
    if (by_n) {
        const int per_group = GROUP_M * num_m_tiles;
        const int g = idx / per_group;
        const int first_n = g * GROUP_M;
        const int gsize = min(GROUP_M, num_n_tiles - first_n);
        const int r = idx - g * per_group;
        nt = first_n + r % gsize;
        mt = r / gsize;
        if (serpentine && (g & 1)) mt = num_m_tiles - 1 - mt;
        return;
    }
    const int per_group = GROUP_M * num_n_tiles;
    const int g = idx / per_group;
    const int first_m = g * GROUP_M;
    const int gsize = min(GROUP_M, num_m_tiles - first_m);
    const int r = idx - g * per_group;
    mt = first_m + r % gsize;
    nt = r / gsize;
    if (serpentine && (g & 1)) nt = num_n_tiles - 1 - nt;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__host__ __device__ constexpr bool epi_has_bias(int e) { return e == MMADA_EPI_BIAS_BF16 || e == MMADA_EPI_BIAS_F32 || e == MMADA_EPI_BIAS_RESID_F32; }
constexpr int EPI_RESID_NORM_F32 = 8;     // internal: MMADA_EPI_RESID_F32 + bf16 copy + per-tile row sums of squares
__host__ __device__ constexpr bool epi_has_resid(int e) { return e == MMADA_EPI_RESID_F32 || e == MMADA_EPI_BIAS_RESID_F32 || e == EPI_RESID_NORM_F32; }
__host__ __device__ constexpr bool epi_out_bf16(int e) { return e == MMADA_EPI_BF16 || e == MMADA_EPI_BIAS_BF16 || e == MMADA_EPI_SWIGLU_BF16 || e == MMADA_EPI_ROPE_BF16; }

// one 32-column chunk of one accumulator row
template <int EPI>
__device__ __forceinline__ void store_chunk(const GemmParams& p, int row, int col0, const uint32_t (&v)[32]) {
    if (row >= p.M) return;
    const int ncols = min(32, p.N - col0);
    if (ncols <= 0) return;
    float f[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        f[t] = __uint_as_float(v[t]);
        if constexpr (epi_has_bias(EPI)) f[t] += (t < ncols) ? __ldg(p.bias + col0 + t) : 0.f;
    }
    if constexpr (epi_out_bf16(EPI)) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)row * p.ldo + col0;
        if (ncols == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(o + 8 * j) =
                    make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                               pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
        } else {
#pragma unroll
            for (int t = 0; t < 32; ++t)
                if (t < ncols) o[t] = __float2bfloat16_rn(f[t]);
        }
    } else {  // fp32 store, optional residual
        float* o = reinterpret_cast<float*>(p.out) + (int64_t)row * p.ldo + col0;
        const float* r = reinterpret_cast<const float*>(p.aux) + (int64_t)row * p.ldo + col0;
        if (ncols == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 a = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                if constexpr (epi_has_resid(EPI)) {
                    const float4 b = *reinterpret_cast<const float4*>(r + 4 * j);
                    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                }
                *reinterpret_cast<float4*>(o + 4 * j) = a;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 32; ++t)
                if (t < ncols) {
                    float x = f[t];
                    if constexpr (epi_has_resid(EPI)) x += r[t];
                    o[t] = x;
                }
        }
    }
}

// Transposes a 32-row x ROW_BYTES block (one row per lane, `w` = the lane's row as 32-bit words) through
// the warp's private staging buffer and writes it with coalesced 16-byte stores: one warp instruction
// covers 32/PIECES whole rows instead of 16 bytes of 32 different rows.  RESID adds an fp32 residual.
template <int ROW_BYTES, bool RESID, bool NORM = false>
__device__ __forceinline__ void staged_store(uint8_t* stage, const uint32_t* w, uint8_t* gptr, const float4* resid,
                                             int64_t pitch_bytes, int rows_valid, int lane, uint8_t* xb_ptr = nullptr,
                                             int64_t xb_pitch_bytes = 0, float* ssq_acc = nullptr) {
    constexpr int PITCH = ROW_BYTES + 16;        // +16: conflict-free 128-bit accesses in both directions
    constexpr int PIECES = ROW_BYTES / 16;
    constexpr int RPI = 32 / PIECES;
#pragma unroll
    for (int j = 0; j < PIECES; ++j)
        *reinterpret_cast<uint4*>(stage + lane * PITCH + j * 16) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    __syncwarp();
    const int piece = lane % PIECES, rsub = lane / PIECES;
#pragma unroll
    for (int i = 0; i < 32 / RPI; ++i) {
        const int r = i * RPI + rsub;
        uint4 v = *reinterpret_cast<const uint4*>(stage + r * PITCH + piece * 16);
        if (r < rows_valid) {
            if constexpr (RESID) {
                const float4 b = resid[i];          // loaded one chunk ahead by the caller (same lane -> (row, piece) map)
                v.x = __float_as_uint(__uint_as_float(v.x) + b.x);
                v.y = __float_as_uint(__uint_as_float(v.y) + b.y);
                v.z = __float_as_uint(__uint_as_float(v.z) + b.z);
                v.w = __float_as_uint(__uint_as_float(v.w) + b.w);
            }
            // (streaming / evict-first stores here were measured and change nothing: 288.5 against 287.3 ms per step)
            *reinterpret_cast<uint4*>(gptr + (int64_t)r * pitch_bytes + piece * 16) = v;
            if constexpr (NORM) {
                const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z), d = __uint_as_float(v.w);
                *reinterpret_cast<uint2*>(xb_ptr + (int64_t)r * xb_pitch_bytes + piece * 8) = make_uint2(pack_bf16(a, b), pack_bf16(c, d));
                ssq_acc[i] += (a * a + b * b) + (c * c + d * d);
            }
        }
    }
    __syncwarp();
}

// NeoX half-split rotary embedding of one head segment held by a thread: lo/hi = the two halves (HALF fp32
// accumulators each) of one head of one token row.  Mirrors the reference's order of operations
// (models/modeling_llada.py:402-428): the projection output is first rounded to the model dtype (bf16), the
// rotation is evaluated in fp32 as t*cos + rotate_half(t)*sin with separately rounded products, and the result is
// rounded to bf16 again.  In place: afterwards lo[0..HALF/2) / hi[0..HALF/2) hold the packed bf16 pairs of the halves.
template <int HALF>
__device__ __forceinline__ void rope_pack(uint32_t* lo, uint32_t* hi, const float* sn, const float* cs, bool rotate) {
#pragma unroll
    for (int j = 0; j < HALF; j += 4) {
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), c4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (rotate) {
            s4 = __ldg(reinterpret_cast<const float4*>(sn + j));
            c4 = __ldg(reinterpret_cast<const float4*>(cs + j));
        }
        const float s[4] = {s4.x, s4.y, s4.z, s4.w}, c[4] = {c4.x, c4.y, c4.z, c4.w};
        float ol[4], oh[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float l = __bfloat162float(__float2bfloat16_rn(__uint_as_float(lo[j + t])));
            const float h = __bfloat162float(__float2bfloat16_rn(__uint_as_float(hi[j + t])));
            ol[t] = rotate ? __fadd_rn(__fmul_rn(l, c[t]), __fmul_rn(-h, s[t])) : l;
            oh[t] = rotate ? __fadd_rn(__fmul_rn(h, c[t]), __fmul_rn(l, s[t])) : h;
        }
        lo[j / 2] = pack_bf16(ol[0], ol[1]);          // slots j/2, j/2+1 <= j: already consumed
        lo[j / 2 + 1] = pack_bf16(ol[2], ol[3]);
        hi[j / 2] = pack_bf16(oh[0], oh[1]);
        hi[j / 2 + 1] = pack_bf16(oh[2], oh[3]);
    }
}

template <int CG, int EPI, int BN, bool CONV>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p) {
    using Cfg = GemmCfg<CG, BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
    // barrier layout (8 bytes each): full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem ptr
    auto full_bar = [&](int s) { return bar_base + 8 * s; };
    auto empty_bar = [&](int s) { return bar_base + 8 * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8 * (2 * Cfg::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8 * (2 * Cfg::STAGES + 2 + a); };
    const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * Cfg::STAGES + 4);
    volatile uint32_t* tmem_ptr_smem =
        reinterpret_cast<volatile uint32_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
    const bool leader = cta_rank == 0;
    const int num_clusters = gridDim.x / CG;
    const int cluster_id = blockIdx.x / CG;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(full_bar(s), 1);       // the leader's arrive.expect_tx (+ transaction bytes of both CTAs)
            mbar_init(empty_bar(s), 1);      // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);          // one tcgen05.commit
            mbar_init(tempty_bar(a), 4 * CG);    // one arrival per epilogue warp of every CTA
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc<CG>(tmem_ptr_addr, 512);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ================================ TMA producer ================================
        const uint32_t full0 = CG == 2 ? mapa_u32(full_bar(0), 0) : full_bar(0);   // leader's barriers
        int stage = 0;
        uint32_t phase = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters) {
            int mt, nt;
            tile_coords(t, p.num_m_tiles, p.num_n_tiles, p.group_m, p.serpentine, p.group_by_n, mt, nt);
            const int m0 = mt * (BM * CG) + (int)cta_rank * BM;
            const int n0 = nt * BN + (int)cta_rank * Cfg::LOAD_N;
            // convolution: the 128 rows are 128 consecutive NHWC pixels = a BW x BH box of one image
            int cb = 0, cy = 0, cx = 0, kpc = 1;
            if constexpr (CONV) {
                const int hw = p.conv_H * p.conv_W;
                cb = m0 / hw;
                const int rem = m0 - cb * hw;
                cy = rem / p.conv_W;
                cx = rem - cy * p.conv_W;
                kpc = p.conv_C / BK;          // k-blocks per tap
            }
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1, 1);
                const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                const uint32_t sb = sa + Cfg::A_BYTES;
                if (elect_one()) {
                    const uint32_t fb = CG == 2 ? full0 + 8 * stage : full_bar(stage);
                    if (CG == 1 || leader) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * CG);
                    // (2-CTA: both CTAs' bytes are accounted on the leader's barrier; the peer's complete_tx may
                    //  precede the leader's expect_tx within the phase — transiently negative tx-count)
                    if constexpr (CONV) {
                        const int tap = kb / kpc, kc = kb - tap * kpc;
                        const int dy = p.conv_taps == 9 ? tap / 3 - 1 : 0, dx = p.conv_taps == 9 ? tap % 3 - 1 : 0;
                        if constexpr (CG == 1) tma_load_4d(sa, &map_a, fb, kc * BK, cx + dx, cy + dy, cb, p.hint_a);
                        else tma_load_4d_2sm(sa, &map_a, fb, kc * BK, cx + dx, cy + dy, cb, p.hint_a);
                    } else {
                        if constexpr (CG == 1) tma_load_2d(sa, &map_a, fb, kb * BK, m0, p.hint_a);
                        else tma_load_2d_2sm(sa, &map_a, fb, kb * BK, m0, p.hint_a);
                    }
                    if constexpr (CG == 1) tma_load_2d(sb, &map_b, fb, kb * BK, n0, p.hint_b);
                    else tma_load_2d_2sm(sb, &map_b, fb, kb * BK, n0, p.hint_b);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN);
            const uint64_t desc_hi = umma_desc_kmajor_sw128(0);          // everything but the start address
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint64_t adesc = desc_hi | (uint64_t)(sa >> 4);
                    const uint64_t bdesc = desc_hi | (uint64_t)((sa + Cfg::A_BYTES) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advance 16 bf16 = 32 bytes along K inside the swizzle row (encoded >> 4)
                            umma_bf16_ss<CG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        if constexpr (CG == 1) umma_commit(empty_bar(stage));
                        else umma_commit_2sm(empty_bar(stage), 0x3);
                        if (kb == num_kb - 1) {
                            if constexpr (CG == 1) umma_commit(tfull_bar(acc));
                            else umma_commit_2sm(tfull_bar(acc), 0x3);
                        }
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue ================================
        const int quarter = warp & 3;                    // TMEM lanes [32*quarter, +32)
        const uint32_t tempty0 = CG == 2 ? mapa_u32(tempty_bar(0), 0) : tempty_bar(0);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters) {
            int mt, nt;
            tile_coords(t, p.num_m_tiles, p.num_n_tiles, p.group_m, p.serpentine, p.group_by_n, mt, nt);
            const int row = mt * (BM * CG) + (int)cta_rank * BM + quarter * 32 + lane;
            uint8_t* stage = smem + Cfg::STAGING_OFF + (warp - 2) * (32 * 144);
            const int row0 = mt * (BM * CG) + (int)cta_rank * BM + quarter * 32;       // first row of this warp
            const int rows_valid = min(32, p.M - row0);
            constexpr int OUT_ES = epi_out_bf16(EPI) ? 2 : 4;
            const int64_t pitch = p.ldo * OUT_ES;
            const bool vec_ok = (pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
                                (!epi_has_resid(EPI) || (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0);
            // residual: each lane reads the (row, 16-byte piece) it will write after the smem transpose, one
            // 32-column chunk ahead of the accumulator — chunk 0 before the accumulator even exists, so the
            // DRAM/L2 latency of the read-modify-write never sits on the epilogue's critical path
            float4 rcur[8], rnext[8];
            const uint8_t* rbase = reinterpret_cast<const uint8_t*>(p.aux) + (int64_t)row0 * pitch + (int64_t)nt * BN * 4;
            auto load_resid = [&](int c, float4 (&dst)[8]) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + (lane >> 3);
                    dst[i] = (r < rows_valid) ? *reinterpret_cast<const float4*>(rbase + (int64_t)r * pitch + c * 128 + (lane & 7) * 16)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            const bool resid_vec = epi_has_resid(EPI) && vec_ok && nt * BN + BN <= p.N;
            if constexpr (epi_has_resid(EPI)) {
                if (resid_vec) load_resid(0, rcur);
            }
            mbar_wait(tfull_bar(acc), acc_phase, 4);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            // folded RMSNorm: this lane's accumulator row is scaled by the row's rstd (1 when not in use)
            [[maybe_unused]] float rs = 1.0f;
            if constexpr (EPI == MMADA_EPI_SWIGLU_BF16 || EPI == MMADA_EPI_ROPE_BF16) {
                if (p.row_ssq != nullptr && row < p.M) {
                    const float* sp = p.row_ssq + (int64_t)row * p.ssq_tiles;
                    float ss = 0.f;
                    for (int i = 0; i < p.ssq_tiles; ++i) ss += sp[i];
                    rs = rsqrtf(ss * p.ssq_inv_d + p.ssq_eps);
                }
            }
            if constexpr (EPI == MMADA_EPI_SWIGLU_BF16) {
                // columns [0,BN/2) = gate, [BN/2,BN) = up for output columns nt*BN/2 + [0,BN/2)
                uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + (int64_t)row0 * pitch + (int64_t)nt * (BN / 2) * 2;
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    uint32_t g[32], u[32];
                    tmem_ld_32x32b_x32(t_addr + c * 32, g);
                    tmem_ld_32x32b_x32(t_addr + BN / 2 + c * 32, u);
                    tmem_ld_wait();
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        w[j] = pack_bf16(silu_f(__uint_as_float(g[2 * j]) * rs) * (__uint_as_float(u[2 * j]) * rs),
                                         silu_f(__uint_as_float(g[2 * j + 1]) * rs) * (__uint_as_float(u[2 * j + 1]) * rs));
                    if (rows_valid > 0) staged_store<64, false>(stage, w, obase + c * 64, nullptr, pitch, rows_valid, lane);
                }
            } else if constexpr (EPI == MMADA_EPI_ROPE_BF16) {
                // fused q|k|v projection: heads inside columns [0, rope_cols) are rotated before the store
                const int pos = (row < p.M ? row : 0) % p.rope_L;
                uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + (int64_t)row0 * pitch;
                if (p.rope_hd == 128) {
                    const float* sn = p.rope_sin + (int64_t)pos * 64;
                    const float* cs = p.rope_cos + (int64_t)pos * 64;
#pragma unroll 1
                    for (int hh = 0; hh < BN / 128; ++hh) {
                        const int col0 = nt * BN + hh * 128;
                        uint32_t lo[64], hi[64];
                        tmem_ld_32x32b_x32(t_addr + hh * 128, lo);
                        tmem_ld_32x32b_x32(t_addr + hh * 128 + 32, lo + 32);
                        tmem_ld_32x32b_x32(t_addr + hh * 128 + 64, hi);
                        tmem_ld_32x32b_x32(t_addr + hh * 128 + 96, hi + 32);
                        tmem_ld_wait();
                        if (p.row_ssq != nullptr) {
#pragma unroll
                            for (int j = 0; j < 64; ++j) {
                                lo[j] = __float_as_uint(__uint_as_float(lo[j]) * rs);
                                hi[j] = __float_as_uint(__uint_as_float(hi[j]) * rs);
                            }
                        }
                        rope_pack<64>(lo, hi, sn, cs, col0 < p.rope_cols);
                        if (rows_valid > 0) {
                            staged_store<128, false>(stage, lo, obase + (int64_t)col0 * 2, nullptr, pitch, rows_valid, lane);
                            staged_store<128, false>(stage, hi, obase + (int64_t)col0 * 2 + 128, nullptr, pitch, rows_valid, lane);
                        }
                    }
                } else {   // head_dim 64
                    const float* sn = p.rope_sin + (int64_t)pos * 32;
                    const float* cs = p.rope_cos + (int64_t)pos * 32;
#pragma unroll 1
                    for (int hh = 0; hh < BN / 64; ++hh) {
                        const int col0 = nt * BN + hh * 64;
                        uint32_t lo[32], hi[32];
                        tmem_ld_32x32b_x32(t_addr + hh * 64, lo);
                        tmem_ld_32x32b_x32(t_addr + hh * 64 + 32, hi);
                        tmem_ld_wait();
                        if (p.row_ssq != nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                lo[j] = __float_as_uint(__uint_as_float(lo[j]) * rs);
                                hi[j] = __float_as_uint(__uint_as_float(hi[j]) * rs);
                            }
                        }
                        rope_pack<32>(lo, hi, sn, cs, col0 < p.rope_cols);
                        if (rows_valid > 0) {   // 64 bytes (low half) then 64 bytes (high half) of the 128-byte head
                            staged_store<64, false>(stage, lo, obase + (int64_t)col0 * 2, nullptr, pitch, rows_valid, lane);
                            staged_store<64, false>(stage, hi, obase + (int64_t)col0 * 2 + 64, nullptr, pitch, rows_valid, lane);
                        }
                    }
                }
            } else if constexpr (epi_out_bf16(EPI)) {
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    const int col0 = nt * BN + c * 64;
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32b_x32(t_addr + c * 64, v0);
                    tmem_ld_32x32b_x32(t_addr + c * 64 + 32, v1);
                    tmem_ld_wait();
                    if (vec_ok && col0 + 64 <= p.N) {
                        uint32_t w[32];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float a0 = __uint_as_float(v0[2 * j]), a1 = __uint_as_float(v0[2 * j + 1]);
                            float b0 = __uint_as_float(v1[2 * j]), b1 = __uint_as_float(v1[2 * j + 1]);
                            if constexpr (epi_has_bias(EPI)) {
                                a0 += __ldg(p.bias + col0 + 2 * j); a1 += __ldg(p.bias + col0 + 2 * j + 1);
                                b0 += __ldg(p.bias + col0 + 32 + 2 * j); b1 += __ldg(p.bias + col0 + 33 + 2 * j);
                            }
                            w[j] = pack_bf16(a0, a1);
                            w[16 + j] = pack_bf16(b0, b1);
                        }
                        if (rows_valid > 0)
                            staged_store<128, false>(stage, w, reinterpret_cast<uint8_t*>(p.out) + (int64_t)row0 * pitch + (int64_t)col0 * 2,
                                                     nullptr, pitch, rows_valid, lane);
                    } else {
                        store_chunk<EPI>(p, row, col0, v0);
                        store_chunk<EPI>(p, row, col0 + 32, v1);
                    }
                }
            } else {
                [[maybe_unused]] float ssq_acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int col0 = nt * BN + c * 32;
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_addr + c * 32, v);
                    if constexpr (epi_has_resid(EPI)) {
                        if (resid_vec && c + 1 < BN / 32) load_resid(c + 1, rnext);
                    }
                    tmem_ld_wait();
                    if (vec_ok && col0 + 32 <= p.N && (!epi_has_resid(EPI) || resid_vec)) {
                        if constexpr (epi_has_bias(EPI)) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + col0 + j));
                        }
                        if constexpr (EPI == EPI_RESID_NORM_F32) {
                            if (rows_valid > 0)
                                staged_store<128, true, true>(
                                    stage, v, reinterpret_cast<uint8_t*>(p.out) + (int64_t)row0 * pitch + (int64_t)col0 * 4, rcur,
                                    pitch, rows_valid, lane,
                                    reinterpret_cast<uint8_t*>(p.xb_out) + ((int64_t)row0 * p.ld_xb + col0) * 2, p.ld_xb * 2, ssq_acc);
                        } else {
                            if (rows_valid > 0)
                                staged_store<128, epi_has_resid(EPI)>(
                                    stage, v, reinterpret_cast<uint8_t*>(p.out) + (int64_t)row0 * pitch + (int64_t)col0 * 4, rcur, pitch,
                                    rows_valid, lane);
                        }
                    } else {
                        store_chunk<EPI>(p, row, col0, v);
                    }
                    if constexpr (epi_has_resid(EPI)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) rcur[i] = rnext[i];
                    }
                }
                if constexpr (EPI == EPI_RESID_NORM_F32) {
                    // lane holds the partial sums of rows 4 i + (lane >> 3) over its 16-byte pieces: add up the 8
                    // lanes of a row (fixed order -> deterministic) and publish the tile's share of the row
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float sacc = ssq_acc[i];
                        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
                        sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
                        const int r = i * 4 + (lane >> 3);
                        if ((lane & 7) == 0 && r < rows_valid) p.ssq_out[(int64_t)(row0 + r) * p.num_n_tiles + nt] = sacc;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(tempty_bar(acc));
                else mbar_arrive_cluster(tempty0 + 8 * acc);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    // teardown: everyone done with TMEM (and, for pairs, the peer done with our smem / barriers)
    __syncwarp();
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, 512);
    }
}

template <int CG, int EPI, int BN, bool CONV>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<CG, BN>;
    auto kern = gemm_kernel<CG, EPI, BN, CONV>;
    static bool configured[kMaxDevices] = {};
    MMADA_CUDA_TRY(ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, configured));
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    int clusters = num_sms() / CG;
    if (clusters > num_tiles) clusters = num_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CG);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MMADA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ma, mb, p));
    return kOk;
}

// plain GEMM: every epilogue at BN = 256 for both CTA-group sizes
template <int CG>
static int dispatch_gemm256(int epi, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t s) {
    switch (epi) {
        case MMADA_EPI_BF16: return launch_gemm<CG, MMADA_EPI_BF16, 256, false>(ma, mb, p, s);
        case MMADA_EPI_F32: return launch_gemm<CG, MMADA_EPI_F32, 256, false>(ma, mb, p, s);
        case MMADA_EPI_RESID_F32: return launch_gemm<CG, MMADA_EPI_RESID_F32, 256, false>(ma, mb, p, s);
        case MMADA_EPI_SWIGLU_BF16: return launch_gemm<CG, MMADA_EPI_SWIGLU_BF16, 256, false>(ma, mb, p, s);
        case MMADA_EPI_BIAS_BF16: return launch_gemm<CG, MMADA_EPI_BIAS_BF16, 256, false>(ma, mb, p, s);
        case MMADA_EPI_BIAS_F32: return launch_gemm<CG, MMADA_EPI_BIAS_F32, 256, false>(ma, mb, p, s);
        case MMADA_EPI_BIAS_RESID_F32: return launch_gemm<CG, MMADA_EPI_BIAS_RESID_F32, 256, false>(ma, mb, p, s);
        case MMADA_EPI_ROPE_BF16: return launch_gemm<CG, MMADA_EPI_ROPE_BF16, 256, false>(ma, mb, p, s);
        case EPI_RESID_NORM_F32: return launch_gemm<CG, EPI_RESID_NORM_F32, 256, false>(ma, mb, p, s);
    }
    return kBadArgument;
}
// narrow outputs (N <= 128): BN = 128, pairs only
template <bool CONV>
static int dispatch_narrow(int epi, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t s) {
    switch (epi) {
        case MMADA_EPI_BF16: return launch_gemm<2, MMADA_EPI_BF16, 128, CONV>(ma, mb, p, s);
        case MMADA_EPI_F32: return launch_gemm<2, MMADA_EPI_F32, 128, CONV>(ma, mb, p, s);
        case MMADA_EPI_BIAS_BF16: return launch_gemm<2, MMADA_EPI_BIAS_BF16, 128, CONV>(ma, mb, p, s);
        case MMADA_EPI_BIAS_F32: return launch_gemm<2, MMADA_EPI_BIAS_F32, 128, CONV>(ma, mb, p, s);
        case MMADA_EPI_BIAS_RESID_F32: return launch_gemm<2, MMADA_EPI_BIAS_RESID_F32, 128, CONV>(ma, mb, p, s);
    }
    return kBadArgument;
}
static int dispatch_conv256(int epi, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t s) {
    switch (epi) {
        case MMADA_EPI_BIAS_BF16: return launch_gemm<2, MMADA_EPI_BIAS_BF16, 256, true>(ma, mb, p, s);
        case MMADA_EPI_BIAS_F32: return launch_gemm<2, MMADA_EPI_BIAS_F32, 256, true>(ma, mb, p, s);
        case MMADA_EPI_BIAS_RESID_F32: return launch_gemm<2, MMADA_EPI_BIAS_RESID_F32, 256, true>(ma, mb, p, s);
    }
    return kBadArgument;
}

}  // namespace mmada

using namespace mmada;

// rasterisation / cache hints.  Defaults measured in round 1 (profiles/r01b_gemm_traffic_probe.txt); an EXPERIMENTS=1
// build can override them from the environment (MMADA_GEMM_GROUP_M, MMADA_GEMM_HINTS=ab with a,b in {n,f,l} = normal /
// evict-first / evict-last for the A and B operand loads) — the product library reads no environment variables.
static void set_tuning(GemmParams& p) {
    int group_m = 0;
    uint64_t ha = kEvictLast, hb = kEvictNormal;          // measured: +2 % sustained vs (8, normal/normal)
#ifdef MMADA_EXPERIMENTS
    static int env_group_m = -1;
    static uint64_t env_ha = kEvictLast, env_hb = kEvictNormal;
    if (env_group_m < 0) {
        const char* g = getenv("MMADA_GEMM_GROUP_M");
        env_group_m = g ? atoi(g) : 0;
        if (env_group_m < 0) env_group_m = 0;
        const char* h = getenv("MMADA_GEMM_HINTS");
        auto dec = [](char c) { return c == 'f' ? kEvictFirst : (c == 'l' ? kEvictLast : kEvictNormal); };
        if (h && h[0] && h[1]) { env_ha = dec(h[0]); env_hb = dec(h[1]); }
    }
    group_m = env_group_m; ha = env_ha; hb = env_hb;
#endif
    // default: 16 m-tiles per group (the A band, 16 x 256 rows x K, stays in L2 while the n-tiles sweep past it);
    // for long K the band outgrows L2 and a squarer wave re-reads less (K = 12288: 3.7 GB at 8 against 4.1 GB at 16;
    // K = 4096: 2.5 GB at 16 against 3.5 GB at 8)
    // round 2 (profiles/r02c_gemm_by_n_probe.txt): grouping the N-tiles instead (a band of B = weights stays in L2 while
    // the m-tiles sweep, odd groups backwards) re-reads less for every block shape: the operand that is streamed once
    // per group is then read ceil(Nt / g) times instead of ceil(Mt / g) — ff_out (K = 12288) 3.94 -> 3.56 GB at g = 8,
    // q|k|v 1.70 -> 1.57 at g = 12, attn_out 1.39 -> 1.25 at g = 8, gate|up 2.70 -> 2.58 at g = 16
    const int by_n = experiment_env("MMADA_GEMM_BY_N", -1);
    p.serpentine = experiment_env("MMADA_GEMM_SERPENTINE", 1);
    if (by_n != 0) {
        p.group_by_n = 1;
        p.group_m = group_m ? group_m : (p.K >= 8192 || p.N <= 4096 ? 8 : (p.N <= 12288 ? 12 : 16));
#ifdef MMADA_EXPERIMENTS
        if (!getenv("MMADA_GEMM_HINTS"))
#endif
        { ha = kEvictNormal; hb = kEvictLast; }     // the resident operand is B now (285.3 -> 284.3 ms per step, same-box ABAB)
    } else {
        p.group_by_n = 0;
        p.group_m = group_m ? group_m : (p.K >= 8192 ? 8 : 16);
    }
    p.hint_a = ha; p.hint_b = hb;
}

// row_ssq / ssq_tiles / norm_dim / eps: folded RMSNorm on the consumer side (SwiGLU epilogue); xb / ld_xb / ssq_out:
// on the producer side (epilogue EPI_RESID_NORM_F32)
static int gemm_entry(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo, const void* aux,
                      const float* bias, int M, int N, int K, int epilogue, int cta_group, void* stream,
                      const float* row_ssq, int ssq_tiles, int norm_dim, float eps, void* xb, int64_t ld_xb, float* ssq_out) {
    if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0) return kBadArgument;
    if ((lda % 8) || (ldb % 8) || (K % 8)) return kUnsupportedShape;        // 16-byte global strides for TMA
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) return kBadArgument;
    if (epilogue < 0 || (epilogue > MMADA_EPI_BIAS_RESID_F32 && epilogue != EPI_RESID_NORM_F32)) return kBadArgument;   // ROPE has its own entry point
    if (epilogue == EPI_RESID_NORM_F32) {
        if (!xb || !ssq_out) return kBadArgument;
        if ((N % 256) || (ldo % 4) || (ld_xb % 8)) return kUnsupportedShape;     // whole tiles, vector stores
        if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(aux) | reinterpret_cast<uintptr_t>(xb)) & 15) return kBadArgument;
    }
    if (row_ssq && (epilogue != MMADA_EPI_SWIGLU_BF16 || ssq_tiles <= 0 || norm_dim <= 0)) return kBadArgument;
    if (epilogue == MMADA_EPI_SWIGLU_BF16 && (N % 256)) return kUnsupportedShape;
    if (epi_has_resid(epilogue) && !aux) return kBadArgument;
    if (epi_has_bias(epilogue) && !bias) return kBadArgument;
    if (cta_group != 1 && cta_group != 2) return kBadArgument;
    const bool narrow = N <= 128 && cta_group == 2 && epilogue != MMADA_EPI_SWIGLU_BF16 && epilogue != MMADA_EPI_RESID_F32;
    const int bn = narrow ? 128 : 256;
    CUtensorMap ma, mb;
    int st = make_tmap_bf16_2d(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM);
    if (st) return st;
    st = make_tmap_bf16_2d(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)(bn / cta_group));
    if (st) return st;
    GemmParams p = {};
    p.out = out; p.aux = aux; p.bias = bias; p.ldo = ldo;
    p.M = M; p.N = N; p.K = K;
    p.num_m_tiles = (M + BM * cta_group - 1) / (BM * cta_group);
    p.num_n_tiles = (N + bn - 1) / bn;
    p.row_ssq = row_ssq; p.ssq_tiles = ssq_tiles; p.ssq_inv_d = norm_dim > 0 ? 1.0f / (float)norm_dim : 0.f; p.ssq_eps = eps;
    p.xb_out = reinterpret_cast<__nv_bfloat16*>(xb); p.ld_xb = ld_xb; p.ssq_out = ssq_out;
    set_tuning(p);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (narrow) return dispatch_narrow<false>(epilogue, ma, mb, p, s);
    return cta_group == 1 ? dispatch_gemm256<1>(epilogue, ma, mb, p, s) : dispatch_gemm256<2>(epilogue, ma, mb, p, s);
}

extern "C" int mmada_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                               const void* aux, const float* bias, int M, int N, int K, int epilogue, int cta_group,
                               void* stream) {
    if (epilogue > MMADA_EPI_BIAS_RESID_F32) return kBadArgument;
    return gemm_entry(A, lda, B, ldb, out, ldo, aux, bias, M, N, K, epilogue, cta_group, stream, nullptr, 0, 0, 0.f, nullptr, 0,
                      nullptr);
}

// x fp32 [M,N] += A . B^T;  xb bf16 [M,N] = bf16(x);  ssq_out fp32 [M, N/256]: per 256-column tile, the row's sum of x^2
extern "C" int mmada_gemm_resid_norm_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* x, int64_t ldx,
                                         void* xb_bf16, int64_t ld_xb, float* ssq_out, int M, int N, int K, int cta_group,
                                         void* stream) {
    return gemm_entry(A, lda, B, ldb, x, ldx, x, nullptr, M, N, K, EPI_RESID_NORM_F32, cta_group, stream, nullptr, 0, 0, 0.f,
                      xb_bf16, ld_xb, ssq_out);
}

// out bf16 [M,N/2] = silu(r*gate) * (r*up), r[m] = rsqrt(sum(row_ssq[m, 0..ssq_tiles)) / norm_dim + eps): SwiGLU epilogue
// with the RMSNorm in front of the projection folded in (its weight multiplied into B's columns by the caller)
extern "C" int mmada_gemm_swiglu_rownorm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                                              const float* row_ssq, int ssq_tiles, int norm_dim, float eps, int M, int N,
                                              int K, int cta_group, void* stream) {
    if (!row_ssq) return kBadArgument;
    return gemm_entry(A, lda, B, ldb, out, ldo, nullptr, nullptr, M, N, K, MMADA_EPI_SWIGLU_BF16, cta_group, stream, row_ssq,
                      ssq_tiles, norm_dim, eps, nullptr, 0, nullptr);
}

// Fused q|k|v projection + rotary embedding: out bf16 [M, N] = A . B^T with RoPE applied to the heads inside
// columns [0, rope_cols) (q and k), position = row % seq_len.
static int qkv_rope_entry(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                          const float* sin_table, const float* cos_table, int M, int N, int K, int rope_cols, int head_dim,
                          int seq_len, int cta_group, void* stream, const float* row_ssq, int ssq_tiles, int norm_dim,
                          float eps) {
    if (row_ssq && (ssq_tiles <= 0 || norm_dim <= 0)) return kBadArgument;
    if (!A || !B || !out || !sin_table || !cos_table || M <= 0 || N <= 0 || K <= 0 || seq_len <= 0) return kBadArgument;
    if ((lda % 8) || (ldb % 8) || (K % 8) || (ldo % 8) || (N % 256)) return kUnsupportedShape;
    if ((head_dim != 64 && head_dim != 128) || rope_cols % head_dim || rope_cols > N) return kUnsupportedShape;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(out) |
         reinterpret_cast<uintptr_t>(sin_table) | reinterpret_cast<uintptr_t>(cos_table)) & 15)
        return kBadArgument;
    if (cta_group != 1 && cta_group != 2) return kBadArgument;
    CUtensorMap ma, mb;
    int st = make_tmap_bf16_2d(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM);
    if (st) return st;
    st = make_tmap_bf16_2d(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)(256 / cta_group));
    if (st) return st;
    GemmParams p = {};
    p.out = out; p.ldo = ldo;
    p.M = M; p.N = N; p.K = K;
    p.num_m_tiles = (M + BM * cta_group - 1) / (BM * cta_group);
    p.num_n_tiles = N / 256;
    p.rope_sin = sin_table; p.rope_cos = cos_table; p.rope_hd = head_dim; p.rope_L = seq_len; p.rope_cols = rope_cols;
    p.row_ssq = row_ssq; p.ssq_tiles = ssq_tiles; p.ssq_inv_d = norm_dim > 0 ? 1.0f / (float)norm_dim : 0.f; p.ssq_eps = eps;
    set_tuning(p);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return cta_group == 1 ? dispatch_gemm256<1>(MMADA_EPI_ROPE_BF16, ma, mb, p, s)
                          : dispatch_gemm256<2>(MMADA_EPI_ROPE_BF16, ma, mb, p, s);
}

extern "C" int mmada_gemm_qkv_rope_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                                        const float* sin_table, const float* cos_table, int M, int N, int K, int rope_cols,
                                        int head_dim, int seq_len, int cta_group, void* stream) {
    return qkv_rope_entry(A, lda, B, ldb, out, ldo, sin_table, cos_table, M, N, K, rope_cols, head_dim, seq_len, cta_group,
                          stream, nullptr, 0, 0, 0.f);
}

// the same with the RMSNorm in front of the projection folded in (see mmada_gemm_swiglu_rownorm_bf16)
extern "C" int mmada_gemm_qkv_rope_rownorm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out,
                                                int64_t ldo, const float* sin_table, const float* cos_table,
                                                const float* row_ssq, int ssq_tiles, int norm_dim, float eps, int M, int N,
                                                int K, int rope_cols, int head_dim, int seq_len, int cta_group,
                                                void* stream) {
    if (!row_ssq) return kBadArgument;
    return qkv_rope_entry(A, lda, B, ldb, out, ldo, sin_table, cos_table, M, N, K, rope_cols, head_dim, seq_len, cta_group,
                          stream, row_ssq, ssq_tiles, norm_dim, eps);
}

// NHWC convolution as an implicit GEMM: out[b,y,x,co] = bias[co] + sum_{tap,c} in[b,y+dy,x+dx,c] * w[co,tap,c] (+ resid)
extern "C" int mmada_conv_nhwc_bf16(const void* in, const void* weight, const float* bias, void* out, const void* resid,
                                    int B, int H, int W, int C_in, int C_out, int taps, int epilogue, void* stream) {
    if (!in || !weight || !bias || !out || B <= 0 || H <= 0 || W <= 0) return kBadArgument;
    if (taps != 9 && taps != 1) return kBadArgument;
    if (C_in % 64) return kUnsupportedShape;                       // whole k-blocks per tap
    const int bw = W >= 128 ? 128 : W;
    if (128 % bw || W % bw) return kUnsupportedShape;
    const int bh = 128 / bw;
    if (H % bh) return kUnsupportedShape;
    if (!epi_has_bias(epilogue)) return kBadArgument;
    if (epi_has_resid(epilogue) && !resid) return kBadArgument;
    CUtensorMap ma, mb;
    const uint64_t dims[4] = {(uint64_t)C_in, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C_in * 2, (uint64_t)W * C_in * 2, (uint64_t)H * W * C_in * 2};
    const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
    int st = make_tmap(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in, dims, strides, box);
    if (st) return st;
    const int K = taps * C_in;
    const int bn = C_out <= 128 ? 128 : 256;
    st = make_tmap_bf16_2d(&mb, weight, (uint64_t)C_out, (uint64_t)K, (uint64_t)K, (uint32_t)(bn / 2));
    if (st) return st;
    GemmParams p = {};
    p.out = out; p.aux = resid; p.bias = bias; p.ldo = C_out;
    p.M = B * H * W; p.N = C_out; p.K = K;
    p.num_m_tiles = (p.M + 255) / 256;
    p.num_n_tiles = (C_out + bn - 1) / bn;
    p.conv_H = H; p.conv_W = W; p.conv_C = C_in; p.conv_taps = taps;
    set_tuning(p);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return bn == 128 ? dispatch_narrow<true>(epilogue, ma, mb, p, s) : dispatch_conv256(epilogue, ma, mb, p, s);
}
