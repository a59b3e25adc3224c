// Microbenchmark: cycles per tcgen05.mma for the operand configurations the attention kernels use.
// One CTA (or CTA pair) per SM issues `iters` back-to-back MMAs from one elected lane, commits, waits, and
// reports clock64 deltas.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mmada_b200/csrc
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace mmada;

struct Cfg { int cg, ts, M, N, b_mn, accs; };   // accs: number of distinct accumulators cycled through

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + 128 * 1024, tptr = bar + 16;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 32 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    if (warp == 0) { tmem_alloc<CG>(tptr, 512); tmem_relinquish<CG>(); }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 128 * 1024 + 16);
    if (warp == 0 && rank == 0) {
        const uint64_t kd = umma_desc_kmajor_sw128(0);
        const uint64_t vd = umma_desc_mnmajor_sw128(0, 16384);
        const uint32_t idesc = umma_idesc_bf16(c.M, c.N, 0, c.b_mn);
        const uint32_t a0 = sbase >> 4, b0 = (sbase + 64 * 1024) >> 4;
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            __syncwarp();
            t0 = clock64();
            if (elect_one()) {
                // descriptors of one group of 8 MMAs are loop-invariant: the loop body is 8 MMA instructions
                const uint32_t d0 = tmem + 256, d1 = tmem + 256 + ((c.accs > 1 ? c.N : 0) % 256);
                uint64_t ad[8], bd[8];
                uint32_t at[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t koff = (uint32_t)(u & 3) * 2;      // walk K inside the swizzle row like the kernels do
                    ad[u] = kd | (uint64_t)(a0 + koff);
                    bd[u] = c.ts ? ((c.b_mn ? vd : kd) | (uint64_t)(b0 + (c.b_mn ? u * 128 : koff))) : (kd | (uint64_t)(b0 + koff));
                    at[u] = tmem + u * 8;
                }
                if (c.ts) {
                    for (int i = 0; i < iters; i += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (CG == 1) umma_bf16_ts((u & 1) ? d1 : d0, at[u], bd[u], idesc, 1);
                            else umma_bf16_ts_cg<2>((u & 1) ? d1 : d0, at[u], bd[u], idesc, 1);
                        }
                    }
                } else {
                    for (int i = 0; i < iters; i += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) umma_bf16_ss<CG>((u & 1) ? d1 : d0, ad[u], bd[u], idesc, 1);
                    }
                }
                if (CG == 1) umma_commit(bar); else umma_commit_2sm(bar, 0x1);
            }
            __syncwarp();
            mbar_wait(bar, rep & 1, 99);
            t1 = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<CG>(tmem, 512); }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    const int smem = 128 * 1024 + 1024 + 64;
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const Cfg cfgs[] = {
        {1, 0, 128, 64, 0, 1},  {1, 0, 128, 128, 0, 1}, {1, 0, 128, 256, 0, 1}, {1, 0, 128, 128, 0, 2},
        {1, 1, 128, 128, 1, 1}, {1, 1, 128, 128, 0, 1}, {1, 1, 128, 64, 1, 1},  {1, 1, 128, 256, 1, 1},
        {2, 0, 256, 128, 0, 1}, {2, 0, 256, 256, 0, 1}, {2, 0, 256, 64, 0, 1},  {2, 0, 256, 128, 0, 2},
        {2, 1, 256, 128, 1, 1}, {2, 1, 256, 128, 0, 1}, {2, 1, 256, 256, 1, 1}, {2, 1, 256, 64, 1, 1},
    };
    const int iters = 512;
    for (const Cfg& c : cfgs) {
        for (int full = 0; full < 2; ++full) {        // one cluster alone vs all SMs busy
            out[0] = 0;
            cudaLaunchConfig_t lc = {};
            const int ctas = full ? 148 : c.cg;
            lc.gridDim = dim3(ctas); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c.cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            cudaError_t e = c.cg == 1 ? cudaLaunchKernelEx(&lc, rate_kernel<1>, c, iters, out)
                                      : cudaLaunchKernelEx(&lc, rate_kernel<2>, c, iters, out);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cfg failed: %s\n", cudaGetErrorString(e)); return 1; }
            const double cyc = (double)out[0] / iters;
            const double flop = 2.0 * c.M * c.N * 16 / c.cg;        // per SM
            printf("cta_group %d  %s  M=%3d N=%3d  B %s  accs %d  %s: %6.1f clk/MMA  -> %6.0f flop/clk/SM\n", c.cg,
                   c.ts ? "A=TMEM" : "A=SMEM", c.M, c.N, c.b_mn ? "MN-major" : "K-major ", c.accs,
                   full ? "148 SMs" : "alone  ", cyc, flop / cyc);
        }
    }
    return 0;
}
