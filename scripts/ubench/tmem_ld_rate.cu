// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM with 1, 2, 4, 8 warps reading concurrently, for the
// two shapes the attention kernels use (16x256b.x8 = m16n8 fragments of 16 lanes, 32x32b.x32 = one lane per thread).
// One CTA of 256 threads on one SM; warp w reads its own lane quarter (w & 3), 64 or 32 columns per instruction.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc, int iters, int nwarps) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tptr;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32 + (SHAPE == 0 ? (warp >> 2) * 16 : 0)) << 16);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {       // 4 instructions per iteration, different columns
                if (SHAPE == 0)
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(base + 64 * c) : "memory");
                else
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(base + 32 * c) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc ^= v[i];
            }
        }
        t1 = clock64();
    }
    out[threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    (void)lane;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1024 * 4); cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    for (int shape = 0; shape < 2; ++shape)
        for (int nw = 1; nw <= 8; nw *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (shape == 0) k<0><<<1, 256>>>(out, cyc, iters, nw);
                else k<1><<<1, 256>>>(out, cyc, iters, nw);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
            }
            // bytes per instruction: 16x256b.x8 = 16 lanes x 64 columns x 4 B = 4096; 32x32b.x32 = 32 lanes x 32 columns x 4 B = 4096
            const double per_ld = (double)cyc[0] / (iters * 4.0);
            printf("%-14s %d warps: %7.1f clk per tcgen05.ld+wait (one warp's view), %6.1f B/clk/SM\n", shape == 0 ? "16x256b.x8" : "32x32b.x32", nw,
                   per_ld, 4096.0 * nw / per_ld);
        }
    return 0;
}
