// Microbenchmark: latency of the per-tile protocol operations of the attention softmax warps, one warp at a time on an
// otherwise idle SM (clock64 around 1000 dependent repetitions):
//   mbarrier.try_wait on a phase that completed long ago, tcgen05.st (16x128b.x8) + wait::st,
//   tcgen05.fence::before_thread_sync + __syncwarp + mbarrier.arrive, tcgen05.fence::after_thread_sync
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(float* out, long long* cyc, int iters) {
    __shared__ uint32_t tptr;
    __shared__ __align__(8) uint64_t bar_done, bar_sink;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&bar_sink)) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_done)) : "memory");   // phase 0 complete
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tptr;
    uint32_t acc = 0;
    if (warp == 0) {
        long long t[6];
        t[0] = clock64();
        for (int it = 0; it < iters; ++it) {            // try_wait on the completed phase 0, result consumed
            uint32_t done;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar_done) + (acc & 0)), "r"(0u) : "memory");
            acc += done;
        }
        t[1] = clock64();
        for (int it = 0; it < iters; ++it) {            // P store of one 64-key half + wait
            uint32_t v = acc + it;
            asm volatile("tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                         ::"r"(tmem + 256), "r"(v) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        t[2] = clock64();
        for (int it = 0; it < iters; ++it) {            // hand-over: fence, syncwarp, one arrive
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_sink)) : "memory");
        }
        t[3] = clock64();
        for (int it = 0; it < iters; ++it) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t[4] = clock64();
        for (int it = 0; it < iters; ++it) {            // test_wait instead of try_wait
            uint32_t done;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar_done) + (acc & 0)), "r"(0u) : "memory");
            acc += done;
        }
        t[5] = clock64();
        if (lane == 0) for (int i = 0; i < 5; ++i) cyc[i] = t[i + 1] - t[i];
    }
    out[threadIdx.x] = __uint_as_float(acc);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1024 * 4); cudaMallocManaged(&cyc, 64);
    const int iters = 1000;
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128>>>(out, cyc, iters);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
    }
    const char* names[5] = {"mbarrier.try_wait (completed phase)", "tcgen05.st 16x128b.x8 + wait::st", "fence::before + syncwarp + arrive",
                            "tcgen05.fence::after_thread_sync", "mbarrier.test_wait (completed phase)"};
    for (int i = 0; i < 5; ++i) printf("%-40s %7.1f clk\n", names[i], (double)cyc[i] / iters);
    return 0;
}
