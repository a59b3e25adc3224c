// Legacy mma.sync (HMMA) issue rate on sm_100a: clocks per m16n8k16 bf16 mma.sync per scheduler, 1 / 2 / 4 warps per
// scheduler, 4 independent accumulator chains per warp.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 hmma_rate.cu
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
__global__ void k(float* out, long long* clk, int iters) {
    float c[4][4] = {};
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
    a[0] += threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][2] + c[3][3];
}
int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 1024);
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        k<<<1, warps * 32>>>(out, clk, iters);
        k<<<1, warps * 32>>>(out, clk, iters);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("mma.sync m16n8k16 bf16: %2d warps/SM (%d per scheduler): %.2f clk per mma per scheduler  (%.0f FLOP/clk/SM)\n", warps,
               warps / 4, (double)h / (iters * 4.0 * (warps / 4)), 4096.0 * iters * 4 * warps / (double)h);
    }
    return 0;
}
