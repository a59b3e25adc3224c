// Microbenchmark: does a packed half-precision ex2 (ex2.approx.f16x2 / .ftz.bf16x2) deliver two exponentials per MUFU
// slot?  (No: cuobjdump shows ptxas splitting every packed ex2 into two scalar MUFU.EX2.F16 / .BF16 operations, so the
// softmax cannot halve its MUFU work that way — not run on the GPU for that reason.)  Same harness as mufu_rate.cu: one CTA on one SM, 8 independent chains per thread, clock64 around the loop.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = 0xb800b800u + threadIdx.x + i;        // (-0.5, -0.5) as f16x2, perturbed
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { float f = __uint_as_float(v[i]); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); v[i] = __float_as_uint(f); }
            if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
            if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
            if (MODE == 3) { unsigned short h = (unsigned short)v[i]; asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h)); v[i] = h; }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8);
    const char* names[4] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "ex2.approx.f16"};
    const int iters = 2000;
    for (int mode = 0; mode < 4; ++mode)
        for (int wps = 1; wps <= 4; wps *= 2) {
            const int threads = wps * 4 * 32;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<1, threads>>>(out, cyc, iters);
                if (mode == 1) k<1><<<1, threads>>>(out, cyc, iters);
                if (mode == 2) k<2><<<1, threads>>>(out, cyc, iters);
                if (mode == 3) k<3><<<1, threads>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            const double per = (double)cyc[0] / (iters * 8.0 * wps);
            printf("%-24s %d warps/scheduler: %6.2f clk per warp-instruction per scheduler\n", names[mode], wps, per);
        }
    return 0;
}
