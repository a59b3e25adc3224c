// Microbenchmark: MUFU.EX2 issue rate per scheduler (SMSP) with 1, 2, 4 warps per SMSP, against FFMA2 and the packed
// polynomial.  One CTA on one SM; every warp runs a loop of 8 independent chains; clock64 around the loop.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float v[8];
    uint64_t w[8];
    for (int i = 0; i < 8; ++i) { v[i] = -0.001f * (threadIdx.x + i); w[i] = (uint64_t)__float_as_uint(v[i]) * 0x100000001ull; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2(v[i]);
            if (MODE == 1) v[i] = rcp(v[i]);
            if (MODE == 2) w[i] = ffma2(w[i], w[(i + 1) & 7], w[i]);
            if (MODE == 3) v[i] = fmaf(v[i], v[(i + 1) & 7], v[i]);
        }
        if (MODE >= 4) {
            // the softmax inner loop's mix, per key pair: FFMA2 (scale, shift) -> 2 x MUFU.EX2 -> FADD2 (row sum) + F2FP (bf16 pack)
            uint32_t pk = 0;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                uint64_t x;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(v[i]), "f"(v[i + 1]));
                x = ffma2(x, w[0], w[1]);
                float a, b;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x));
                a = ex2(a); b = ex2(b);
                if (MODE >= 5) {
                    uint64_t e;
                    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(e) : "f"(a), "f"(b));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[2 + (i >> 1)]) : "l"(e));
                    uint32_t h = 0;
                    if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
                    if (MODE == 7) {        // integer pack: round half up, take the upper halves (ALU pipe only)
                        const uint32_t ua = __float_as_uint(a) + 0x8000u, ub = __float_as_uint(b) + 0x8000u;
                        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h) : "r"(ua), "r"(ub));
                    }
                    pk ^= h;
                }
                v[i] = a - 3.0f; v[i + 1] = b - 3.0f;
            }
            w[7] ^= pk;
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float((uint32_t)w[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8);
    const char* names[8] = {"MUFU.EX2 (ex2.approx.ftz.f32)", "MUFU.RCP", "FFMA2 (fma.rn.f32x2)", "FFMA", "FFMA2 + 2 MUFU.EX2 (per MUFU)", "FFMA2 + 2 MUFU + FADD2 + F2FP (per MUFU)", "FFMA2 + 2 MUFU + FADD2 (per MUFU)", "FFMA2 + 2 MUFU + FADD2 + int pack (per MUFU)"};
    const int iters = 2000;
    for (int mode = 0; mode < 8; ++mode)
        for (int wps = 1; wps <= 8; wps *= 2) {      // warps per scheduler
            const int threads = wps * 4 * 32;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<1, threads>>>(out, cyc, iters);
                if (mode == 1) k<1><<<1, threads>>>(out, cyc, iters);
                if (mode == 2) k<2><<<1, threads>>>(out, cyc, iters);
                if (mode == 3) k<3><<<1, threads>>>(out, cyc, iters);
                if (mode == 4) k<4><<<1, threads>>>(out, cyc, iters);
                if (mode == 5) k<5><<<1, threads>>>(out, cyc, iters);
                if (mode == 6) k<6><<<1, threads>>>(out, cyc, iters);
                if (mode == 7) k<7><<<1, threads>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            const double per = (double)cyc[0] / (iters * 8.0 * wps);       // clocks per warp-instruction per scheduler
            printf("%-32s %d warps/scheduler: %6.2f clk per warp-instruction per scheduler  (%5.1f lanes/clk/SM)\n", names[mode], wps, per, 128.0 / per);
        }
    return 0;
}
