// Probe: how fast can ONE CTA per 32 KiB row stream three [8192 x 8192] fp32 tensors (the access pattern of the t2i
// sampling kernel: cond, uncond, noise) when it does nothing else?  Variants: loads of all three rows up front, or the
// third row after a block barrier (like the kernel); 256 or 512 threads; 1 or 2 rows per CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 stream3_probe.cu -o stream3_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int THREADS, int ROWS, bool STAGGER>
__global__ void __launch_bounds__(THREADS) probe(const float4* __restrict__ a, const float4* __restrict__ b,
                                                 const float4* __restrict__ c, float* out) {
    constexpr int VEC = 2048 / THREADS;
    float s = 0.f;
    for (int r = 0; r < ROWS; ++r) {
        const size_t row = (size_t)blockIdx.x * ROWS + r;
        const float4 *pa = a + row * 2048, *pb = b + row * 2048, *pc = c + row * 2048;
        float4 x[VEC], y[VEC], z[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) x[i] = __ldcs(pa + i * THREADS + threadIdx.x);
#pragma unroll
        for (int i = 0; i < VEC; ++i) y[i] = __ldcs(pb + i * THREADS + threadIdx.x);
        if (STAGGER) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) s += x[i].x * y[i].y + x[i].z * y[i].w;
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) z[i] = __ldcs(pc + i * THREADS + threadIdx.x);
#pragma unroll
        for (int i = 0; i < VEC; ++i) s += x[i].x + y[i].y + z[i].z + z[i].w;
    }
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = s;
}
template <int THREADS, int ROWS, bool STAGGER>
void run(const char* name, float4* a, float4* b, float4* c, float* out) {
    const int rows = 8192;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) probe<THREADS, ROWS, STAGGER><<<rows / ROWS, THREADS>>>(a, b, c, out);
    cudaEventRecord(e0);
    for (int w = 0; w < 10; ++w) probe<THREADS, ROWS, STAGGER><<<rows / ROWS, THREADS>>>(a, b, c, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-44s %.4f ms  %.0f GB/s\n", name, ms, 3.0 * rows * 32768 / ms / 1e6);
}
int main() {
    float4 *a, *b, *c; float* out;
    const size_t n = (size_t)8192 * 32768;
    cudaMalloc(&a, n); cudaMalloc(&b, n); cudaMalloc(&c, n); cudaMalloc(&out, 8192 * 512 * 4);
    cudaMemset(a, 0, n); cudaMemset(b, 0, n); cudaMemset(c, 0, n);
    run<256, 1, false>("256 thr, 1 row/CTA, all loads up front", a, b, c, out);
    run<256, 1, true>("256 thr, 1 row/CTA, third row after a barrier", a, b, c, out);
    run<512, 1, false>("512 thr, 1 row/CTA, all loads up front", a, b, c, out);
    run<512, 1, true>("512 thr, 1 row/CTA, third row after a barrier", a, b, c, out);
    run<256, 2, false>("256 thr, 2 rows/CTA, all loads up front", a, b, c, out);
    run<256, 4, true>("256 thr, 4 rows/CTA, third row after a barrier", a, b, c, out);
    run<1024, 1, false>("1024 thr, 1 row/CTA, all loads up front", a, b, c, out);
    return 0;
}
