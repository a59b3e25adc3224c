// Softmax arithmetic shared by the attention kernels: MUFU / FMA-pipe exponentials, packed fp32x2 ops.
#pragma once
#include "common.cuh"

namespace mmada {

// 2^x on the MUFU pipe (one instruction; flush-to-zero, -inf -> 0)
__device__ __forceinline__ float ex2_mufu(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// packed fp32x2 arithmetic (sm_100): two lanes per instruction
__device__ __forceinline__ uint64_t pk2(float2 a) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 upk2(uint64_t r) {
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
    return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
    return upk2(rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    uint64_t rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(rd);
}
__device__ __forceinline__ float2 fadd2_rm(float2 a, float2 b) {     // round towards -inf
    uint64_t rd;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(rd);
}
// 2^x for two values on the FMA pipe: floor via the 1.5*2^23 magic add (round-down), degree-3 minimax
// polynomial of 2^f on [0,1), exponent patched in with an integer multiply-add.  Relative error ~1e-4 —
// P is rounded to bf16 (2^-9) anyway.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fminf(fmaxf(x.x, -126.0f), 127.0f);       // above 2^127 the exponent patch would wrap: saturate (callers that
    x.y = fminf(fmaxf(x.y, -126.0f), 127.0f);       // speculate on the range detect 2^127 in their row sums)
    const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
    const float2 r = fadd2_rm(x, magic);
    const float2 fl = fadd2(r, nmagic);                                   // floor(x), exact
    const float2 f = ffma2(fl, make_float2(-1.0f, -1.0f), x);             // x - floor(x) in [0,1)
    float2 p = ffma2(f, make_float2(0.077119089663028717f, 0.077119089663028717f),
                     make_float2(0.227564394474029541f, 0.227564394474029541f));
    p = ffma2(p, f, make_float2(0.695146143436431885f, 0.695146143436431885f));
    p = ffma2(p, f, make_float2(1.0f, 1.0f));
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23)));
}

}  // namespace mmada
