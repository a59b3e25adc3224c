// Host-side helpers shared by the launchers: error mapping, device attributes, TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace mmada {

// Status codes returned over the C ABI (0 = OK).  CUDA runtime errors are returned as 1000 + code.
enum Status : int {
    kOk = 0,
    kBadArgument = 1,
    kUnsupportedShape = 2,
    kDriverEntryPoint = 3,
    kTensorMap = 4,
};

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? kOk : 1000 + (int)e; }

#define MMADA_CUDA_TRY(expr)                           \
    do {                                               \
        cudaError_t _e = (expr);                       \
        if (_e != cudaSuccess) return cuda_status(_e); \
    } while (0)

constexpr int kMaxDevices = 64;

inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// SM count of the CURRENT device (cached per device: a process may drive several)
inline int num_sms() {
    static int cached[kMaxDevices] = {};
    const int dev = current_device();
    if (cached[dev] == 0) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    return cached[dev];
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: `done` is the launcher's per-kernel flag array
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kern, int bytes, bool (&done)[kMaxDevices]) {
    const int dev = current_device();
    if (done[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done[dev] = true;
    return e;
}

// Tuning switches read from the environment exist only in EXPERIMENTS=1 builds (csrc/Makefile); the product library
// always takes the default.
inline int experiment_env(const char* name, int dflt) {
#ifdef MMADA_EXPERIMENTS
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled encode_tiled_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Row-major tensor of `rank` dims; dims[0] is the innermost (contiguous) extent in ELEMENTS, strides
// are in BYTES for dims 1..rank-1.  128-byte swizzle, out-of-bounds elements read as zero.
inline int make_tmap(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled fn = encode_tiled_fn();
    if (!fn) return kDriverEntryPoint;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "mmada: cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu box %u %u\n", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
        return kTensorMap;
    }
    return kOk;
}

// 2-D bf16 matrix [rows][cols], row pitch `ld` elements, box = 64 columns x box_rows rows.
inline int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows) {
    uint64_t dims[2] = {cols, rows};
    uint64_t strides[1] = {ld * 2};
    uint32_t box[2] = {64, box_rows};
    return make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box);
}

}  // namespace mmada
