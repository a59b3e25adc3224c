// Library-level entry points of the C ABI.
#include "host_utils.h"
#include "../../include/mmada_b200.h"

extern "C" int mmada_abi_version(void) { return MMADA_ABI_VERSION; }

extern "C" int mmada_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
    return major * 10 + minor;
}
