// api.cu -- version / status / device helpers of the C ABI.
#include "common.cuh"

namespace hdn {
int64_t g_launches = 0;

int sm_count() {
    static int cached[64] = {0};  // per device: a process may drive several GPUs
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;  // B200
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}
}  // namespace hdn

extern "C" int hdn_abi_version(void) { return HDN_ABI_VERSION; }

extern "C" int64_t hdn_launch_count(void) { return __atomic_load_n(&hdn::g_launches, __ATOMIC_RELAXED); }

extern "C" const char *hdn_status_string(int status) {
    switch (status) {
        case HDN_OK: return "ok";
        case HDN_ERR_NULL: return "required pointer is NULL";
        case HDN_ERR_SHAPE: return "invalid shape";
        case HDN_ERR_ALIGN: return "pointer not 4-byte aligned";
        case HDN_ERR_UNSUPPORTED: return "unsupported argument combination";
        case HDN_ERR_DEVICE: return "no usable CUDA device";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown hdn status";
    }
}

extern "C" int hdn_device_info(int *sm, int *major, int *minor) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return HDN_ERR_DEVICE;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return HDN_ERR_DEVICE;
    if (sm) *sm = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return HDN_OK;
}
