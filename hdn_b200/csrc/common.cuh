// common.cuh -- shared device helpers for libhdn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hdn_b200.h"

namespace hdn {

extern int64_t g_launches;  // api.cu
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HDN_OK : (int)e;
}

int sm_count();  // api.cu: SM count of the CURRENT device (cached per device)

// One-time per-DEVICE configuration of a kernel (cudaFuncSetAttribute applies to the current device only, so a process that
// drives several GPUs must opt in to > 48 KB of dynamic shared memory on each of them).  Benign race: the attribute set is idempotent.
struct DeviceOnce {
    unsigned long long done = 0;  // bit d = device d configured (devices >= 64 are simply re-configured every call)
    template <class F>
    int run(F &&configure) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return HDN_ERR_DEVICE;
        const unsigned long long bit = dev < 64 ? 1ull << dev : 0ull;
        if (bit && (__atomic_load_n(&done, __ATOMIC_ACQUIRE) & bit)) return HDN_OK;
        const cudaError_t e = configure();
        if (e != cudaSuccess) return (int)e;
        if (bit) __atomic_fetch_or(&done, bit, __ATOMIC_RELEASE);
        return HDN_OK;
    }
};

// Up to HDN_MAX_PROBLEMS same-shape correlation problems of one launch (device pointers, passed by value).
struct XProblems {
    const float *x[HDN_MAX_PROBLEMS];
    const float *k[HDN_MAX_PROBLEMS];
    float *out[HDN_MAX_PROBLEMS];
};
// xcorr_fft.cu: FFT correlation for the FMA-bound shapes.  Returns HDN_ERR_UNSUPPORTED when the shape has no FFT kernel.
int xcorr_fft_dispatch(const XProblems &P, int n, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular, long long kbs, int variant, cudaStream_t st);
bool xcorr_fft_applicable(int C, int Hx, int Wx, int Hk, int Wk, int circular);
// shared template with cached row spectra (xcorr_fft.cu): size of a template's spectra in floats (0: shape has no such kernel),
// the producer (P.k = templates, P.out = spectra) and the consumer (P.k = spectra)
long long xcorr_spectra_floats(int C, int Hx, int Wx, int Hk, int Wk, int circular);
int xcorr_spectra_dispatch(const XProblems &P, int n, int C, int Hx, int Wx, int Hk, int Wk, int circular, cudaStream_t st);
int xcorr_fft_spec_dispatch(const XProblems &P, int n, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular, int variant, cudaStream_t st);

// ---- mbarrier + bulk async copy (TMA, 1-D) -------------------------------------------------
// SASS: cp.async.bulk -> UBLKCP, expect_tx -> SYNCS.ARRIVE.TRANS64 (B200_PROFILING.md).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // try_wait suspends the thread in hardware until the phase completes or the time hint expires; with the default (short)
    // limit a waiting warp re-issued the poll ~20 times per wait and took issue slots from the warps doing the work.
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier. 16-B aligned src/dst/size.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by bulk async-groups.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// order generic-proxy smem writes before async-proxy (TMA) reads of the same bytes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace hdn
