// umma.cuh -- tcgen05 / TMEM helpers and the argument block shared by the tensor-core convolution kernels (conv_gemm.cu, conv_shift.cu).
#pragma once
#include "common.cuh"

namespace hdn {

constexpr int CG_BM = 128, CG_BK = 32, CG_THREADS = 256;
#ifndef HDN_CG_KCB
#define HDN_CG_KCB 8
#endif
constexpr int CG_KCB = HDN_CG_KCB;  // K blocks per TMEM accumulation chunk (8 x 32 = 256 of K)

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=0 (no swizzle) [61,64)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t zero = 0;  // disable-output-lane mask: all lanes enabled
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(zero)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// Up to HDN_MAX_PROBLEMS same-shape convolutions per launch (the 3 levels x {cls, loc} branches of a BAN head share one shape):
// blockIdx.z = problem * B + image.
struct ConvGemmArgs {
    const float *x[HDN_MAX_PROBLEMS], *wpk[HDN_MAX_PROBLEMS], *scale[HDN_MAX_PROBLEMS], *shift[HDN_MAX_PROBLEMS],
        *residual[HDN_MAX_PROBLEMS];  // wpk: hdn_conv_pack_weight_f32 output
    float *out[HDN_MAX_PROBLEMS];
    const float *w2[HDN_MAX_PROBLEMS];  // PROJECT mode: second 1x1 convolution [L, Cout] row-major (device); out = partial sums
    int B, L;
    int splitk;  // > 1: a cluster of `splitk` CTAs shares one output tile, each taking 1/splitk of K (blockIdx.x = tile * splitk + rank)
    int Cin, Cout, H, W, taps, dil, relu;
    int Ho, Wo, off;  // output extent; input pixel of output (r, c) under the CENTRE tap = (r * stride + off, c * stride + off):
                      // off = (ksize / 2) * dil - pad   ('same' 3x3: 0; 'valid' 3x3: d; 1x1: -pad)
    int stride;       // 1 or 2
};

// conv_gemm_ts.cu: the large launches with the activations in tensor memory (A operand from TMEM)
int launch_conv_gemm_ts(const ConvGemmArgs &a, int nprob, cudaStream_t st);
// conv_shift.cu: 3x3 'valid' convolutions with the activations staged ONCE per channel block (9 shifted operand windows).
bool conv_shift_applicable(const ConvGemmArgs &a, int ksize, int valid);
int launch_conv_shift(const ConvGemmArgs &a, int nprob, cudaStream_t st);

}  // namespace hdn
