// xcorr_fft.cu -- K1/K2 in the transform domain for the shapes where the direct correlation is FMA-bound (29x29 and 15x15 templates).
//
// Replaces hdn/core/xcorr.py:37-46 / :48-61 at 256/512 crops (61x61 (*) 29x29, 29x29 circular (*) 29x29) and the 15x15
// large-displacement window (39x39 (*) 15x15).  Algorithm and phase functions: xcorr_fft.cuh; 64-point FFT: fft64.cuh.
//
// Kernel structure (persistent CTAs, two or three per SM; G = 2 planes per group, 4-6 warps per CTA):
//   * the x and k planes of a group are two contiguous byte ranges -> two 1-D TMA bulk copies (UBLKCP) onto an mbarrier.  Two
//     planes of odd size start 0 or 8 bytes past a 16-byte boundary, so the copy fetches the enclosing 16-byte-aligned window (it
//     stays inside the tensor: C % 4 == 0 makes the tensor's own ends aligned) and the phases index from the offset.  The NEXT
//     group's copies are issued as soon as phase R has consumed the landing buffer, so they fly during the column stage;
//   * phase R (row FFTs) -> column stage (direct complex correlation per frequency column, a dense FFMA2 loop) -> phase O
//     (inverse row FFTs), separated by __syncthreads().  An FFT task = one half of a 64-point FFT, register-resident; R and O
//     share ONE copy of the half-FFT code (the phase only selects the load / store code around it) -- a fully specialised
//     straight-line kernel (one FFT body per phase, 130 KB of SASS) spent half of its issue slots waiting for instruction fetch;
//   * the CTAs of an SM run unsynchronised, so one CTA's shared-memory-bound load section overlaps another's arithmetic;
//   * the finished tile (staged over the dead row-spectrum buffer) is copied out with coalesced stores.
// HBM traffic is the algorithmic bytes (each plane read once, each output written once).
#include "common.cuh"
#include "xcorr_fft.cuh"

namespace hdn {

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::CTAS)
    xcorr_fft_kernel(XProblems P, int groups_per_problem, int n_groups, int C, long long k_bstride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *raw = reinterpret_cast<float *>(smem_raw);
    float2 *XR = reinterpret_cast<float2 *>(raw + Cfg::RAW_FLOATS);
    float2 *KR = XR + Cfg::G * Cfg::XR_PLANE;
    float2 *CT = KR + Cfg::G * Cfg::KR_PLANE;
    uint64_t *full = reinterpret_cast<uint64_t *>(CT + Cfg::G * Cfg::CT_PLANE);
    float *so = reinterpret_cast<float *>(XR);  // output tile: XR is dead once the column stage is done
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // group g -> problem, element offsets of its x / k / out planes
    auto locate = [&](int g, int &prob, long long &xoff, long long &koff, long long &ooff) {
        prob = g / groups_per_problem;
        const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
        const long long b = plane0 / C, c0 = plane0 - b * C;
        xoff = plane0 * Cfg::XPL;
        koff = b * k_bstride + c0 * Cfg::KPL;
        ooff = plane0 * Cfg::OPL;
    };
    auto issue = [&](int g) {  // elected thread only
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        mbar_expect_tx(full, Cfg::RAW_FLOATS * 4);
        bulk_g2s(raw, P.x[prob] + (xoff & ~3ll), Cfg::XWIN * 4, full);
        bulk_g2s(raw + Cfg::XWIN, P.k[prob] + (koff & ~3ll), Cfg::KWIN * 4, full);
    };
    if (tid == 0 && (int)blockIdx.x < n_groups) issue(blockIdx.x);

    int it = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        const FftBufs bufs{raw + (int)(xoff & 3), raw + Cfg::XWIN + (int)(koff & 3), XR, KR, CT, so};
        mbar_wait(full, it & 1);
#pragma unroll 1
        for (int ph = 0; ph < FFT_PHASES; ++ph) {
            const int ntask = fftc_tasks<Cfg>(ph);
#pragma unroll 1
            for (int t0 = 0; t0 < ntask; t0 += Cfg::NT) {
                const int t = t0 + tid;
                const int h = fft_task_half<Cfg>(ph, t), unit = fft_task_unit<Cfg>(ph, t);
                float re[32], im[32];
                if (t < ntask && fftc_load<Cfg>(ph, bufs, unit, h, re, im)) {
                    if (h) fft::half_twiddle(re, im);
                    fft::fft32_fwd(re, im);
                    fftc_store<Cfg>(ph, bufs, unit, h, re, im);
                }
            }
            __syncthreads();
            if (ph == FFT_PH_R) {
                if (tid == 0) {  // landing buffer consumed
                    const int gn = g + gridDim.x;
                    if (gn < n_groups) issue(gn);
                }
#pragma unroll 1
                for (int t = tid; t < Cfg::COL_TASKS; t += Cfg::NT) fftc_col<Cfg>(bufs, t);
                __syncthreads();
            }
        }
        float *dst = P.out[prob] + ooff;
#pragma unroll 2
        for (int e = tid; e < Cfg::OUT_FLOATS; e += Cfg::NT) dst[e] = so[e];
        __syncthreads();  // the tile aliases XR, which the next group's phase R writes
    }
}

template <class Cfg>
static int launch_fft(const XProblems &P, int n, int B, int C, long long kbs, cudaStream_t st) {
    static DeviceOnce once;
    if (int e = once.run([] { return cudaFuncSetAttribute(xcorr_fft_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM); }))
        return e;
    const int gpp = (int)(((long long)B * C) / Cfg::G);
    const int total = gpp * n;
    const int slots = sm_count() * Cfg::CTAS;
    const int grid = total < slots ? total : slots;
    xcorr_fft_kernel<Cfg><<<grid, Cfg::NT, Cfg::SMEM, st>>>(P, gpp, total, C, kbs);
    count_launch();
    return launch_status();
}

//                     KH  KW  HX  WX  circ  G   NT
// threads per CTA (tunable at build time for A/B runs): the column stage splits a plane's output rows over NT / 64 warps
#ifndef HDN_FFT_NT1
#define HDN_FFT_NT1 192
#endif
#ifndef HDN_FFT_NT2
#define HDN_FFT_NT2 128
#endif
#ifndef HDN_FFT_NT3
#define HDN_FFT_NT3 128
#endif
#ifndef HDN_FFT_G1
#define HDN_FFT_G1 2
#endif
using F256 = FCfg<29, 29, 61, 61, false, HDN_FFT_G1, HDN_FFT_NT1>;    // 256/512 crops, similarity branch
using F256Lp = FCfg<29, 29, 29, 29, true, 2, HDN_FFT_NT2>;   // 256/512 crops, log-polar branch (INSTANCE_SIZE = 512)
using FWin15 = FCfg<15, 15, 39, 39, false, 2, HDN_FFT_NT3>;  // 15x15 large-displacement window

#define HDN_FFT_SHAPES(X) X(F256) X(F256Lp) X(FWin15)

bool xcorr_fft_applicable(int C, int Hx, int Wx, int Hk, int Wk, int circular) {
#define HDN_IS(CFG) \
    if (Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % 4 == 0) return true;
    HDN_FFT_SHAPES(HDN_IS)
#undef HDN_IS
    return false;
}

int xcorr_fft_dispatch(const XProblems &P, int n, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular, long long kbs, cudaStream_t st) {
#define HDN_TRY(CFG) \
    if (Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % 4 == 0) \
        return launch_fft<CFG>(P, n, B, C, kbs, st);
    HDN_FFT_SHAPES(HDN_TRY)
#undef HDN_TRY
    return HDN_ERR_UNSUPPORTED;
}

}  // namespace hdn
