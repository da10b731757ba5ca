// xcorr_fft.cu -- K1/K2 by 64x64 FFT for the shapes where the direct correlation is FMA-bound (29x29 and 15x15 templates).
//
// Replaces hdn/core/xcorr.py:37-46 / :48-61 at 256/512 crops (61x61 (*) 29x29, 29x29 circular (*) 29x29) and the 15x15
// large-displacement window (39x39 (*) 15x15).  Algorithm and phase functions: xcorr_fft.cuh; 64-point FFT: fft64.cuh.
//
// Kernel structure (one persistent CTA per SM, G = 4 planes per group, 256 threads = 8 warps):
//   * the x and k planes of a group are two contiguous byte ranges -> two 1-D TMA bulk copies (UBLKCP) onto an mbarrier;
//     the NEXT group's copies are issued as soon as phase R has consumed the landing buffer, so they fly during the column phases;
//   * five phases R -> CX -> CK -> CI -> O separated by __syncthreads(); a task = one half of a 64-point FFT, register-resident;
//     all phases share ONE copy of the half-FFT code (the phase only selects the load / store code around it), so the hot loop
//     stays resident in the instruction cache -- a fully specialised straight-line kernel (one FFT body per phase, 130 KB of
//     SASS) spent half of its issue slots waiting for instruction fetch;
//   * the finished G x HO x WO tile leaves through a double-buffered TMA bulk store.
// HBM traffic is exactly the algorithmic bytes (each plane read once, each output written once).
#include "common.cuh"
#include "xcorr_fft.cuh"

namespace hdn {

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 1)
    xcorr_fft_kernel(XProblems P, int groups_per_problem, int n_groups, int C, long long k_bstride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *raw = reinterpret_cast<float *>(smem_raw);
    float *sout = raw + Cfg::RAW_FLOATS;
    float2 *XR = reinterpret_cast<float2 *>(sout + 2 * Cfg::OUT_FLOATS);
    float2 *KR = XR + Cfg::G * Cfg::XR_PLANE;
    uint64_t *full = reinterpret_cast<uint64_t *>(KR + Cfg::G * Cfg::KR_PLANE);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int g) {  // elected thread only
        const int prob = g / groups_per_problem;
        const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
        const long long b = plane0 / C, c0 = plane0 - b * C;
        mbar_expect_tx(full, Cfg::RAW_FLOATS * 4);
        bulk_g2s(raw, P.x[prob] + plane0 * Cfg::XPL, Cfg::G * Cfg::XPL * 4, full);
        bulk_g2s(raw + Cfg::G * Cfg::XPL, P.k[prob] + b * k_bstride + c0 * Cfg::KPL, Cfg::G * Cfg::KPL * 4, full);
    };
    if (tid == 0 && (int)blockIdx.x < n_groups) issue(blockIdx.x);

    int it = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
        mbar_wait(full, it & 1);
        const FftBufs bufs{raw, raw + Cfg::G * Cfg::XPL, XR, KR, sout + (it & 1) * Cfg::OUT_FLOATS};
#pragma unroll 1
        for (int ph = 0; ph < FFT_PHASES; ++ph) {
            const int ntask = fftc_tasks<Cfg>(ph);
#pragma unroll 1
            for (int t0 = 0; t0 < ntask; t0 += Cfg::NT) {
                const int t = t0 + tid;
                const int h = fft_task_half(t), unit = fft_task_unit(t);
                float re[64], im[64];
                const bool active = t < ntask && fftc_load<Cfg>(ph, bufs, unit, re, im);
                if (ph == FFT_PH_CX) __syncthreads();  // columns are transformed in place: both halves have read before either writes
                if (active) {
                    fft::half_butterfly(h, re, im);
                    fft::fft32_fwd(re, im);
                    if (h == 0) fftc_store<Cfg, 0>(ph, bufs, unit, re, im);
                    else fftc_store<Cfg, 1>(ph, bufs, unit, re, im);
                }
            }
            if (ph == FFT_PH_O) fence_proxy_async_smem();  // my output-tile writes -> visible to the TMA store
            __syncthreads();
            if (tid == 0) {
                if (ph == FFT_PH_R) {  // landing buffer consumed
                    const int gn = g + gridDim.x;
                    if (gn < n_groups) issue(gn);
                    bulk_wait_read<1>();  // the store issued two groups ago has left the tile buffer phase O of this group fills
                } else if (ph == FFT_PH_O) {
                    const int prob = g / groups_per_problem;
                    const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
                    bulk_s2g(P.out[prob] + plane0 * Cfg::OPL, bufs.out, Cfg::OUT_FLOATS * 4);
                    bulk_commit();
                }
            }
        }
    }
    if (tid == 0) bulk_wait_all<0>();
}

template <class Cfg>
static int launch_fft(const XProblems &P, int n, int B, int C, long long kbs, cudaStream_t st) {
    static bool configured = false;  // benign race: idempotent attribute set
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(xcorr_fft_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int gpp = (int)(((long long)B * C) / Cfg::G);
    const int total = gpp * n;
    const int grid = total < sm_count() ? total : sm_count();
    xcorr_fft_kernel<Cfg><<<grid, Cfg::NT, Cfg::SMEM, st>>>(P, gpp, total, C, kbs);
    count_launch();
    return launch_status();
}

//                     KH  KW  HX  WX  circ  G   NT
using F256 = FCfg<29, 29, 61, 61, false, 4, 256>;    // 256/512 crops, similarity branch
using F256Lp = FCfg<29, 29, 29, 29, true, 4, 256>;   // 256/512 crops, log-polar branch (INSTANCE_SIZE = 512)
using FWin15 = FCfg<15, 15, 39, 39, false, 4, 256>;  // 15x15 large-displacement window

#define HDN_FFT_SHAPES(X) X(F256) X(F256Lp) X(FWin15)

bool xcorr_fft_applicable(int C, int Hx, int Wx, int Hk, int Wk, int circular) {
#define HDN_IS(CFG) \
    if (Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % CFG::G == 0) return true;
    HDN_FFT_SHAPES(HDN_IS)
#undef HDN_IS
    return false;
}

int xcorr_fft_dispatch(const XProblems &P, int n, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular, long long kbs, cudaStream_t st) {
#define HDN_TRY(CFG) \
    if (Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % CFG::G == 0) \
        return launch_fft<CFG>(P, n, B, C, kbs, st);
    HDN_FFT_SHAPES(HDN_TRY)
#undef HDN_TRY
    return HDN_ERR_UNSUPPORTED;
}

}  // namespace hdn
