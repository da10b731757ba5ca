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
    uint64_t *full = reinterpret_cast<uint64_t *>(CT + Cfg::G * Cfg::CT_PLANE), *kfull = full + 1;
    float *so = reinterpret_cast<float *>(XR);  // output tile: XR is dead once the column stage is done
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_init(kfull, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // group g -> problem, element offsets of its x / k / out planes.  KSPEC: P.k holds the template's row spectra [C][KR_PLANE] complex
    // (one template for the whole batch), koff = float offset of the group's first channel.
    auto locate = [&](int g, int &prob, long long &xoff, long long &koff, long long &ooff) {
        prob = g / groups_per_problem;
        const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
        const long long b = plane0 / C, c0 = plane0 - b * C;
        xoff = plane0 * Cfg::XPL;
        koff = Cfg::KSPEC ? c0 * (2 * Cfg::KR_PLANE) : b * k_bstride + c0 * Cfg::KPL;
        ooff = plane0 * Cfg::OPL;
    };
    auto issue = [&](int g) {  // elected thread only
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        mbar_expect_tx(full, Cfg::RAW_FLOATS * 4);
        bulk_g2s(raw, P.x[prob] + (xoff & ~3ll), Cfg::XWIN * 4, full);
        if (!Cfg::KSPEC) bulk_g2s(raw + Cfg::XWIN, P.k[prob] + (koff & ~3ll), Cfg::KWIN * 4, full);
    };
    auto issue_k = [&](int g) {  // KSPEC, elected thread only: the group's spectra land directly in KR (free once the column stage is done)
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        mbar_expect_tx(kfull, Cfg::G * Cfg::KR_PLANE * 8);
        bulk_g2s(KR, P.k[prob] + koff, Cfg::G * Cfg::KR_PLANE * 8, kfull);
    };
    if (tid == 0 && (int)blockIdx.x < n_groups) {
        issue(blockIdx.x);
        if (Cfg::KSPEC) issue_k(blockIdx.x);
    }

    int it = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        const FftBufs bufs{raw + (int)(xoff & 3), raw + Cfg::XWIN + (Cfg::KSPEC ? 0 : (int)(koff & 3)), XR, KR, CT, so};
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
                if (Cfg::KSPEC) mbar_wait(kfull, it & 1);
#pragma unroll 1
                for (int t = tid; t < Cfg::COL_TASKS; t += Cfg::NT) fftc_col<Cfg>(bufs, t);
                __syncthreads();
                if (Cfg::KSPEC && tid == 0 && g + (int)gridDim.x < n_groups) issue_k(g + gridDim.x);  // KR is free again
            }
        }
        float *dst = P.out[prob] + ooff;
#pragma unroll 2
        for (int e = tid; e < Cfg::OUT_FLOATS; e += Cfg::NT) dst[e] = so[e];
        __syncthreads();  // the tile aliases XR, which the next group's phase R writes
    }
}

// ---- software-pipelined variant ----------------------------------------------------------------------------------------------
// The three-phase kernel above leaves warps at barriers: phase O has half as many FFT tasks as phase R (one complex inverse FFT
// serves two planes), so half of the CTA idles through it, and every phase ends with the skew of its slowest warp (ncu: 1.0-1.2
// barrier-stall cycles per issued instruction, the largest stall; 12 resident warps per SM).  Here the CTA has exactly as many warps
// as phase R and phase O have FFT tasks TOGETHER (6 + 3 at 61x61, 4 + 2 at 29x29 circular and 39x39) and a group takes two rounds:
//
//     column round   COL(g)                      all warps (an odd warp out copies the finished tile of g-1 to global memory)
//     FFT round      O(g)  ||  R(g+1)            inverse row FFTs of this group next to the row FFTs of the NEXT one
//
// so every warp runs one FFT task and one column task per group, there are two barriers per group instead of four, and 18
// instead of 12 warps are resident per SM at 61x61.  Costs: the output tile no longer aliases the row spectra (R(g+1) writes them
// while O(g) writes the tile; +8.7 KB) and the column stage is cut into 4 instead of 3 row segments per plane.
template <class Cfg>
struct PipeCfg {
    static constexpr int RW = Cfg::R_TASKS / 32, OW = Cfg::O_TASKS / 32, NW = Cfg::NT / 32;
    static_assert(RW + OW == NW, "a warp per FFT task of R and O together");
    static_assert(Cfg::COL_TASKS <= Cfg::NT, "one column task per thread");
    static constexpr int SPARE = Cfg::NT - Cfg::COL_TASKS;  // threads without a column task: they copy the previous tile out
    static constexpr unsigned long long SMEM = Cfg::SMEM + (unsigned long long)Cfg::OUT_FLOATS * 4;
    static constexpr int CTAS = SMEM <= 75 * 1024 ? 3 : (SMEM <= 113 * 1024 ? 2 : 1);
    static_assert(Cfg::OUT_FLOATS % 2 == 0, "tiles are copied as float2");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, PipeCfg<Cfg>::CTAS)
    xcorr_fft_pipe_kernel(XProblems P, int groups_per_problem, int n_groups, int C, long long k_bstride) {
    using PC = PipeCfg<Cfg>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *raw = reinterpret_cast<float *>(smem_raw);
    float2 *XR = reinterpret_cast<float2 *>(raw + Cfg::RAW_FLOATS);
    float2 *KR = XR + Cfg::G * Cfg::XR_PLANE;
    float2 *CT = KR + Cfg::G * Cfg::KR_PLANE;
    float *so = reinterpret_cast<float *>(CT + Cfg::G * Cfg::CT_PLANE);  // its own buffer: O(g) fills it while R(g+1) fills XR
    uint64_t *full = reinterpret_cast<uint64_t *>(so + Cfg::OUT_FLOATS), *kfull = full + 1;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_init(kfull, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto locate = [&](int g, int &prob, long long &xoff, long long &koff, long long &ooff) {
        prob = g / groups_per_problem;
        const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
        const long long b = plane0 / C, c0 = plane0 - b * C;
        xoff = plane0 * Cfg::XPL;
        koff = Cfg::KSPEC ? c0 * (2 * Cfg::KR_PLANE) : b * k_bstride + c0 * Cfg::KPL;
        ooff = plane0 * Cfg::OPL;
    };
    auto issue = [&](int g) {  // elected thread only
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        mbar_expect_tx(full, Cfg::RAW_FLOATS * 4);
        bulk_g2s(raw, P.x[prob] + (xoff & ~3ll), Cfg::XWIN * 4, full);
        if (!Cfg::KSPEC) bulk_g2s(raw + Cfg::XWIN, P.k[prob] + (koff & ~3ll), Cfg::KWIN * 4, full);
    };
    auto issue_k = [&](int g) {  // KSPEC, elected thread only: ready-made template spectra straight into KR
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        mbar_expect_tx(kfull, Cfg::G * Cfg::KR_PLANE * 8);
        bulk_g2s(KR, P.k[prob] + koff, Cfg::G * Cfg::KR_PLANE * 8, kfull);
    };
    // one FFT task of the round: phase R of group gr (warps < RW; gr < 0: none) or phase O of the group whose spectra are in CT
    auto fft_round = [&](int gr, int it_r, bool do_o) {
        const bool is_r = warp < PC::RW;
        const int ph = is_r ? FFT_PH_R : FFT_PH_O;
        if (is_r ? gr < 0 : !do_o) return;
        int prob = 0;
        long long xoff = 0, koff = 0, ooff = 0;
        if (is_r) {
            locate(gr, prob, xoff, koff, ooff);
            mbar_wait(full, it_r & 1);
        }
        const FftBufs bufs{raw + (int)(xoff & 3), raw + Cfg::XWIN + (Cfg::KSPEC ? 0 : (int)(koff & 3)), XR, KR, CT, so};
        const int t = is_r ? tid : tid - PC::RW * 32;
        const int h = fft_task_half<Cfg>(ph, t), unit = fft_task_unit<Cfg>(ph, t);
        float re[32], im[32];
        if (fftc_load<Cfg>(ph, bufs, unit, h, re, im)) {
            if (h) fft::half_twiddle(re, im);
            fft::fft32_fwd(re, im);
            fftc_store<Cfg>(ph, bufs, unit, h, re, im);
        }
    };
    auto copy_out = [&](float *dst, int first, int step) {
        float2 *d2 = reinterpret_cast<float2 *>(dst);  // a group starts at an even plane: 8-byte aligned
        const float2 *s2 = reinterpret_cast<const float2 *>(so);
#pragma unroll 2
        for (int e = first; e < Cfg::OUT_FLOATS / 2; e += step) d2[e] = s2[e];
    };

    const int g0 = blockIdx.x;
    if (g0 >= n_groups) return;
    if (tid == 0) {
        issue(g0);
        if (Cfg::KSPEC) issue_k(g0);
    }
    fft_round(g0, 0, false);  // prologue: R(g0) alone.  (A single loop with the prologue as iteration -1 measured 5 % slower.)
    __syncthreads();
    if (tid == 0 && g0 + (int)gridDim.x < n_groups) issue(g0 + gridDim.x);

    float *prev_dst = nullptr;
    int it = 0;
#pragma unroll 1
    for (int g = g0; g < n_groups; g += gridDim.x, ++it) {
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        const FftBufs bufs{raw, raw, XR, KR, CT, so};
        // ---- column round (+ the previous tile on its way out)
        if (PC::SPARE > 0) {
            if (tid < Cfg::COL_TASKS) {
                if (Cfg::KSPEC) mbar_wait(kfull, it & 1);
                fftc_col<Cfg>(bufs, tid);
            } else if (prev_dst) copy_out(prev_dst, tid - Cfg::COL_TASKS, PC::SPARE > 0 ? PC::SPARE : 1);
        } else {
            if (prev_dst) copy_out(prev_dst, tid, Cfg::NT);
            if (Cfg::KSPEC) mbar_wait(kfull, it & 1);
            fftc_col<Cfg>(bufs, tid);
        }
        __syncthreads();
        // ---- FFT round: O(g) next to R(g + grid)
        const int gn = g + gridDim.x;
        if (Cfg::KSPEC && tid == 0 && gn < n_groups) issue_k(gn);  // KR is free again
        fft_round(gn < n_groups ? gn : -1, it + 1, true);
        __syncthreads();
        if (tid == 0 && gn + (int)gridDim.x < n_groups) issue(gn + gridDim.x);  // R(gn) has consumed the landing buffer
        prev_dst = P.out[prob] + ooff;
    }
    copy_out(prev_dst, tid, Cfg::NT);
}

template <class Cfg>
static int launch_fft_pipe(const XProblems &P, int n, int B, int C, long long kbs, cudaStream_t st) {
    using PC = PipeCfg<Cfg>;
    static DeviceOnce once;
    if (int e = once.run([] { return cudaFuncSetAttribute(xcorr_fft_pipe_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC::SMEM); }))
        return e;
    const int gpp = (int)(((long long)B * C) / Cfg::G);
    const int total = gpp * n;
    const int slots = sm_count() * PC::CTAS;
    const int grid = total < slots ? total : slots;
    xcorr_fft_pipe_kernel<Cfg><<<grid, Cfg::NT, PC::SMEM, st>>>(P, gpp, total, C, kbs);
    count_launch();
    return launch_status();
}

// ---- warp-specialised variant -----------------------------------------------------------------------------------------------
// One persistent CTA per SM whose warps keep ONE role for the whole launch and hand groups to each other through double-buffered
// shared memory and mbarriers -- no CTA-wide barrier anywhere:
//
//     R warps (one per 32 row-FFT tasks)    RAW[s] -> XR[s], KR[s]      wait raw_full[s], xr_free[s];  arrive xr_full[s], raw_free[s]
//     COL warps (one per plane x segment)   XR[s], KR[s] -> CT[s]       wait xr_full[s],  ct_free[s];  arrive ct_full[s], xr_free[s]
//     O warps (one per 32 inverse tasks)    CT[s] -> tile[s]            wait ct_full[s],  so_free[s];  arrive so_full[s], ct_free[s]
//     one copy warp                         tile[s] -> global           wait so_full[s];               arrive so_free[s]
//
// Every role has about the same number of instructions per group and warp (one half-FFT task or one column segment), so in
// steady state all 18 warps (61x61) issue all the time and a warp only waits when it is really starved.  The landing buffer of
// group i+2 is requested by the first R thread as soon as all R warps have released stage s.
template <class Cfg>
struct WsCfg {
    static constexpr int RW = Cfg::R_TASKS / 32, CW = Cfg::COL_TASKS / 32, OW = Cfg::O_TASKS / 32, NW = RW + CW + OW + 1, NT = NW * 32;
    static constexpr int STAGE_FLOATS = Cfg::RAW_FLOATS + 2 * Cfg::G * (Cfg::XR_PLANE + Cfg::KR_PLANE + Cfg::CT_PLANE) + Cfg::OUT_FLOATS;
    static_assert(Cfg::RAW_FLOATS % 4 == 0 && STAGE_FLOATS % 2 == 0 && Cfg::OUT_FLOATS % 2 == 0, "alignment of the stage's buffers");
    static constexpr int STAGE_PAD = (STAGE_FLOATS + 31) / 32 * 32;  // 128-byte multiple: the landing buffer of stage 1 stays 16-byte aligned
    static constexpr unsigned long long SMEM = 2ull * STAGE_PAD * 4 + 16 * 8;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    enum { RAW_FULL = 0, RAW_FREE = 2, XR_FULL = 4, XR_FREE = 6, CT_FULL = 8, CT_FREE = 10, SO_FULL = 12, SO_FREE = 14 };
};

template <class Cfg>
__global__ void __launch_bounds__(WsCfg<Cfg>::NT, 1)
    xcorr_fft_ws_kernel(XProblems P, int groups_per_problem, int n_groups, int C, long long k_bstride) {
    using W = WsCfg<Cfg>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *base = reinterpret_cast<float *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + 2 * W::STAGE_PAD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[W::RAW_FULL + s], 1);
            mbar_init(&bars[W::RAW_FREE + s], W::RW * 32);
            mbar_init(&bars[W::XR_FULL + s], W::RW * 32);
            mbar_init(&bars[W::XR_FREE + s], W::CW * 32);
            mbar_init(&bars[W::CT_FULL + s], W::CW * 32);
            mbar_init(&bars[W::CT_FREE + s], W::OW * 32);
            mbar_init(&bars[W::SO_FULL + s], W::OW * 32);
            mbar_init(&bars[W::SO_FREE + s], 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    auto locate = [&](int g, int &prob, long long &xoff, long long &koff, long long &ooff) {
        prob = g / groups_per_problem;
        const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
        const long long b = plane0 / C, c0 = plane0 - b * C;
        xoff = plane0 * Cfg::XPL;
        koff = b * k_bstride + c0 * Cfg::KPL;
        ooff = plane0 * Cfg::OPL;
    };
    auto stage = [&](int s, int xo, int ko) {
        float *raw = base + s * W::STAGE_PAD;
        float2 *XR = reinterpret_cast<float2 *>(raw + Cfg::RAW_FLOATS);
        float2 *KR = XR + Cfg::G * Cfg::XR_PLANE;
        float2 *CT = KR + Cfg::G * Cfg::KR_PLANE;
        float *so = reinterpret_cast<float *>(CT + Cfg::G * Cfg::CT_PLANE);
        return FftBufs{raw + xo, raw + Cfg::XWIN + ko, XR, KR, CT, so};
    };
    auto issue = [&](int g, int s) {  // one thread
        int prob;
        long long xoff, koff, ooff;
        locate(g, prob, xoff, koff, ooff);
        float *raw = base + s * W::STAGE_PAD;
        mbar_expect_tx(&bars[W::RAW_FULL + s], Cfg::RAW_FLOATS * 4);
        bulk_g2s(raw, P.x[prob] + (xoff & ~3ll), Cfg::XWIN * 4, &bars[W::RAW_FULL + s]);
        bulk_g2s(raw + Cfg::XWIN, P.k[prob] + (koff & ~3ll), Cfg::KWIN * 4, &bars[W::RAW_FULL + s]);
    };
    const int g0 = blockIdx.x, step = gridDim.x;
    if (g0 >= n_groups) return;

    const bool is_r = warp < W::RW, is_o = warp >= W::RW + W::CW && warp < W::RW + W::CW + W::OW;
    if (is_r || is_o) {  // --------------------------------------- row FFTs (R) and inverse row FFTs (O): ONE copy of the half-FFT code
        if (tid == 0) {
            issue(g0, 0);
            if (g0 + step < n_groups) issue(g0 + step, 1);
        }
        const int ph = is_r ? FFT_PH_R : FFT_PH_O;
        const int t = is_r ? tid : tid - (W::RW + W::CW) * 32;
        const int h = fft_task_half<Cfg>(ph, t), unit = fft_task_unit<Cfg>(ph, t);
        // R: wait xr_free / raw_full, arrive xr_full / raw_free;   O: wait so_free / ct_full, arrive so_full / ct_free
        uint64_t *wait_free = &bars[is_r ? W::XR_FREE : W::SO_FREE], *wait_full = &bars[is_r ? W::RAW_FULL : W::CT_FULL];
        uint64_t *done_full = &bars[is_r ? W::XR_FULL : W::SO_FULL], *done_free = &bars[is_r ? W::RAW_FREE : W::CT_FREE];
        int i = 0;
#pragma unroll 1
        for (int g = g0; g < n_groups; g += step, ++i) {
            const int s = i & 1, par = (i >> 1) & 1;
            if (tid == 0 && i >= 1 && g + step < n_groups) {  // group i+1 lands in stage s^1 once every R warp has released it (groups 0, 1: above)
                mbar_wait(&bars[W::RAW_FREE + (s ^ 1)], ((i - 1) >> 1) & 1);
                issue(g + step, s ^ 1);
            }
            int xo = 0, ko = 0;
            if (is_r) {
                int prob;
                long long xoff, koff, ooff;
                locate(g, prob, xoff, koff, ooff);
                xo = (int)(xoff & 3), ko = (int)(koff & 3);
            }
            const FftBufs bufs = stage(s, xo, ko);
            mbar_wait(wait_free + s, par ^ 1);
            mbar_wait(wait_full + s, par);
            float re[32], im[32];
            if (fftc_load<Cfg>(ph, bufs, unit, h, re, im)) {
                if (h) fft::half_twiddle(re, im);
                fft::fft32_fwd(re, im);
                fftc_store<Cfg>(ph, bufs, unit, h, re, im);
            }
            mbar_arrive(done_full + s);
            mbar_arrive(done_free + s);
        }
    } else if (warp < W::RW + W::CW) {  // ------------------------------------------------------------------ column stage
        const int t = tid - W::RW * 32;
        int i = 0;
#pragma unroll 1
        for (int g = g0; g < n_groups; g += step, ++i) {
            const int s = i & 1, par = (i >> 1) & 1;
            const FftBufs bufs = stage(s, 0, 0);
            mbar_wait(&bars[W::CT_FREE + s], par ^ 1);
            mbar_wait(&bars[W::XR_FULL + s], par);
            fftc_col<Cfg>(bufs, t);
            mbar_arrive(&bars[W::CT_FULL + s]);
            mbar_arrive(&bars[W::XR_FREE + s]);
        }
    } else {  // ------------------------------------------------------------------------------------------------ tiles out
        int i = 0;
        for (int g = g0; g < n_groups; g += step, ++i) {
            const int s = i & 1, par = (i >> 1) & 1;
            int prob;
            long long xoff, koff, ooff;
            locate(g, prob, xoff, koff, ooff);
            const FftBufs bufs = stage(s, 0, 0);
            float2 *d2 = reinterpret_cast<float2 *>(P.out[prob] + ooff);  // a group starts at an even plane: 8-byte aligned
            const float2 *s2 = reinterpret_cast<const float2 *>(bufs.out);
            mbar_wait(&bars[W::SO_FULL + s], par);
#pragma unroll 4
            for (int e = lane; e < Cfg::OUT_FLOATS / 2; e += 32) d2[e] = s2[e];
            mbar_arrive(&bars[W::SO_FREE + s]);
        }
    }
}

template <class Cfg>
static int launch_fft_ws(const XProblems &P, int n, int B, int C, long long kbs, cudaStream_t st) {
    using W = WsCfg<Cfg>;
    static DeviceOnce once;
    if (int e = once.run([] { return cudaFuncSetAttribute(xcorr_fft_ws_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W::SMEM); }))
        return e;
    const int gpp = (int)(((long long)B * C) / Cfg::G);
    const int total = gpp * n;
    const int grid = total < sm_count() ? total : sm_count();
    xcorr_fft_ws_kernel<Cfg><<<grid, W::NT, W::SMEM, st>>>(P, gpp, total, C, kbs);
    count_launch();
    return launch_status();
}

// ---- template row spectra, once per template ------------------------------------------------------------------------------------
// K'(u,f) of a template [C,KH,KW] in the layout the KSPEC kernels land in KR: [C][KH][33] complex.  It is phase R itself -- the units
// (x_j, k_j) with an all-zero x -- so the spectra are the values the per-pair kernels compute (Z = 0 + i*k_j separates exactly).
// A shared template (k_batch_stride == 0: config 3's broadcast template, the tracker's per-sequence template) pays for its 29 row
// transforms once instead of once per pair and frame: -31 % of phase R at 61x61, -50 % at 29x29 circular.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::CTAS) xcorr_spectra_kernel(XProblems P, int groups_per_problem, int n_groups) {
    static_assert(!Cfg::KSPEC, "the producer runs the plain phase R");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *raw = reinterpret_cast<float *>(smem_raw);
    float2 *XR = reinterpret_cast<float2 *>(raw + Cfg::RAW_FLOATS);
    float2 *KR = XR + Cfg::G * Cfg::XR_PLANE;
    float2 *CT = KR + Cfg::G * Cfg::KR_PLANE;
    uint64_t *full = reinterpret_cast<uint64_t *>(CT + Cfg::G * Cfg::CT_PLANE);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_fence_init();
    }
    for (int e = tid; e < Cfg::XWIN; e += Cfg::NT) raw[e] = 0.f;  // the x rows of every unit
    __syncthreads();
    int it = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
        const int prob = g / groups_per_problem;
        const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G, koff = plane0 * Cfg::KPL;
        if (tid == 0) {
            mbar_expect_tx(full, Cfg::KWIN * 4);
            bulk_g2s(raw + Cfg::XWIN, P.k[prob] + (koff & ~3ll), Cfg::KWIN * 4, full);
        }
        const FftBufs bufs{raw, raw + Cfg::XWIN + (int)(koff & 3), XR, KR, CT, reinterpret_cast<float *>(XR)};
        mbar_wait(full, it & 1);
#pragma unroll 1
        for (int t0 = 0; t0 < Cfg::R_TASKS; t0 += Cfg::NT) {
            const int t = t0 + tid;
            const int h = fft_task_half<Cfg>(FFT_PH_R, t), unit = fft_task_unit<Cfg>(FFT_PH_R, t);
            float re[32], im[32];
            if (t < Cfg::R_TASKS && fftc_load<Cfg>(FFT_PH_R, bufs, unit, h, re, im)) {
                if (h) fft::half_twiddle(re, im);
                fft::fft32_fwd(re, im);
                fftc_store<Cfg>(FFT_PH_R, bufs, unit, h, re, im);
            }
        }
        __syncthreads();
        float2 *dst = reinterpret_cast<float2 *>(P.out[prob]) + plane0 * Cfg::KR_PLANE;
        for (int e = tid; e < Cfg::G * Cfg::KR_PLANE; e += Cfg::NT) dst[e] = KR[e];
        __syncthreads();  // KR and the landing buffer are reused by the next group
    }
}

template <class Cfg>
static int launch_spectra(const XProblems &P, int n, int C, cudaStream_t st) {
    static DeviceOnce once;
    if (int e = once.run([] { return cudaFuncSetAttribute(xcorr_spectra_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM); }))
        return e;
    const int gpp = C / Cfg::G, total = gpp * n, slots = sm_count() * Cfg::CTAS;
    xcorr_spectra_kernel<Cfg><<<total < slots ? total : slots, Cfg::NT, Cfg::SMEM, st>>>(P, gpp, total);
    count_launch();
    return launch_status();
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

// shared template with ready-made row spectra (KSPEC): three-phase and pipelined CTAs
using S256 = FCfg<29, 29, 61, 61, false, 2, 192, HDN_FFT_TB, true>;
using S256P = FCfg<29, 29, 61, 61, false, 2, 224, HDN_FFT_TB, true>;    // 4 R + 3 O warps; 6 column warps + one warp copying tiles out
using S256Lp = FCfg<29, 29, 29, 29, true, 2, 128, HDN_FFT_TB, true>;
using S256LpP = FCfg<29, 29, 29, 29, true, 2, 128, HDN_FFT_TB, true>;   // 2 R + 2 O warps, 4 column warps: three CTAs per SM
// the pipelined kernel's CTAs: one warp per FFT task of R and O together
using P256 = FCfg<29, 29, 61, 61, false, 2, 288>;
using P256Lp = FCfg<29, 29, 29, 29, true, 2, 192>;
using PWin15 = FCfg<15, 15, 39, 39, false, 2, 192>;

// (three-phase config, pipelined / warp-specialised config, variant that measured fastest on a B200: 1 = three-phase, 2 = pipelined)
//   61x61 (*) 29x29:       1.035 / 1.081 / 1.297 ms (three-phase / pipelined / warp-specialised), 6 x 64 x 256 planes
//   29x29 circ (*) 29x29:  0.769 / 0.723 / 0.780 ms
//   39x39 (*) 15x15:       2.196 / 2.345 / 2.613 ms (6 x 256 x 256 planes)
#define HDN_FFT_SHAPES(X) X(F256, P256, 1) X(F256Lp, P256Lp, 2) X(FWin15, PWin15, 1)

bool xcorr_fft_applicable(int C, int Hx, int Wx, int Hk, int Wk, int circular) {
#define HDN_IS(CFG, PCFG, BEST) \
    if (Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % 4 == 0) return true;
    HDN_FFT_SHAPES(HDN_IS)
#undef HDN_IS
    return false;
}

// variant: 0 = the fastest measured kernel of the shape, 1 = three-phase, 2 = pipelined, 3 = warp-specialised
int xcorr_fft_dispatch(const XProblems &P, int n, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular, long long kbs, int variant,
                       cudaStream_t st) {
#define HDN_TRY(CFG, PCFG, BEST) \
    if (Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % 4 == 0) \
        switch (variant ? variant : BEST) { \
            case 1: return launch_fft<CFG>(P, n, B, C, kbs, st); \
            case 3: return launch_fft_ws<PCFG>(P, n, B, C, kbs, st); \
            default: return launch_fft_pipe<PCFG>(P, n, B, C, kbs, st); \
        }
    HDN_FFT_SHAPES(HDN_TRY)
#undef HDN_TRY
    return HDN_ERR_UNSUPPORTED;
}

// ---- shared template, cached spectra -----------------------------------------------------------------------------------------
// variant of the KSPEC kernels picked by variant == 0 (1 = three-phase, 2 = pipelined), by measurement on a B200 (B = 64, 6 problems):
//   61x61:          three-phase 0.961 ms, pipelined 1.004 ms   (per-pair templates: 1.032 ms)
//   29x29 circular: three-phase 0.750 ms, pipelined 0.692 ms   (per-pair templates: 0.723 ms)
// The gain is smaller than the -31 % / -50 % of phase R: a round of the pipelined kernel is one FFT task per warp however many warps
// phase R needs, and the three-phase kernel only sheds issue pressure (4 of 6 warps busy in phase R).
#ifndef HDN_FFT_SPEC_K1
#define HDN_FFT_SPEC_K1 1
#endif
#ifndef HDN_FFT_SPEC_K2
#define HDN_FFT_SPEC_K2 2
#endif

long long xcorr_spectra_floats(int C, int Hx, int Wx, int Hk, int Wk, int circular) {
    if (C % 4 != 0 || Hk != 29 || Wk != 29) return 0;
    if (!circular && Hx == 61 && Wx == 61) return (long long)C * F256::KR_PLANE * 2;
    if (circular && Hx == 29 && Wx == 29) return (long long)C * F256Lp::KR_PLANE * 2;
    return 0;
}

// P.k[i] = template i [C,Hk,Wk];  P.out[i] = its spectra
int xcorr_spectra_dispatch(const XProblems &P, int n, int C, int Hx, int Wx, int Hk, int Wk, int circular, cudaStream_t st) {
    if (!xcorr_spectra_floats(C, Hx, Wx, Hk, Wk, circular)) return HDN_ERR_UNSUPPORTED;
    return circular ? launch_spectra<F256Lp>(P, n, C, st) : launch_spectra<F256>(P, n, C, st);
}

// P.k[i] = spectra of problem i's (shared) template
int xcorr_fft_spec_dispatch(const XProblems &P, int n, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular, int variant, cudaStream_t st) {
    if (!xcorr_spectra_floats(C, Hx, Wx, Hk, Wk, circular)) return HDN_ERR_UNSUPPORTED;
    if (!circular) {
        if ((variant ? variant : HDN_FFT_SPEC_K1) == 1) return launch_fft<S256>(P, n, B, C, 0, st);
        return launch_fft_pipe<S256P>(P, n, B, C, 0, st);
    }
    if ((variant ? variant : HDN_FFT_SPEC_K2) == 1) return launch_fft<S256Lp>(P, n, B, C, 0, st);
    return launch_fft_pipe<S256LpP>(P, n, B, C, 0, st);
}

}  // namespace hdn
