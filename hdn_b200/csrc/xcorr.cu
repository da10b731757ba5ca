// xcorr.cu -- K1/K2: depth-wise (optionally circular) cross-correlation for sm_100a.
//
// Replaces hdn/core/xcorr.py:37-46 (xcorr_depthwise) and :48-61 (xcorr_depthwise_circular).
//
// Design (DESIGN.md "K1/K2"):
//   * A (b,c) plane is tiny (29x29 .. 61x61 fp32) and contiguous in NCHW, so a group of G planes of
//     x, of k and of out are three contiguous byte ranges.  One persistent CTA per SM walks groups
//     g = blockIdx.x, += gridDim.x; an elected thread stages x|k of group g+STAGES with 1-D TMA
//     bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier, and drains the finished
//     output tile with a bulk store from shared memory.  HBM sees only full-line, fully coalesced
//     traffic, each byte exactly once.
//   * Compute: one thread owns one OUTPUT ROW: WO accumulators and the WX-wide input row live in
//     registers, kernel taps are warp-broadcast LDS.  Row pitch is odd for every shape the network
//     produces, so lanes (= consecutive rows) hit distinct banks without padding.
//   * The circular variant never materialises the padded tensor: rows wrap with one conditional
//     add, columns clamp at COMPILE time (the clamped column is just a different register).
//   * Accumulation order is fixed (u outer, v inner, split-K halves combined in order) -> results
//     are run-to-run deterministic.
//   * Shapes outside the table fall back to a plain one-thread-per-output kernel (still on device).
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace hdn {

// PPW_ > 0: a warp owns PPW_ whole planes (lane = plane-in-warp * HO + row, the remaining lanes idle) instead of the dense
// thread = row numbering.  A plane's rows are an odd pitch apart, so lanes inside one plane never share a bank, but a warp that
// straddles two planes in the dense numbering does (plane pitch = pitch^2: 4 of 7 straddling lanes collide at 29x29, doubling the
// wavefronts of every load) -- which made the HBM-bound 5x5 kernel shared-memory-bound (84 % pipe utilisation, 43 % conflicts).
// RP_: row-pair mode for the HBM-bound 5x5 shape (implies two planes per warp): a thread owns TWO consecutive output rows and
// keeps them as the halves of float2 registers, so one packed FFMA2 (fma.rn.f32x2) updates both -- half the FMA issue slots for the
// same shared loads (each thread needs the same rows as before).  Lane stride is then two rows (even), which still maps the 13
// row pairs of a plane to distinct even banks, and the neighbouring plane (pitch^2 = 9 mod 32) to the odd ones.
template <int KH_, int KW_, int HX_, int WX_, bool CIRC_, int G_, int NT_, int STAGES_, int KSPLIT_, bool SPILL_, int CTAS_ = 1, int PPW_ = 0,
          bool RP_ = false>
struct XCfg {
    static constexpr bool RP = RP_;
    static constexpr int KH = KH_, KW = KW_, HX = HX_, WX = WX_, G = G_, NT = NT_, STAGES = STAGES_, KSPLIT = KSPLIT_, CTAS = CTAS_, PPW = PPW_;
    static constexpr bool CIRC = CIRC_, SPILL = SPILL_;
    static constexpr int PH = CIRC ? HX / 2 : 0, PW = CIRC ? WX / 2 : 0;
    static constexpr int HO = HX + 2 * PH - KH + 1, WO = WX + 2 * PW - KW + 1;
    static constexpr int XPL = HX * WX, KPL = KH * KW, OPL = HO * WO;
    static constexpr int STAGE_FLOATS = G * (XPL + KPL);
    static constexpr int OUT_FLOATS = G * OPL;
    // STAGES == 1: single-buffered input and output tile -- half the shared memory per CTA, twice the resident CTAs, which overlap each
    // other's copies instead of a CTA overlapping its own
    static constexpr int OUT_BUFS = STAGES == 1 ? 1 : 2;
    static constexpr size_t SMEM = (size_t)(STAGES * STAGE_FLOATS + OUT_BUFS * OUT_FLOATS) * 4 + STAGES * 8 + 16;
    static_assert(G % 4 == 0, "bulk copies need 16-byte multiples");
    static_assert(!SPILL || (HO == 33 && NT == 32 * G * KSPLIT), "row-spill mapping is for 33-row outputs");
    static_assert(SMEM * CTAS <= 227 * 1024 - 1024 * CTAS, "shared memory budget");
    static_assert(256 % G == 0, "G must divide the network's 256 channels or the staged path is never taken");
    static_assert(PPW == 0 || (KSPLIT == 1 && !SPILL && PPW * ((HX + 2 * PH - KH + 1 + (RP ? 1 : 0)) / (RP ? 2 : 1)) <= 32 && NT * PPW == 32 * G),
                  "warp-per-plane mapping");
    static_assert(!RP || ((PPW == 2 || PPW == 0) && KSPLIT == 1 && !SPILL && NT >= (PPW == 2 ? 16 * G : G * ((HX + 2 * PH - KH + 2) / 2))),
                  "row-pair mode: two planes per warp or dense numbering");
};

// Accumulate kernel rows [u0,u1) of output row i into acc[WO].
template <class Cfg>
__device__ __forceinline__ void row_accumulate(const float *__restrict__ xp, const float *__restrict__ kp, int i, int u0, int u1,
                                               float (&acc)[Cfg::WO]) {
#pragma unroll 1
    for (int u = u0; u < u1; ++u) {
        int r = i + u - Cfg::PH;
        if (Cfg::CIRC) {
            if (r < 0) r += Cfg::HX;
            else if (r >= Cfg::HX) r -= Cfg::HX;
        }
        const float *xr = xp + r * Cfg::WX;
        float xv[Cfg::WX];
#pragma unroll
        for (int c = 0; c < Cfg::WX; ++c) xv[c] = xr[c];
        const float *kr = kp + u * Cfg::KW;
#pragma unroll
        for (int v = 0; v < Cfg::KW; ++v) {
            const float kv = kr[v];
#pragma unroll
            for (int c = 0; c < Cfg::WO; ++c) {
                int q = c + v - Cfg::PW;  // compile-time after unrolling
                q = q < 0 ? 0 : (q > Cfg::WX - 1 ? Cfg::WX - 1 : q);
                acc[c] = fmaf(xv[q], kv, acc[c]);
            }
        }
    }
}

template <class Cfg>
__device__ __forceinline__ void compute_group(const float *__restrict__ sx, const float *__restrict__ sk, float *__restrict__ so, int tid) {
    constexpr int ROWS = Cfg::G * Cfg::HO;
    if constexpr (Cfg::RP) {
        constexpr int NP = (Cfg::HO + 1) / 2;  // row pairs per plane
        // PPW == 2: a warp owns two whole planes (lane = plane-in-warp * NP + pair).
        // PPW == 0: thread = (pair, plane) with the PLANE in the lane: consecutive lanes are one plane pitch apart, and the pitch of every
        //           network shape is odd (169 = 9 mod 32), so the 32 lanes of a load hit 32 different banks whatever the row
        int p, j;
        bool active;
        if (Cfg::PPW == 2) {
            const int lane = tid & 31, pl = lane / NP;
            p = (tid >> 5) * 2 + pl; j = lane - pl * NP; active = pl < 2;
        } else {
            j = tid / Cfg::G; p = tid - j * Cfg::G; active = tid < Cfg::G * NP;
        }
        if (active) {
            const int i = 2 * j;
            const float *xp = sx + p * Cfg::XPL, *kp = sk + p * Cfg::KPL;
            float2 acc[Cfg::WO];
#pragma unroll
            for (int c = 0; c < Cfg::WO; ++c) acc[c] = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int u = 0; u < Cfg::KH; ++u) {
                // input rows of output rows i (-> .x) and i+1 (-> .y).  Plain: rows i+u, i+u+1 -- for the last, odd pair the second row lies
                // past the plane: it is read (the stage buffer continues with the next plane / the templates), its results are never
                // stored.  Circular: both rows wrap inside the plane.
                int r0 = i + u - Cfg::PH, r1 = r0 + 1;
                if (Cfg::CIRC) {
                    if (r0 < 0) r0 += Cfg::HX;
                    else if (r0 >= Cfg::HX) r0 -= Cfg::HX;
                    if (r1 < 0) r1 += Cfg::HX;
                    else if (r1 >= Cfg::HX) r1 -= Cfg::HX;
                }
                const float *x0 = xp + r0 * Cfg::WX, *x1 = xp + r1 * Cfg::WX;
                float2 xv[Cfg::WX];
#pragma unroll
                for (int c = 0; c < Cfg::WX; ++c) xv[c] = make_float2(x0[c], x1[c]);
                const float *kr = kp + u * Cfg::KW;
#pragma unroll
                for (int v = 0; v < Cfg::KW; ++v) {
                    const float kv = kr[v];
                    const float2 kk = make_float2(kv, kv);
#pragma unroll
                    for (int c = 0; c < Cfg::WO; ++c) {
                        int q = c + v - Cfg::PW;  // compile-time after unrolling: replicate padding of the columns
                        q = q < 0 ? 0 : (q > Cfg::WX - 1 ? Cfg::WX - 1 : q);
                        acc[c] = __ffma2_rn(xv[q], kk, acc[c]);
                    }
                }
            }
            float *o = so + p * Cfg::OPL + i * Cfg::WO;
#pragma unroll
            for (int c = 0; c < Cfg::WO; ++c) o[c] = acc[c].x;
            if (i + 1 < Cfg::HO) {
#pragma unroll
                for (int c = 0; c < Cfg::WO; ++c) o[Cfg::WO + c] = acc[c].y;
            }
        }
    } else if constexpr (Cfg::PPW > 0) {
        const int lane = tid & 31, pl = lane / Cfg::HO, i = lane - pl * Cfg::HO;
        if (pl < Cfg::PPW) {
            const int p = (tid >> 5) * Cfg::PPW + pl;
            float acc[Cfg::WO];
#pragma unroll
            for (int c = 0; c < Cfg::WO; ++c) acc[c] = 0.f;
            row_accumulate<Cfg>(sx + p * Cfg::XPL, sk + p * Cfg::KPL, i, 0, Cfg::KH, acc);
            float *o = so + p * Cfg::OPL + i * Cfg::WO;
#pragma unroll
            for (int c = 0; c < Cfg::WO; ++c) o[c] = acc[c];
        }
    } else if constexpr (!Cfg::SPILL) {
#pragma unroll 1
        for (int ks = 0; ks < Cfg::KSPLIT; ++ks) {
            // split-K slices are combined in slice order (deterministic)
#pragma unroll 1
            for (int t = tid; t < ROWS * Cfg::KSPLIT; t += Cfg::NT) {
                const int myks = t / ROWS;
                if (myks != ks) continue;
                const int rrow = t - myks * ROWS;
                const int p = rrow / Cfg::HO, i = rrow - p * Cfg::HO;
                float acc[Cfg::WO];
#pragma unroll
                for (int c = 0; c < Cfg::WO; ++c) acc[c] = 0.f;
                row_accumulate<Cfg>(sx + p * Cfg::XPL, sk + p * Cfg::KPL, i, ks * Cfg::KH / Cfg::KSPLIT, (ks + 1) * Cfg::KH / Cfg::KSPLIT, acc);
                float *o = so + p * Cfg::OPL + i * Cfg::WO;
                if (ks == 0) {
#pragma unroll
                    for (int c = 0; c < Cfg::WO; ++c) o[c] = acc[c];
                } else {
#pragma unroll
                    for (int c = 0; c < Cfg::WO; ++c) o[c] += acc[c];
                }
            }
            if (ks + 1 < Cfg::KSPLIT) __syncthreads();
        }
    } else {
        // 33-row outputs: warp = (plane, K-slice); lane = rows 0..31; row 32 is spread over the lanes
        // (lane l -> column l, column 32 via a warp reduction) so every lane carries the same load.
        const int warp = tid >> 5, lane = tid & 31;
        const int p = warp % Cfg::G, ks = warp / Cfg::G;
        const int u0 = ks * Cfg::KH / Cfg::KSPLIT, u1 = (ks + 1) * Cfg::KH / Cfg::KSPLIT;
        const float *xp = sx + p * Cfg::XPL, *kp = sk + p * Cfg::KPL;
        float acc[Cfg::WO];
#pragma unroll
        for (int c = 0; c < Cfg::WO; ++c) acc[c] = 0.f;
        row_accumulate<Cfg>(xp, kp, lane, u0, u1, acc);
        float e = 0.f, e32 = 0.f;  // out[32][lane], partial of out[32][32]
#pragma unroll 1
        for (int u = u0; u < u1; ++u) {
            const float *xr = xp + (32 + u) * Cfg::WX;
            const float *kr = kp + u * Cfg::KW;
#pragma unroll
            for (int v = 0; v < Cfg::KW; ++v) e = fmaf(xr[lane + v], kr[v], e);
            if (lane < Cfg::KW) e32 = fmaf(xr[32 + lane], kr[lane], e32);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) e32 += __shfl_xor_sync(0xffffffffu, e32, off);
        float *o = so + p * Cfg::OPL;
#pragma unroll 1
        for (int s = 0; s < Cfg::KSPLIT; ++s) {
            if (s == ks) {
                if (s == 0) {
#pragma unroll
                    for (int c = 0; c < Cfg::WO; ++c) o[lane * Cfg::WO + c] = acc[c];
                    o[32 * Cfg::WO + lane] = e;
                    if (lane == 0) o[32 * Cfg::WO + 32] = e32;
                } else {
#pragma unroll
                    for (int c = 0; c < Cfg::WO; ++c) o[lane * Cfg::WO + c] += acc[c];
                    o[32 * Cfg::WO + lane] += e;
                    if (lane == 0) o[32 * Cfg::WO + 32] += e32;
                }
            }
            if (s + 1 < Cfg::KSPLIT) __syncthreads();
        }
    }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::CTAS)
    xcorr_staged_kernel(XProblems P, int groups_per_problem, int n_groups, int C, long long k_bstride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sin = reinterpret_cast<float *>(smem_raw);
    float *sout = sin + Cfg::STAGES * Cfg::STAGE_FLOATS;
    uint64_t *full = reinterpret_cast<uint64_t *>(sout + Cfg::OUT_BUFS * Cfg::OUT_FLOATS);
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int stage, int g) {  // elected thread only
        const int prob = g / groups_per_problem;
        const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
        const long long b = plane0 / C, c0 = plane0 - b * C;
        float *dst = sin + stage * Cfg::STAGE_FLOATS;
        mbar_expect_tx(&full[stage], Cfg::STAGE_FLOATS * 4);
        bulk_g2s(dst, P.x[prob] + plane0 * Cfg::XPL, Cfg::G * Cfg::XPL * 4, &full[stage]);
        bulk_g2s(dst + Cfg::G * Cfg::XPL, P.k[prob] + b * k_bstride + c0 * Cfg::KPL, Cfg::G * Cfg::KPL * 4, &full[stage]);
    };

    if (tid == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            const int g = blockIdx.x + s * gridDim.x;
            if (g < n_groups) issue(s, g);
        }
    }

    int it = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
        const int s = it % Cfg::STAGES;
        if (Cfg::OUT_BUFS == 1) {  // the previous tile must have left the (only) output buffer before anyone writes it again
            if (tid == 0) bulk_wait_read<0>();
            __syncthreads();
        }
        mbar_wait(&full[s], (it / Cfg::STAGES) & 1);
        const float *sx = sin + s * Cfg::STAGE_FLOATS;
        float *so = sout + (Cfg::OUT_BUFS == 2 ? (it & 1) * Cfg::OUT_FLOATS : 0);
        compute_group<Cfg>(sx, sx + Cfg::G * Cfg::XPL, so, tid);
        if (Cfg::OUT_BUFS == 2 && tid == 0) bulk_wait_read<0>();  // store of iteration it-1 has left its buffer (reused at it+1)
        fence_proxy_async_smem();           // my so[] writes -> visible to the TMA store
        __syncthreads();                    // all rows written; all reads of stage s finished
        if (tid == 0) {
            const int prob = g / groups_per_problem;
            const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
            bulk_s2g(P.out[prob] + plane0 * Cfg::OPL, so, Cfg::OUT_FLOATS * 4);
            bulk_commit();
            const int gn = g + Cfg::STAGES * gridDim.x;
            if (gn < n_groups) issue(s, gn);
        }
    }
    if (tid == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------------
// Vectorised variant for the FMA-bound shapes (29x29 kernels).  Same TMA staging, but the landed planes are
// re-pitched once per group into a work tile whose row pitch XP (and kernel-row pitch KP) is a multiple of 4
// floats with XP/4 odd: every row starts 16-byte aligned, so a thread reads its input row and the kernel row
// with LDS.128, and lanes (= rows) of a quarter-warp still fall into distinct 16-byte bank groups.  Per kernel
// row u that is ceil(WX/4) + ceil(KW/4) shared loads feeding WO*KW FFMA (16 + 8 -> 957 for 61 (*) 29).
// The raw landing buffer is free as soon as the re-pitch is done, so the next group's TMA overlaps the compute.
template <int KH_, int KW_, int HX_, int WX_, bool CIRC_, int G_, int NT_, int KSPLIT_, int XP_, int KP_, bool TAIL_>
struct VCfg {
    static constexpr int KH = KH_, KW = KW_, HX = HX_, WX = WX_, G = G_, NT = NT_, KSPLIT = KSPLIT_, XP = XP_, KP = KP_, CTAS = 1;
    static constexpr bool CIRC = CIRC_, TAIL = TAIL_;
    static constexpr int PH = CIRC ? HX / 2 : 0, PW = CIRC ? WX / 2 : 0;
    static constexpr int HO = HX + 2 * PH - KH + 1, WO = WX + 2 * PW - KW + 1;
    static constexpr int XPL = HX * WX, KPL = KH * KW, OPL = HO * WO;
    static constexpr int NX4 = (WX + 3) / 4, NK4 = (KW + 3) / 4;
    static constexpr int RAW_FLOATS = G * (XPL + KPL), WX_FLOATS = G * HX * XP, WK_FLOATS = G * KH * KP, OUT_FLOATS = G * OPL;
    static constexpr size_t SMEM = (size_t)(RAW_FLOATS + WX_FLOATS + WK_FLOATS + 2 * OUT_FLOATS) * 4 + 8 + 16;
    static_assert(G % 4 == 0 && 256 % G == 0, "group size");
    static_assert(XP % 4 == 0 && (XP / 4) % 2 == 1 && XP >= NX4 * 4, "x pitch: 16-byte rows, odd number of 16-byte groups");
    static_assert(KP % 4 == 0 && (KP / 4) % 2 == 1 && KP >= NK4 * 4, "k pitch");
    static_assert(!TAIL || (HO == 33 && NT == 32 * G * KSPLIT && KH <= 32), "tail-row mapping is for 33-row outputs");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// One kernel row: acc[c] += sum_v x[r][c+v-PW] * k[u][v], operands fetched with 128-bit shared loads.
template <class Cfg>
__device__ __forceinline__ void row_step_vec(const float *__restrict__ xrow, const float *__restrict__ krow, float (&acc)[Cfg::WO]) {
    float xv[Cfg::NX4 * 4], kv[Cfg::NK4 * 4];
#pragma unroll
    for (int q = 0; q < Cfg::NX4; ++q) {
        const float4 t = reinterpret_cast<const float4 *>(xrow)[q];
        xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int q = 0; q < Cfg::NK4; ++q) {
        const float4 t = reinterpret_cast<const float4 *>(krow)[q];
        kv[4 * q] = t.x; kv[4 * q + 1] = t.y; kv[4 * q + 2] = t.z; kv[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int v = 0; v < Cfg::KW; ++v) {
#pragma unroll
        for (int c = 0; c < Cfg::WO; ++c) {
            int q = c + v - Cfg::PW;
            q = q < 0 ? 0 : (q > Cfg::WX - 1 ? Cfg::WX - 1 : q);
            acc[c] = fmaf(xv[q], kv[v], acc[c]);
        }
    }
}

template <class Cfg>
__device__ __forceinline__ void rows_vec(const float *__restrict__ xp, const float *__restrict__ kp, int i, int u0, int u1, float (&acc)[Cfg::WO]) {
#pragma unroll 1
    for (int u = u0; u < u1; ++u) {
        int r = i + u - Cfg::PH;
        if (Cfg::CIRC) {
            if (r < 0) r += Cfg::HX;
            else if (r >= Cfg::HX) r -= Cfg::HX;
        }
        row_step_vec<Cfg>(xp + r * Cfg::XP, kp + u * Cfg::KP, acc);
    }
}

template <class Cfg>
__device__ __forceinline__ void compute_group_vec(const float *__restrict__ wx, const float *__restrict__ wk, float *__restrict__ so, int tid) {
    constexpr int ROWS = Cfg::G * Cfg::HO;
    if constexpr (!Cfg::TAIL) {
#pragma unroll 1
        for (int ks = 0; ks < Cfg::KSPLIT; ++ks) {
#pragma unroll 1
            for (int t = tid; t < ROWS * Cfg::KSPLIT; t += Cfg::NT) {
                const int myks = t / ROWS;
                if (myks != ks) continue;
                const int rrow = t - myks * ROWS;
                const int p = rrow / Cfg::HO, i = rrow - p * Cfg::HO;
                float acc[Cfg::WO];
#pragma unroll
                for (int c = 0; c < Cfg::WO; ++c) acc[c] = 0.f;
                rows_vec<Cfg>(wx + p * Cfg::HX * Cfg::XP, wk + p * Cfg::KH * Cfg::KP, i, ks * Cfg::KH / Cfg::KSPLIT, (ks + 1) * Cfg::KH / Cfg::KSPLIT, acc);
                float *o = so + p * Cfg::OPL + i * Cfg::WO;
                if (ks == 0) {
#pragma unroll
                    for (int c = 0; c < Cfg::WO; ++c) o[c] = acc[c];
                } else {
#pragma unroll
                    for (int c = 0; c < Cfg::WO; ++c) o[c] += acc[c];
                }
            }
            if (ks + 1 < Cfg::KSPLIT) __syncthreads();
        }
    } else {
        // 33-row outputs.  warp = (plane, K-slice); lane = output rows 0..31 over the slice's kernel rows.
        // Row 32 is done by the slice-0 warp with lane = KERNEL row: every lane runs one row step (its own input
        // row 32+u against kernel row u), then a butterfly transpose-reduction leaves column j's total in lane j.
        // Slice 0 has the shorter u-range (KH/KSPLIT rounded down), so the extra step balances the slices.
        const int warp = tid >> 5, lane = tid & 31;
        const int p = warp % Cfg::G, ks = warp / Cfg::G;
        const int u0 = ks * Cfg::KH / Cfg::KSPLIT, u1 = (ks + 1) * Cfg::KH / Cfg::KSPLIT;
        const float *xp = wx + p * Cfg::HX * Cfg::XP, *kp = wk + p * Cfg::KH * Cfg::KP;
        float acc[Cfg::WO];
#pragma unroll
        for (int c = 0; c < Cfg::WO; ++c) acc[c] = 0.f;
        rows_vec<Cfg>(xp, kp, lane, u0, u1, acc);
        float *o = so + p * Cfg::OPL;
        if (ks == 0) {
#pragma unroll
            for (int c = 0; c < Cfg::WO; ++c) o[lane * Cfg::WO + c] = acc[c];
            float t[Cfg::WO];
#pragma unroll
            for (int c = 0; c < Cfg::WO; ++c) t[c] = 0.f;
            const int u = lane < Cfg::KH ? lane : Cfg::KH - 1;
            row_step_vec<Cfg>(xp + (32 + u) * Cfg::XP, kp + u * Cfg::KP, t);
            if (lane >= Cfg::KH) {
#pragma unroll
                for (int c = 0; c < Cfg::WO; ++c) t[c] = 0.f;
            }
            float last = t[32];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) last += __shfl_xor_sync(0xffffffffu, last, off);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {  // after the step with `off`, a lane holds `off` partial columns
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int j = 0; j < off; ++j) {
                    const float send = upper ? t[j] : t[j + off];
                    const float keep = upper ? t[j + off] : t[j];
                    t[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            o[32 * Cfg::WO + lane] = t[0];
            if (lane == 0) o[32 * Cfg::WO + 32] = last;
        }
#pragma unroll 1
        for (int s = 1; s < Cfg::KSPLIT; ++s) {
            __syncthreads();
            if (s == ks) {
#pragma unroll
                for (int c = 0; c < Cfg::WO; ++c) o[lane * Cfg::WO + c] += acc[c];
            }
        }
    }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 1)
    xcorr_vec_kernel(XProblems P, int groups_per_problem, int n_groups, int C, long long k_bstride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *raw = reinterpret_cast<float *>(smem_raw);
    float *wx = raw + Cfg::RAW_FLOATS;
    float *wk = wx + Cfg::WX_FLOATS;
    float *sout = wk + Cfg::WK_FLOATS;
    uint64_t *full = reinterpret_cast<uint64_t *>(sout + 2 * Cfg::OUT_FLOATS);
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
        // re-pitch: dense planes -> 16-byte-aligned rows (pad columns are never used as operands)
#pragma unroll 4
        for (int e = tid; e < Cfg::G * Cfg::XPL; e += Cfg::NT) {
            const int row = e / Cfg::WX, c = e - row * Cfg::WX;
            wx[row * Cfg::XP + c] = raw[e];
        }
        for (int e = tid; e < Cfg::G * Cfg::KPL; e += Cfg::NT) {
            const int row = e / Cfg::KW, c = e - row * Cfg::KW;
            wk[row * Cfg::KP + c] = raw[Cfg::G * Cfg::XPL + e];
        }
        __syncthreads();  // work tile complete; raw buffer free
        if (tid == 0) {
            const int gn = g + gridDim.x;
            if (gn < n_groups) issue(gn);  // lands while this group computes
        }
        float *so = sout + (it & 1) * Cfg::OUT_FLOATS;
        compute_group_vec<Cfg>(wx, wk, so, tid);
        if (tid == 0) bulk_wait_read<0>();
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const int prob = g / groups_per_problem;
            const long long plane0 = (long long)(g - prob * groups_per_problem) * Cfg::G;
            bulk_s2g(P.out[prob] + plane0 * Cfg::OPL, so, Cfg::OUT_FLOATS * 4);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all<0>();
}

// Any shape, any alignment: one thread per output element, straight from global memory.
__global__ void xcorr_generic_kernel(XProblems P, int nprob, int B, int C, int Hx, int Wx, int Hk, int Wk, int ph, int pw, int Ho, int Wo,
                                     long long k_bstride, long long plane_begin) {
    const long long planes = (long long)B * C - plane_begin;
    const long long per_prob = planes * Ho * Wo;
    const long long total = per_prob * nprob;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int prob = (int)(e / per_prob);
        long long r = e - prob * per_prob;
        const int j = (int)(r % Wo);
        r /= Wo;
        const int i = (int)(r % Ho);
        const long long plane = r / Ho + plane_begin;
        const long long b = plane / C, c = plane - b * C;
        const float *xp = P.x[prob] + plane * Hx * Wx;
        const float *kp = P.k[prob] + b * k_bstride + c * Hk * Wk;
        float acc = 0.f;
        for (int u = 0; u < Hk; ++u) {
            int rr = i + u - ph;
            if (ph) rr = ((rr % Hx) + Hx) % Hx;
            for (int v = 0; v < Wk; ++v) {
                int cc = j + v - pw;
                cc = cc < 0 ? 0 : (cc > Wx - 1 ? Wx - 1 : cc);
                acc = fmaf(__ldg(xp + rr * Wx + cc), __ldg(kp + u * Wk + v), acc);
            }
        }
        P.out[prob][(plane * Ho + i) * Wo + j] = acc;
    }
}

template <class Cfg, class = void>
struct KernelOf {
    static constexpr auto fn = xcorr_staged_kernel<Cfg>;
};
template <class Cfg>
struct KernelOf<Cfg, std::void_t<decltype(Cfg::XP)>> {
    static constexpr auto fn = xcorr_vec_kernel<Cfg>;
};

template <class Cfg>
static int launch_staged(const XProblems &P, int n, int B, int C, long long kbs, cudaStream_t st) {
    static DeviceOnce once;
    if (int e = once.run([] { return cudaFuncSetAttribute(KernelOf<Cfg>::fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM); }))
        return e;
    const int gpp = (int)(((long long)B * C) / Cfg::G);
    const int total = gpp * n;
    const int slots = sm_count() * Cfg::CTAS;
    const int grid = total < slots ? total : slots;
    KernelOf<Cfg>::fn<<<grid, Cfg::NT, Cfg::SMEM, st>>>(P, gpp, total, C, kbs);
    count_launch();
    return launch_status();
}

//                       KH  KW  HX  WX  circ   G   NT  ST KS spill
#ifndef HDN_NATIVE_DENSE  // A/B switch: the dense thread = row numbering this shape used before (0.963 ms vs 0.932 ms at batch 512)
#ifndef HDN_NATIVE_ROWPAIR_OFF
#ifndef HDN_NATIVE_STAGES
#define HDN_NATIVE_STAGES 1
#endif
// 127/255 crops (HBM-bound): a warp = two planes x 13 row pairs, FFMA2; single-buffered 4-warp CTAs, four per SM
using CfgNative = XCfg<5, 5, 29, 29, false, 8, 128, HDN_NATIVE_STAGES, 1, false, HDN_NATIVE_STAGES == 1 ? 4 : 2, 2, true>;
#else
using CfgNative = XCfg<5, 5, 29, 29, false, 8, 256, 2, 1, false, 2, 1>;  // a warp = one plane (25 rows), scalar FFMA
#endif
#else
using CfgNative = XCfg<5, 5, 29, 29, false, 8, 224, 2, 1, false, 2>;
#endif
// lp branch, 127 crops: FMA-bound (28 flop/B).  Row pairs + FFMA2; a warp = one row pair of 32 planes (conflict-free, all lanes busy)
#ifndef HDN_NATIVE_LP_SCALAR
using CfgNativeLp = XCfg<13, 13, 13, 13, true, 32, 224, 1, 1, false, 3, 0, true>;  // 32 planes x 7 row pairs; single-buffered, 3 CTAs/SM
#else
using CfgNativeLp = XCfg<13, 13, 13, 13, true, 32, 416, 3, 1, false>;
#endif
//                     KH  KW  HX  WX  circ   G   NT KS  XP  KP  tail
using Cfg256 = VCfg<29, 29, 61, 61, false, 4, 256, 2, 68, 36, true>;    // 256/512 crops (FMA-bound), LDS.128 operands
using Cfg256Lp = VCfg<29, 29, 29, 29, true, 8, 256, 1, 36, 36, false>;  // lp branch, INSTANCE_SIZE=512
using CfgWin15 = XCfg<15, 15, 39, 39, false, 8, 224, 2, 1, false>;     // 15x15 window sweep

static thread_local int g_xcorr_algo = HDN_XCORR_AUTO;  // hdn_xcorr_set_algo: per calling thread, so concurrent callers cannot flip each other's choice

// AUTO: the transform-domain kernel wherever one exists -- it measured faster than the direct sum on a B200 for all three shapes
// (61x61 (*) 29x29: 2.6x, 29x29 circular (*) 29x29: 2.3x, 39x39 (*) 15x15: 1.2x)
static bool fft_selected(int C, int Hx, int Wx, int Hk, int Wk, int circular) {
    return g_xcorr_algo != HDN_XCORR_DIRECT && xcorr_fft_applicable(C, Hx, Wx, Hk, Wk, circular);
}

static long long g_generic_launches = 0;  // hdn_xcorr_generic_launches

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static bool staged_applicable(int C, int Hx, int Wx, int Hk, int Wk, int circular, long long kbs) {
    if (!(kbs == 0 || kbs == (long long)C * Hk * Wk)) return false;
#define HDN_IS(CFG) \
    if (Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % CFG::G == 0) return true;
    HDN_IS(CfgNative) HDN_IS(CfgNativeLp) HDN_IS(Cfg256) HDN_IS(Cfg256Lp) HDN_IS(CfgWin15)
#undef HDN_IS
    return false;
}

static int xcorr_dispatch(const XProblems &P, int n, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular, long long kbs,
                          cudaStream_t st) {
    bool fast = true;
    for (int i = 0; i < n; ++i) fast = fast && aligned16(P.x[i]) && aligned16(P.k[i]) && aligned16(P.out[i]);
    fast = fast && (kbs == 0 || kbs == (long long)C * Hk * Wk);
    // FMA-bound shapes (29x29 / 15x15 templates): 64x64 FFT correlation, ~5x fewer instructions than the direct sum (xcorr_fft.cu)
    if (fast && fft_selected(C, Hx, Wx, Hk, Wk, circular))
        return xcorr_fft_dispatch(P, n, B, C, Hx, Wx, Hk, Wk, circular, kbs, g_xcorr_algo >= HDN_XCORR_FFT_PHASED ? g_xcorr_algo - HDN_XCORR_FFT_PHASED + 1 : 0, st);
#define HDN_TRY(CFG)                                                                                                               \
    if (fast && Hk == CFG::KH && Wk == CFG::KW && Hx == CFG::HX && Wx == CFG::WX && (circular != 0) == CFG::CIRC && C % CFG::G == 0) \
        return launch_staged<CFG>(P, n, B, C, kbs, st);
    HDN_TRY(CfgNative)
    HDN_TRY(CfgNativeLp)
    HDN_TRY(Cfg256)
    HDN_TRY(Cfg256Lp)
    HDN_TRY(CfgWin15)
#undef HDN_TRY
    // No tiled kernel for this shape / alignment: one thread per output, straight from global memory.  Correct but slow
    // (no staging, no register tiling) -- say so once per process so that a mis-sized crop does not pass silently.
    static int warned = 0;
    __atomic_fetch_add(&g_generic_launches, 1ll, __ATOMIC_RELAXED);
    if (!__atomic_exchange_n(&warned, 1, __ATOMIC_RELAXED) && !getenv("HDN_B200_QUIET"))
        fprintf(stderr, "hdn_b200: xcorr %dx%d (*) %dx%d%s, C=%d%s has no tiled kernel: running the generic one-thread-per-output kernel\n", Hx, Wx, Hk,
                Wk, circular ? " circular" : "", C, fast ? "" : " (unaligned pointers or strided template)");
    const int ph = circular ? Hx / 2 : 0, pw = circular ? Wx / 2 : 0;
    const int Ho = Hx + 2 * ph - Hk + 1, Wo = Wx + 2 * pw - Wk + 1;
    const long long total = (long long)n * B * C * Ho * Wo;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    xcorr_generic_kernel<<<(int)blocks, 256, 0, st>>>(P, n, B, C, Hx, Wx, Hk, Wk, ph, pw, Ho, Wo, kbs, 0);
    count_launch();
    return launch_status();
}

}  // namespace hdn

using namespace hdn;

extern "C" int hdn_xcorr_dw_multi_f32(int n, const float *const *x_host, const float *const *k_host, float *const *out_host, int B, int C,
                                      int Hx, int Wx, int Hk, int Wk, int circular, int64_t k_batch_stride, hdn_stream_t stream) {
    if (!x_host || !k_host || !out_host) return HDN_ERR_NULL;
    if (n < 1 || n > HDN_MAX_PROBLEMS) return HDN_ERR_UNSUPPORTED;
    if (B < 1 || C < 1 || Hx < 1 || Wx < 1 || Hk < 1 || Wk < 1) return HDN_ERR_SHAPE;
    const int ph = circular ? Hx / 2 : 0, pw = circular ? Wx / 2 : 0;
    if (Hk > Hx + 2 * ph || Wk > Wx + 2 * pw) return HDN_ERR_SHAPE;
    if (k_batch_stride != 0 && k_batch_stride < (int64_t)C * Hk * Wk) return HDN_ERR_SHAPE;
    XProblems P;
    for (int i = 0; i < n; ++i) {
        if (!x_host[i] || !k_host[i] || !out_host[i]) return HDN_ERR_NULL;
        if ((reinterpret_cast<uintptr_t>(x_host[i]) | reinterpret_cast<uintptr_t>(k_host[i]) | reinterpret_cast<uintptr_t>(out_host[i])) & 3u)
            return HDN_ERR_ALIGN;
        P.x[i] = x_host[i];
        P.k[i] = k_host[i];
        P.out[i] = out_host[i];
    }
    for (int i = n; i < HDN_MAX_PROBLEMS; ++i) P.x[i] = P.k[i] = P.out[i] = nullptr;
    return xcorr_dispatch(P, n, B, C, Hx, Wx, Hk, Wk, circular, (long long)k_batch_stride, (cudaStream_t)stream);
}

extern "C" int64_t hdn_xcorr_spectra_floats(int C, int Hx, int Wx, int Hk, int Wk, int circular) {
    return (int64_t)xcorr_spectra_floats(C, Hx, Wx, Hk, Wk, circular);
}

static int fill_problems(XProblems &P, int n, const float *const *a, const float *const *b, float *const *c) {
    if (!a || !b || !c) return HDN_ERR_NULL;
    if (n < 1 || n > HDN_MAX_PROBLEMS) return HDN_ERR_UNSUPPORTED;
    for (int i = 0; i < HDN_MAX_PROBLEMS; ++i) P.x[i] = P.k[i] = P.out[i] = nullptr;
    for (int i = 0; i < n; ++i) {
        if (!a[i] || !b[i] || !c[i]) return HDN_ERR_NULL;
        if ((reinterpret_cast<uintptr_t>(a[i]) | reinterpret_cast<uintptr_t>(b[i]) | reinterpret_cast<uintptr_t>(c[i])) & 15u) return HDN_ERR_ALIGN;
        P.x[i] = a[i];
        P.k[i] = b[i];
        P.out[i] = c[i];
    }
    return HDN_OK;
}

extern "C" int hdn_xcorr_template_spectra_f32(int n, const float *const *k_host, float *const *spectra_host, int C, int Hx, int Wx, int Hk,
                                              int Wk, int circular, hdn_stream_t stream) {
    XProblems P;
    if (int e = fill_problems(P, n, k_host, k_host, spectra_host)) return e;
    return xcorr_spectra_dispatch(P, n, C, Hx, Wx, Hk, Wk, circular, (cudaStream_t)stream);
}

extern "C" int hdn_xcorr_dw_multi_spec_f32(int n, const float *const *x_host, const float *const *spectra_host, float *const *out_host, int B,
                                           int C, int Hx, int Wx, int Hk, int Wk, int circular, hdn_stream_t stream) {
    if (B < 1) return HDN_ERR_SHAPE;
    XProblems P;
    if (int e = fill_problems(P, n, x_host, spectra_host, out_host)) return e;
    const int variant = g_xcorr_algo >= HDN_XCORR_FFT_PHASED ? g_xcorr_algo - HDN_XCORR_FFT_PHASED + 1 : 0;
    return xcorr_fft_spec_dispatch(P, n, B, C, Hx, Wx, Hk, Wk, circular, variant == 3 ? 0 : variant, (cudaStream_t)stream);
}

extern "C" int hdn_xcorr_dw_f32(const float *x, const float *k, float *out, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular,
                                int64_t k_batch_stride, hdn_stream_t stream) {
    return hdn_xcorr_dw_multi_f32(1, &x, &k, &out, B, C, Hx, Wx, Hk, Wk, circular, k_batch_stride, stream);
}

extern "C" int hdn_xcorr_set_algo(int algo) {
    if (algo < HDN_XCORR_AUTO || algo > HDN_XCORR_FFT_WS) return HDN_ERR_UNSUPPORTED;
    g_xcorr_algo = algo;
    return HDN_OK;
}

extern "C" int hdn_xcorr_uses_fft(int C, int Hx, int Wx, int Hk, int Wk, int circular, int64_t k_batch_stride) {
    // mirrors xcorr_dispatch: dense or shared template (and 16-byte aligned pointers, which the caller knows) + a transform-domain shape
    const bool kbs_ok = k_batch_stride == 0 || k_batch_stride == (int64_t)C * Hk * Wk;
    return (kbs_ok && fft_selected(C, Hx, Wx, Hk, Wk, circular)) ? 1 : 0;
}

extern "C" int64_t hdn_xcorr_generic_launches(void) { return __atomic_load_n(&g_generic_launches, __ATOMIC_RELAXED); }

extern "C" int hdn_xcorr_is_staged(int C, int Hx, int Wx, int Hk, int Wk, int circular, int64_t k_batch_stride) {
    return staged_applicable(C, Hx, Wx, Hk, Wk, circular, (long long)k_batch_stride) ? 1 : 0;
}
