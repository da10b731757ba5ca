// xcorr_fft.cuh -- per-thread phase code of the transform-domain depth-wise correlation (K1/K2 at the FMA-bound shapes).
//
//   out[i,j] = sum_{u,v} xp[i+u, j+v] * k[u,v]     xp = x (K1) or its circular-row / replicate-column padding (K2)
//
// Direct evaluation costs KH*KW FMAs per output (841 at 29x29: 81 flop/B, FMA-bound at 6-7x the HBM time).  Here the sum
// over v (along a row) is taken in the frequency domain and the sum over u (across rows) directly:
//
//   R    rows     X'(r,f) = FFT64_j xp[r,j],  K'(u,f) = FFT64_v k[u,v]        (the padded row fits 64 points: no wrap-around
//                 reaches a valid output).  Rows are REAL, so two of them share one complex FFT (z = a + i*b,
//                 2A(f) = Z(f) + conj Z(-f), 2B(f) = (Z(f) - conj Z(-f))/i): (x_r, k_r) for r < KH, then (x_r, x_{r+np}).
//                 Only f = 0..32 is kept (Hermitian); the real values at f = 0 and f = 32 share slot 0 -> 32 complex per row.
//   COL  columns  c~(i,f) = sum_u X'(i+u,f) * conj K'(u,f)       a 1-D complex correlation per frequency column: KH complex
//                 MACs (3 FFMAs each) per output instead of KH*KW real ones.  (Slot 0 holds two real columns: component-wise.)
//   O    outputs  out(i,:) = IFFT64_f c~(i,f), two planes per complex FFT:  Q(f) = c~_p(i,f) + i*c~_q(i,f), extended to f > 32
//                 by Hermitian symmetry;  out_p(i,:) + i*out_q(i,:) = IFFT(Q) = conj(FFT(conj Q)).
//
// Per output this is ~3*KH + 2 * (64-point FFT / 33) instead of KH*KW operations, fp32-accurate (3e-7 of max|out|).  A full
// 2-D FFT (columns transformed too) has the same operation count but needs five barrier-separated phases of twiddle-heavy
// code; the direct column stage is a dense FFMA loop like the direct kernel's.
//
// Every 64-point FFT is done by TWO threads (halves h = 0/1 own the even/odd outputs, fft64.cuh) entirely in registers, and
// phases R and O run the SAME forward half-FFT code (the phase only selects the load / store code around it), so the hot
// loops stay resident in the instruction cache.  Shared memory per plane: XR [rows][33], KR [KH][33], CT [HO][33] complex
// (33 = odd pitch: conflict-free for row-wise and column-wise access).  Scale: 2 * 2 * 64 = 2^8 (exact).
//
// Host-compilable: tests/native/host_fft_check.cpp runs these phases task by task on the CPU against a direct correlation.
#pragma once
#include "fft64.cuh"

#if !defined(__CUDACC__)
struct float2 {
    float x, y;
};
#endif

#ifndef HDN_FFT_TB
#define HDN_FFT_TB 8  // taps per register-window block of the column stage (build-time tunable)
#endif

namespace hdn {

// KSPEC_: the template's row spectra K'(u,f) arrive ready-made (computed once per template by the same phase-R code, see
// xcorr_spectra in xcorr_fft.cu) and land directly in KR -- phase R then transforms x rows only, in pairs (x_j, x_{j + R_PAIRS}).
template <int KH_, int KW_, int HX_, int WX_, bool CIRC_, int G_, int NT_, int TB_ = HDN_FFT_TB, bool KSPEC_ = false>
struct FCfg {
    static constexpr int KH = KH_, KW = KW_, HX = HX_, WX = WX_, G = G_, NT = NT_;
    static constexpr bool CIRC = CIRC_, KSPEC = KSPEC_;
    static constexpr int PH = CIRC ? HX / 2 : 0, PW = CIRC ? WX / 2 : 0;
    static constexpr int HP = HX + 2 * PH, WP = WX + 2 * PW;  // padded input extent
    static constexpr int HO = HP - KH + 1, WO = WP - KW + 1;
    static constexpr int XPL = HX * WX, KPL = KH * KW, OPL = HO * WO;
    static constexpr int PITCH = 33;
    // column stage: a warp = (plane, segment of SEG consecutive output rows), lane = frequency column
    static constexpr int NSEG = NT / (32 * G) > 0 ? NT / (32 * G) : 1, SEG = (HO + NSEG - 1) / NSEG;
    static constexpr int TB = TB_;  // taps per register-window block of the column stage
    static constexpr int XR_ROWS = NSEG * SEG + KH - 1 > HP ? NSEG * SEG + KH - 1 : HP;  // rows past HP are read (never used) by the last segment
    static constexpr int XR_PLANE = XR_ROWS * PITCH, KR_PLANE = KH * PITCH, CT_PLANE = HO * PITCH;  // complex elements
    // A group's x / k planes are fetched as 16-byte-aligned windows (TMA bulk copies need 16-byte addresses and sizes): a group of 2
    // planes of odd size starts 0 or 2 floats past a 16-byte boundary, so the window is the group rounded out to multiples of 4 floats.
    // (offset 0: the window ends 2 floats into the next group -- never the last one, whose offset is 2 because B*C % 4 == 0.)
    static constexpr int XWIN = (G * XPL + (G % 4 && XPL % 2 ? 2 : 0) + 3) / 4 * 4, KWIN = KSPEC ? 0 : (G * KPL + (G % 4 && KPL % 2 ? 2 : 0) + 3) / 4 * 4;
    static_assert(!KSPEC || (G * KR_PLANE * 8) % 16 == 0, "a group's spectra are one 16-byte-multiple bulk copy");
    static constexpr int RAW_FLOATS = XWIN + KWIN, OUT_FLOATS = G * OPL;
    // Rows that phase R transforms.  K2's padded plane repeats itself: padded row r is source row (r - PH) mod HX with the SAME
    // replicate-padded columns, so only the HX distinct source rows are transformed and each spectrum is stored at every padded
    // row it stands for (57 padded rows -> 29 transforms at 29x29: phase R drops from two half-empty passes to one full pass).
    static constexpr bool DEDUP = CIRC && HX >= KH;
    static constexpr int XSRC = DEDUP ? HX : HP;
    static constexpr int R_PAIRS = KSPEC ? (XSRC + 1) / 2 : (XSRC - KH + 1) / 2;  // the x rows without a kernel row are packed in pairs (r, r + R_PAIRS):
    static constexpr int R_KROWS = KSPEC ? 0 : KH;                                // units (x_j, k_j), j < R_KROWS, come first
    static constexpr int R_UNITS_PLANE = R_KROWS + R_PAIRS;   // consecutive units read consecutive rows (odd pitch: no bank conflicts)
    static constexpr int R_UNITS = G * R_UNITS_PLANE, O_UNITS = (G / 2) * HO;
    static constexpr int pad32(int n) { return (n + 31) / 32 * 32; }
    // FFT tasks = (unit, half); warps alternate halves so a warp runs one half only:  task t -> h = (t/32) % 2, unit = (t/64)*32 + t%32
    // Phase O with 33 units would fill 2 + 2 warps, the last two with one lane each: its tasks are numbered densely instead
    // (h = t / units), one warp then mixes both halves (it runs the h = 1 twiddles predicated) and a warp is saved.
    static constexpr bool O_DENSE = pad32(2 * O_UNITS) < 2 * pad32(O_UNITS);
    static constexpr int R_TASKS = 2 * pad32(R_UNITS), O_TASKS = O_DENSE ? pad32(2 * O_UNITS) : 2 * pad32(O_UNITS), COL_TASKS = G * NSEG * 32;
    // the output tile is staged over XR (dead once the column stage is done)
    static_assert((unsigned long long)G * XR_PLANE * 8 >= (unsigned long long)OUT_FLOATS * 4, "output tile must fit the XR region it aliases");
    static constexpr unsigned long long SMEM = (unsigned long long)RAW_FLOATS * 4 + (unsigned long long)G * (XR_PLANE + KR_PLANE + CT_PLANE) * 8 + 16;
    static constexpr int CTAS = SMEM <= 75 * 1024 ? 3 : (SMEM <= 113 * 1024 ? 2 : 1);  // resident CTAs per SM (228 KB of shared memory)
    static_assert(HP <= 64 && WP <= 64, "padded input must fit the 64-point transform");
    static_assert(KH <= HP && KW <= WP && HO <= 64 && WO <= 64, "shape");
    static_assert(G % 2 == 0, "planes are paired in phase O");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct FftBufs {
    const float *rawx, *rawk;  // landed planes of the group (dense)
    float2 *XR, *KR, *CT;
    float *out;  // output tile of the group (dense); aliases XR
};

enum { FFT_PH_R = 0, FFT_PH_O = 1, FFT_PHASES = 2 };

template <class Cfg>
HDN_HD int fft_task_half(int ph, int t) {
    return (ph == FFT_PH_O && Cfg::O_DENSE) ? (t >= Cfg::O_UNITS ? 1 : 0) : (t >> 5) & 1;
}
template <class Cfg>
HDN_HD int fft_task_unit(int ph, int t) {  // may be >= the phase's unit count: fftc_load() rejects it
    return (ph == FFT_PH_O && Cfg::O_DENSE) ? (t >= Cfg::O_UNITS ? t - Cfg::O_UNITS : t) : ((t >> 6) << 5) | (t & 31);
}

// source row of padded row r (circular rows for K2)
template <class Cfg>
HDN_HD int fft_src_row(int r) {
    int sr = r - Cfg::PH;
    if (Cfg::CIRC) {
        if (sr < 0) sr += Cfg::HX;
        else if (sr >= Cfg::HX) sr -= Cfg::HX;
    }
    return sr;
}

// source row of transformed row j: j itself when the padded plane's repeated rows are de-duplicated (or there is no padding)
template <class Cfg>
HDN_HD int fft_x_row(int j) {
    return Cfg::DEDUP ? j : fft_src_row<Cfg>(j);
}

// ---- loads: the 64 complex inputs a[n] of the unit's transform, folded on the fly into the half's 32 values ---------------------
//      s[n] = a[n] + sgn * a[n+32]      sgn = +1 (h = 0) / -1 (h = 1), one FFMA per value; 64 live registers instead of 128
template <class Cfg>
HDN_HD void fftc_fold_row(const float *row, bool valid, float sgn, float (&s)[32]) {  // a[n] = padded row, zero past WP
#pragma unroll
    for (int n = 0; n < 32; ++n) {
        int q = n - Cfg::PW, q2 = n + 32 - Cfg::PW;  // compile-time: replicate padding of the columns
        q = q < 0 ? 0 : (q > Cfg::WX - 1 ? Cfg::WX - 1 : q);
        q2 = q2 < 0 ? 0 : (q2 > Cfg::WX - 1 ? Cfg::WX - 1 : q2);
        const float lo = (n < Cfg::WP && valid) ? row[q] : 0.f;
        s[n] = (n + 32 < Cfg::WP && valid) ? row[q2] * sgn + lo : lo;
    }
}

template <class Cfg>
HDN_HD bool fftc_load(int ph, const FftBufs &b, int unit, int h, float (&re)[32], float (&im)[32]) {
    const float sgn = h ? -1.f : 1.f;
    if (ph == FFT_PH_R) {
        if (unit >= Cfg::R_UNITS) return false;
        const int p = unit / Cfg::R_UNITS_PLANE, j = unit - p * Cfg::R_UNITS_PLANE;
        fftc_fold_row<Cfg>(b.rawx + p * Cfg::XPL + fft_x_row<Cfg>(j) * Cfg::WX, true, sgn, re);
        if (j < Cfg::R_KROWS) {  // (x_j, k_j)
            const float *krow = b.rawk + p * Cfg::KPL + j * Cfg::KW;
#pragma unroll
            for (int n = 0; n < 32; ++n) {
                const float lo = n < Cfg::KW ? krow[n] : 0.f;
                im[n] = n + 32 < Cfg::KW ? krow[n + 32] * sgn + lo : lo;
            }
        } else {  // (x_j, x_{j + R_PAIRS})
            const int r2 = j + Cfg::R_PAIRS;
            const bool has2 = r2 < Cfg::XSRC;
            fftc_fold_row<Cfg>(b.rawx + p * Cfg::XPL + fft_x_row<Cfg>(has2 ? r2 : j) * Cfg::WX, has2, sgn, im);
        }
        return true;
    }
    if (unit >= Cfg::O_UNITS) return false;
    const int m = unit / Cfg::HO, i = unit - m * Cfg::HO;
    const float2 *rp = b.CT + (2 * m) * Cfg::CT_PLANE + i * Cfg::PITCH, *rq = rp + Cfg::CT_PLANE;
    // Stored rows are conj(c~) (slot 0 = (c~(0), -c~(32)), both real).  The transform's input is a = conj(Q), Q(f) = c~_p(f) + i*c~_q(f):
    //   a[f]    = (A.x + C.y,  A.y - C.x)      f = 1..31, A = rp[f], C = rq[f]
    //   a[64-f] = (A.x - C.y, -A.y - C.x)      (Hermitian extension)
    //   a[0]    = (A0.x, -C0.x),  a[32] = (-A0.y, C0.y)
    // s[n] = a[n] + sgn * a[n+32] pairs slot n with slot 32 - n, so n and 32 - n are produced from the same four loads.
    {
        const float2 A = rp[0], C = rq[0];
        re[0] = A.x - sgn * A.y; im[0] = sgn * C.y - C.x;
        const float2 A16 = rp[16], C16 = rq[16];
        re[16] = (A16.x + C16.y) + sgn * (A16.x - C16.y);
        im[16] = (A16.y - C16.x) - sgn * (A16.y + C16.x);
    }
#pragma unroll
    for (int n = 1; n < 16; ++n) {
        const float2 A = rp[n], C = rq[n], B = rp[32 - n], D = rq[32 - n];
        const float ar = A.x + C.y, ai = A.y - C.x, amr = A.x - C.y, ami = -A.y - C.x;  // a[n], a[64 - n]
        const float br = B.x + D.y, bi = B.y - D.x, bmr = B.x - D.y, bmi = -B.y - D.x;  // a[32 - n], a[32 + n]
        re[n] = bmr * sgn + ar; im[n] = bmi * sgn + ai;            // a[n] + sgn * a[n + 32]
        re[32 - n] = amr * sgn + br; im[32 - n] = ami * sgn + bi;  // a[32 - n] + sgn * a[64 - n]
    }
    return true;
}

// ---- stores: the half's 32 outputs X[2m + h] are at (re, im)[POS32(m)] ------------------------------------------------------
// Phase R needs the half as a compile-time constant (the partner X[64 - f] of X[f] sits at a register index that depends on it).
template <class Cfg, int H>
HDN_HD void fftc_store_R(const FftBufs &b, int unit, const float (&re)[32], const float (&im)[32]) {
    using namespace fft;
    const int p = unit / Cfg::R_UNITS_PLANE, j = unit - p * Cfg::R_UNITS_PLANE;
    const bool ktype = j < Cfg::R_KROWS;
    float2 *xr = b.XR + p * Cfg::XR_PLANE;
    // first padded row of x row j (DEDUP: source row j stands for padded rows j + PH - HX, j + PH, j + PH + HX inside [0, HP))
    const int j2 = j + Cfg::R_PAIRS;
    const int r1 = Cfg::DEDUP ? (j + Cfg::PH >= Cfg::HX ? j + Cfg::PH - Cfg::HX : j + Cfg::PH) : j;
    const int r2 = Cfg::DEDUP ? (j2 + Cfg::PH >= Cfg::HX ? j2 + Cfg::PH - Cfg::HX : j2 + Cfg::PH) : j2;
    float2 *d1 = xr + r1 * Cfg::PITCH;
    float2 *d2 = ktype ? b.KR + p * Cfg::KR_PLANE + j * Cfg::PITCH : xr + r2 * Cfg::PITCH;
    const bool has2 = ktype || j2 < Cfg::XSRC;
    // the same spectrum again HX rows further down (rows past HP are never read by a valid output)
    const bool rep1 = Cfg::DEDUP && r1 + Cfg::HX < Cfg::HP, rep2 = Cfg::DEDUP && !ktype && has2 && r2 + Cfg::HX < Cfg::HP;
    constexpr int REP = Cfg::HX * Cfg::PITCH;
    if (H == 0) {
        const float2 v1 = float2{2.f * re[HPOS(0)], 2.f * re[HPOS(32)]}, v2 = float2{2.f * im[HPOS(0)], 2.f * im[HPOS(32)]};
        d1[0] = v1;
        if (rep1) d1[REP] = v1;
        if (has2) d2[0] = v2;
        if (rep2) d2[REP] = v2;
    }
#pragma unroll
    for (int f = 2 - H; f < 32; f += 2) {  // 2A(f) = Z(f) + conj Z(-f),  2B(f) = (Z(f) - conj Z(-f)) / i
        const float ar = re[HPOS(f)], ai = im[HPOS(f)], br = re[HPOS(64 - f)], bi = im[HPOS(64 - f)];
        const float2 v1 = float2{ar + br, ai - bi}, v2 = float2{ai + bi, br - ar};
        d1[f] = v1;
        if (rep1) d1[f + REP] = v1;
        if (has2) d2[f] = v2;
        if (rep2) d2[f + REP] = v2;
    }
}

template <class Cfg>
HDN_HD void fftc_store(int ph, const FftBufs &b, int unit, int h, const float (&re)[32], const float (&im)[32]) {
    using namespace fft;
    if (ph == FFT_PH_R) {
        if (h == 0) fftc_store_R<Cfg, 0>(b, unit, re, im);
        else fftc_store_R<Cfg, 1>(b, unit, re, im);
        return;
    }
    const int m = unit / Cfg::HO, i = unit - m * Cfg::HO;
    constexpr float SCALE = 1.0f / 256.0f;
    float *op = b.out + (2 * m) * Cfg::OPL + i * Cfg::WO + h, *oq = op + Cfg::OPL;
#pragma unroll
    for (int mm = 0; 2 * mm < Cfg::WO; ++mm) {
        if (2 * mm + 1 < Cfg::WO || h == 0) {
            op[2 * mm] = re[POS32(mm)] * SCALE;
            oq[2 * mm] = im[POS32(mm)] * -SCALE;
        }
    }
}

template <class Cfg>
HDN_HD int fftc_tasks(int ph) {
    return ph == FFT_PH_R ? Cfg::R_TASKS : Cfg::O_TASKS;
}

// ---- column stage ----------------------------------------------------------------------------------------------------------
// conj c~(i,f) = sum_u conj X'(i+u,f) * K'(u,f)  for the SEG output rows of the task's segment.  NTAP taps per block; the
// SEG + NTAP - 1 input rows of a block are held in registers (one shared load feeds ~NTAP complex MACs).  A complex MAC costs
// THREE FMAs (Gauss), issued as 1.5 packed FFMA2: with  a1 = sum xr*kr,  a2 = sum xi*ki,  a3 = sum (xr + xi)*(kr - ki)
//     Re = a1 + a2,    Im(x * conj k) = a3 - a1 + a2        (the stored conjugate has the imaginary part a1 - a2 - a3).
// Slot 0 (f == 0) packs two REAL columns (f = 0 in .x, f = 32 in .y) whose products must not mix: it simply keeps (a1, -a2).
// d = a * b + c on both halves: one packed FFMA2 (fma.rn.f32x2) on sm_100, i.e. one issue slot for two FMAs
HDN_HD float2 fftc_fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return float2{a.x * b.x + c.x, a.y * b.y + c.y};
#endif
}

// Accumulators: a12[i] = (a1, a2) of output row i  -- its update is ONE FFMA2 of the loaded pairs (xr, xi) * (kr, ki);
//               a3p[m] = a3 of rows (2m, 2m+1)     -- one FFMA2 of (xs[d], xs[d+1]) * (kd, kd), xs = xr + xi kept as overlapping pairs
template <class Cfg, int NTAP>
HDN_HD void fftc_col_block(const float2 *xc, const float2 *kc, int u0, float2 (&a12)[Cfg::SEG], float2 (&a3p)[(Cfg::SEG + 1) / 2]) {
    constexpr int NW = Cfg::SEG + NTAP - 1;
    float2 w[NW], ws2[NW];
#pragma unroll
    for (int d = 0; d < NW; ++d) {
        w[d] = xc[(u0 + d) * Cfg::PITCH];
        const float sum = w[d].x + w[d].y;
        ws2[d].x = sum;
        ws2[d].y = 0.f;
        if (d > 0) ws2[d - 1].y = sum;
    }
#pragma unroll
    for (int tt = 0; tt < NTAP; ++tt) {
        const float2 k = kc[(u0 + tt) * Cfg::PITCH];
        const float kd = k.x - k.y;
        const float2 kdd = float2{kd, kd};
#pragma unroll
        for (int i = 0; i < Cfg::SEG; ++i) a12[i] = fftc_fma2(w[i + tt], k, a12[i]);
#pragma unroll
        for (int m = 0; m < (Cfg::SEG + 1) / 2; ++m) a3p[m] = fftc_fma2(ws2[2 * m + tt], kdd, a3p[m]);  // odd SEG: the last .y is never used
    }
}

template <class Cfg>
HDN_HD void fftc_col(const FftBufs &b, int task) {
    const int f = task & 31, ws = task >> 5;
    const int p = ws / Cfg::NSEG, i0 = (ws - p * Cfg::NSEG) * Cfg::SEG;
    const float2 *xc = b.XR + p * Cfg::XR_PLANE + i0 * Cfg::PITCH + f;
    const float2 *kc = b.KR + p * Cfg::KR_PLANE + f;
    float2 a12[Cfg::SEG], a3p[(Cfg::SEG + 1) / 2];
#pragma unroll
    for (int i = 0; i < Cfg::SEG; ++i) a12[i] = float2{0.f, 0.f};
#pragma unroll
    for (int m = 0; m < (Cfg::SEG + 1) / 2; ++m) a3p[m] = float2{0.f, 0.f};
    constexpr int FULL = Cfg::KH / Cfg::TB * Cfg::TB, REM = Cfg::KH - FULL;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int u0 = 0; u0 < FULL; u0 += Cfg::TB) fftc_col_block<Cfg, Cfg::TB>(xc, kc, u0, a12, a3p);
    if (REM > 0) fftc_col_block<Cfg, (REM > 0 ? REM : 1)>(xc, kc, FULL, a12, a3p);
    float2 *ct = b.CT + p * Cfg::CT_PLANE + i0 * Cfg::PITCH + f;
#pragma unroll
    for (int i = 0; i < Cfg::SEG; ++i) {
        const float a1 = a12[i].x, a2 = a12[i].y, a3 = (i & 1) ? a3p[i / 2].y : a3p[i / 2].x;
        if (i0 + i < Cfg::HO) ct[i * Cfg::PITCH] = f == 0 ? float2{a1, -a2} : float2{a1 + a2, a1 - a2 - a3};
    }
}

}  // namespace hdn
