// xcorr_fft.cuh -- per-thread phase code of the FFT depth-wise correlation (K1/K2 at the FMA-bound shapes).
//
//   out[i,j] = sum_{u,v} xp[i+u, j+v] * k[u,v]     xp = x (K1) or its circular-row / replicate-column padding (K2)
//
// is a 64x64 circular correlation whenever the padded input fits 64x64 (no wrap reaches a valid output):
//   OUT = IFFT2( FFT2(xp) * conj(FFT2(k)) ).
// Direct evaluation costs KH*KW FMAs per output (841 at 29x29: 81 flop/B, FMA-bound at 6-7x the HBM time); this route
// costs ~160 64-point FFTs per plane, ~5x fewer instructions.
//
// Every 64-point FFT is done by TWO threads (halves h = 0/1 own the even/odd outputs, fft64.cuh) entirely in registers;
// shared memory only carries the transposes.  All five phases run the SAME forward half-FFT code (inverse transforms use
// IFFT(a) = conj(FFT(conj a)), the conjugations are folded into the neighbouring phases), so the kernel is a loop
// phase -> { load (phase-specific) ; half-FFT (shared) ; store (phase-specific) } whose hot body stays in the instruction cache.
//
// Per plane (shared-memory buffers XR [64][33] complex, KR [max(KH,HO)][33] complex; 33 = odd pitch, conflict-free both ways):
//   R   rows     unit = two REAL rows packed into one complex row z = a + i*b:  (x_r, k_r) for r < KH, then (x_r, x_{r+np}).
//                Z = FFT(z);  2A(f) = Z(f) + conj Z(-f),  2B(f) = (Z(f) - conj Z(-f)) / i.   Only f = 0..32 is kept
//                (Hermitian); A(0), A(32) are real and share slot 0 -> 32 complex per row.  -> XR rows (x), KR rows (k).
//   CX  columns  unit = frequency column f < 32 of XR:  X^ = FFT_r(2X[:, f])  (in place, 64 rows).
//   CK  columns  K^ = FFT_r(2K[:, f]);  conj(P) = conj(X^) * K^ stored over X^.   Column 0 is the packed pair of the
//                (real-row) columns 0 and 32: lane 0 separates them, multiplies, re-packs.
//   CI  columns  conj(c~[:, f]) = FFT_fr(conj P)  -> KR rows i < HO   (c~ = IFFT_fr(P): spatial rows, frequency columns)
//   O   outputs  unit = output row i of a PAIR of planes (p, q):  Q(f) = c~_p(i,f) + i*c~_q(i,f), extended to f > 32 by
//                Hermitian symmetry;  out_p(i,:) + i*out_q(i,:) = IFFT(Q) = conj(FFT(conj Q)).
// Scale: two factors of 2 and two unnormalised inverse transforms = 4 * 64 * 64 = 2^14 (exact).
//
// Host-compilable: tests/native/host_fft_check.cpp runs these phases task by task on the CPU against a direct correlation.
#pragma once
#include "fft64.cuh"

#if !defined(__CUDACC__)
struct float2 {
    float x, y;
};
#endif

namespace hdn {

template <int KH_, int KW_, int HX_, int WX_, bool CIRC_, int G_, int NT_>
struct FCfg {
    static constexpr int KH = KH_, KW = KW_, HX = HX_, WX = WX_, G = G_, NT = NT_;
    static constexpr bool CIRC = CIRC_;
    static constexpr int PH = CIRC ? HX / 2 : 0, PW = CIRC ? WX / 2 : 0;
    static constexpr int HP = HX + 2 * PH, WP = WX + 2 * PW;  // padded input extent
    static constexpr int HO = HP - KH + 1, WO = WP - KW + 1;
    static constexpr int XPL = HX * WX, KPL = KH * KW, OPL = HO * WO;
    static constexpr int PITCH = 33;
    static constexpr int KR_ROWS = KH > HO ? KH : HO;
    static constexpr int XR_PLANE = 64 * PITCH, KR_PLANE = KR_ROWS * PITCH;  // complex elements
    // A group's x / k planes are fetched as 16-byte-aligned windows (TMA bulk copies need 16-byte addresses and sizes): a group of 2
    // planes of odd size starts 0 or 2 floats past a 16-byte boundary, so the window is the group rounded out to multiples of 4 floats.
    // (offset 0: the window ends 2 floats into the next group -- never the last one, whose offset is 2 because B*C % 4 == 0.)
    static constexpr int XWIN = (G * XPL + (G % 4 && XPL % 2 ? 2 : 0) + 3) / 4 * 4, KWIN = (G * KPL + (G % 4 && KPL % 2 ? 2 : 0) + 3) / 4 * 4;
    static constexpr int RAW_FLOATS = XWIN + KWIN, OUT_FLOATS = G * OPL;
    static constexpr int OUT_BUFS = G % 4 ? 1 : 2;  // 4-plane tiles leave by (asynchronous) TMA bulk store: double-buffered
    static constexpr int R_PAIRS = (HP - KH + 1) / 2;         // the x rows without a kernel row are packed in pairs (r, r + R_PAIRS):
    static constexpr int R_UNITS_PLANE = KH + R_PAIRS;        // consecutive units read consecutive rows (odd pitch: no bank conflicts)
    static constexpr int R_UNITS = G * R_UNITS_PLANE, C_UNITS = G * 32, O_UNITS = (G / 2) * HO;
    static constexpr int pad32(int n) { return (n + 31) / 32 * 32; }
    // tasks = (unit, half); warps alternate halves so a warp runs one half only:  task t -> h = (t/32) % 2, unit = (t/64)*32 + t%32
    static constexpr int R_TASKS = 2 * pad32(R_UNITS), C_TASKS = 2 * pad32(C_UNITS), O_TASKS = 2 * pad32(O_UNITS);
    static constexpr unsigned long long SMEM =
        (unsigned long long)(RAW_FLOATS + OUT_BUFS * OUT_FLOATS) * 4 + (unsigned long long)G * (XR_PLANE + KR_PLANE) * 8 + 16;
    static constexpr int CTAS = SMEM <= 75 * 1024 ? 3 : (SMEM <= 113 * 1024 ? 2 : 1);  // resident CTAs per SM (228 KB of shared memory)
    static_assert(HP <= 64 && WP <= 64, "padded input must fit the 64-point transform");
    static_assert(KH <= HP && KW <= WP && HO <= 64 && WO <= 64, "shape");
    static_assert(G % 2 == 0, "planes are paired in phase O");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct FftBufs {
    const float *rawx, *rawk;  // landed planes of the group (dense)
    float2 *XR, *KR;
    float *out;  // output tile of the group (dense)
};

enum { FFT_PH_R = 0, FFT_PH_CX = 1, FFT_PH_CK = 2, FFT_PH_CI = 3, FFT_PH_O = 4, FFT_PHASES = 5 };

HDN_HD int fft_task_half(int t) { return (t >> 5) & 1; }
HDN_HD int fft_task_unit(int t) { return ((t >> 6) << 5) | (t & 31); }

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

// ---- loads: all 64 complex inputs of the unit's transform -----------------------------------------------------------------
template <class Cfg>
HDN_HD bool fftc_load(int ph, const FftBufs &b, int unit, float (&re)[64], float (&im)[64]) {
    if (ph == FFT_PH_R) {
        if (unit >= Cfg::R_UNITS) return false;
        const int p = unit / Cfg::R_UNITS_PLANE, j = unit - p * Cfg::R_UNITS_PLANE;
        const bool ktype = j < Cfg::KH;
        const int r1 = j;  // j < KH: (x_j, k_j);  else (x_j, x_{j + R_PAIRS})
        const float *xrow = b.rawx + p * Cfg::XPL + fft_src_row<Cfg>(r1) * Cfg::WX;
#pragma unroll
        for (int n = 0; n < 64; ++n) {
            int q = n - Cfg::PW;  // compile-time: replicate padding of the columns
            q = q < 0 ? 0 : (q > Cfg::WX - 1 ? Cfg::WX - 1 : q);
            re[n] = n < Cfg::WP ? xrow[q] : 0.f;
        }
        if (ktype) {
            const float *krow = b.rawk + p * Cfg::KPL + j * Cfg::KW;
#pragma unroll
            for (int n = 0; n < 64; ++n) im[n] = n < Cfg::KW ? krow[n] : 0.f;
        } else {
            const int r2 = r1 + Cfg::R_PAIRS;
            const bool has2 = r2 < Cfg::HP;
            const float *xrow2 = b.rawx + p * Cfg::XPL + fft_src_row<Cfg>(has2 ? r2 : r1) * Cfg::WX;
#pragma unroll
            for (int n = 0; n < 64; ++n) {
                int q = n - Cfg::PW;
                q = q < 0 ? 0 : (q > Cfg::WX - 1 ? Cfg::WX - 1 : q);
                im[n] = (n < Cfg::WP && has2) ? xrow2[q] : 0.f;
            }
        }
        return true;
    }
    if (ph == FFT_PH_O) {
        if (unit >= Cfg::O_UNITS) return false;
        const int m = unit / Cfg::HO, i = unit - m * Cfg::HO;
        const float2 *rp = b.KR + (2 * m) * Cfg::KR_PLANE + i * Cfg::PITCH, *rq = rp + Cfg::KR_PLANE;
        // stored rows are conj(c~) (slot 0 = (c~(0), -c~(32)), both real); build conj(Q), Q(f) = c~_p(f) + i*c~_q(f), Q(-f) by symmetry
        {
            const float2 a = rp[0], c = rq[0];
            re[0] = a.x; im[0] = -c.x; re[32] = -a.y; im[32] = c.y;
        }
#pragma unroll
        for (int f = 1; f < 32; ++f) {
            const float2 a = rp[f], c = rq[f];
            re[f] = a.x + c.y; im[f] = a.y - c.x;
            re[64 - f] = a.x - c.y; im[64 - f] = -a.y - c.x;
        }
        return true;
    }
    // column phases: unit = (plane, frequency column f)
    if (unit >= Cfg::C_UNITS) return false;
    const int p = unit >> 5, f = unit & 31;
    if (ph == FFT_PH_CK) {
        const float2 *kc = b.KR + p * Cfg::KR_PLANE + f;
#pragma unroll
        for (int r = 0; r < 64; ++r) {
            if (r < Cfg::KH) {
                const float2 t = kc[r * Cfg::PITCH];
                re[r] = t.x; im[r] = t.y;
            } else {
                re[r] = im[r] = 0.f;
            }
        }
    } else if (ph == FFT_PH_CX) {  // row spectra: HP rows, the rest is zero padding (never read)
        const float2 *xc = b.XR + p * Cfg::XR_PLANE + f;
#pragma unroll
        for (int r = 0; r < 64; ++r) {
            if (r < Cfg::HP) {
                const float2 t = xc[r * Cfg::PITCH];
                re[r] = t.x; im[r] = t.y;
            } else {
                re[r] = im[r] = 0.f;
            }
        }
    } else {  // CI: conj(P), all 64 rows
        const float2 *xc = b.XR + p * Cfg::XR_PLANE + f;
#pragma unroll
        for (int r = 0; r < 64; ++r) {
            const float2 t = xc[r * Cfg::PITCH];
            re[r] = t.x; im[r] = t.y;
        }
    }
    return true;
}

// ---- stores: the half's 32 outputs X[2m + h] are at (re, im)[POS32(m)] ------------------------------------------------------
// Phase R needs the half as a compile-time constant (the partner X[64 - f] of X[f] sits at a register index that depends on it).
template <class Cfg, int H>
HDN_HD void fftc_store_R(const FftBufs &b, int unit, const float (&re)[64], const float (&im)[64]) {
    using namespace fft;
    const int p = unit / Cfg::R_UNITS_PLANE, j = unit - p * Cfg::R_UNITS_PLANE;
    const bool ktype = j < Cfg::KH;
    const int r1 = j;
    float2 *d1 = b.XR + p * Cfg::XR_PLANE + r1 * Cfg::PITCH;
    float2 *d2 = ktype ? b.KR + p * Cfg::KR_PLANE + j * Cfg::PITCH : d1 + Cfg::R_PAIRS * Cfg::PITCH;
    const bool has2 = ktype || r1 + Cfg::R_PAIRS < Cfg::HP;
    if (H == 0) {
        d1[0] = float2{2.f * re[HPOS(0)], 2.f * re[HPOS(32)]};
        if (has2) d2[0] = float2{2.f * im[HPOS(0)], 2.f * im[HPOS(32)]};
    }
#pragma unroll
    for (int f = 2 - H; f < 32; f += 2) {  // 2A(f) = Z(f) + conj Z(-f),  2B(f) = (Z(f) - conj Z(-f)) / i
        const float ar = re[HPOS(f)], ai = im[HPOS(f)], br = re[HPOS(64 - f)], bi = im[HPOS(64 - f)];
        d1[f] = float2{ar + br, ai - bi};
        if (has2) d2[f] = float2{ai + bi, br - ar};
    }
}

template <class Cfg>
HDN_HD void fftc_store(int ph, const FftBufs &b, int unit, int h, const float (&re)[64], const float (&im)[64]) {
    using namespace fft;
    if (ph == FFT_PH_R) {
        if (h == 0) fftc_store_R<Cfg, 0>(b, unit, re, im);
        else fftc_store_R<Cfg, 1>(b, unit, re, im);
        return;
    }
    if (ph == FFT_PH_O) {
        const int m = unit / Cfg::HO, i = unit - m * Cfg::HO;
        constexpr float SCALE = 1.0f / 16384.0f;
        float *op = b.out + (2 * m) * Cfg::OPL + i * Cfg::WO + h, *oq = op + Cfg::OPL;
#pragma unroll
        for (int mm = 0; 2 * mm < Cfg::WO; ++mm) {
            if (2 * mm + 1 < Cfg::WO || h == 0) {
                op[2 * mm] = re[POS32(mm)] * SCALE;
                oq[2 * mm] = im[POS32(mm)] * -SCALE;
            }
        }
        return;
    }
    const int p = unit >> 5, f = unit & 31;
    float2 *xc = b.XR + p * Cfg::XR_PLANE + f + h * Cfg::PITCH;  // row 2m + h of column f
    if (ph == FFT_PH_CX) {
#pragma unroll
        for (int mm = 0; mm < 32; ++mm) xc[2 * mm * Cfg::PITCH] = float2{re[POS32(mm)], im[POS32(mm)]};
    } else if (ph == FFT_PH_CI) {
        float2 *kc = b.KR + p * Cfg::KR_PLANE + f + h * Cfg::PITCH;
#pragma unroll
        for (int mm = 0; 2 * mm < Cfg::HO; ++mm)
            if (2 * mm + 1 < Cfg::HO || h == 0) kc[2 * mm * Cfg::PITCH] = float2{re[POS32(mm)], im[POS32(mm)]};
    } else if (f != 0) {  // CK: conj(P) = conj(X^) * K^ over X^
#pragma unroll
        for (int mm = 0; mm < 32; ++mm) {
            const float2 x = xc[2 * mm * Cfg::PITCH];
            const float kr = re[POS32(mm)], ki = im[POS32(mm)];
            xc[2 * mm * Cfg::PITCH] = float2{x.x * kr + x.y * ki, x.x * ki - x.y * kr};
        }
    } else {
        // CK, column 0 = the packed pair of real-row columns (0 and 32): W = V0 + i*V32 with V0, V32 Hermitian in fr.  For (fr, -fr):
        //   4*X0 = A + conj B,  4*X32 = (A - conj B)/i   (A = Wx(fr), B = Wx(-fr));  likewise K from C = Wk(fr), D = Wk(-fr)
        //   Q(fr) = P0 + i*P32,  Q(-fr) = conj P0 + i*conj P32,  P = X * conj K;  conj(Q) is stored; 1/4 restores the scale.
        // One lane per plane runs this, so it is a ROLLED loop over shared memory (K^ parked in the spare pad column 32 of XR) to
        // keep it out of the instruction-cache footprint; fr and -fr have the parity of h, i.e. both were written by this thread.
        float2 *kp = xc + 32;
#pragma unroll
        for (int mm = 0; mm < 32; ++mm) kp[2 * mm * Cfg::PITCH] = float2{re[POS32(mm)], im[POS32(mm)]};
        float2 *c0 = b.XR + p * Cfg::XR_PLANE;
#pragma unroll 1
        for (int fr = h; fr <= 32; fr += 2) {
            const int mf = (64 - fr) & 63;
            const float2 A = c0[fr * Cfg::PITCH], B = c0[mf * Cfg::PITCH], Cc = c0[fr * Cfg::PITCH + 32], D = c0[mf * Cfg::PITCH + 32];
            const float ur = A.x + B.x, ui = A.y - B.y, vr = A.y + B.y, vi = B.x - A.x;
            const float sr = Cc.x + D.x, si = Cc.y - D.y, tr = Cc.y + D.y, ti = D.x - Cc.x;
            const float p0r = 0.25f * (ur * sr + ui * si), p0i = 0.25f * (ui * sr - ur * si);
            const float p1r = 0.25f * (vr * tr + vi * ti), p1i = 0.25f * (vi * tr - vr * ti);
            c0[fr * Cfg::PITCH] = float2{p0r - p1i, -p0i - p1r};
            if (mf != fr) c0[mf * Cfg::PITCH] = float2{p0r + p1i, p0i - p1r};
        }
    }
}

// number of tasks of a phase
template <class Cfg>
HDN_HD int fftc_tasks(int ph) {
    return ph == FFT_PH_R ? Cfg::R_TASKS : (ph == FFT_PH_O ? Cfg::O_TASKS : Cfg::C_TASKS);
}

}  // namespace hdn
