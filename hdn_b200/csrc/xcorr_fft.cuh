// xcorr_fft.cuh -- the three per-thread phases of the FFT depth-wise correlation (K1/K2 at the FMA-bound shapes).
//
//   out[i,j] = sum_{u,v} xp[i+u, j+v] * k[u,v]     xp = x (K1) or its circular-row / replicate-column padding (K2)
//
// is a 64x64 circular correlation whenever the padded input fits 64x64 (no wrap reaches a valid output), i.e.
//   OUT = IFFT2( FFT2(xp) * conj(FFT2(k)) ).
// Direct evaluation costs KH*KW FMAs per output (841 at 29x29: 81 flop/B, FMA-bound at 6-7x the HBM time); the FFT
// route costs ~175 register-resident 64-point FFTs per plane, ~5x fewer instructions.
//
// Per plane (all buffers in shared memory, one thread = one 64-point FFT held in registers, fft64.cuh):
//   R  rows     thread = padded input row r:  z = xp[r,:] + i*k[r,:]  -> Z = FFT(z);  since both are real,
//               2X(f) = Z(f) + conj Z(-f),  2K(f) = (Z(f) - conj Z(-f)) / i.   Only f = 0..32 is kept (Hermitian);
//               X(0) and X(32) are real and share slot 0 as (X(0), X(32)) -> 32 complex per row.
//               Rows r >= KH carry no kernel row: real-input FFT (half the outputs are dead code).
//   C  columns  thread = frequency column f (one warp per plane): X^ = FFT_r(2X[:,f]), K^ = FFT_r(2K[:,f]) (29 non-zero
//               inputs, pruned), P = X^ * conj(K^) in registers, c~[:,f] = IFFT_r(P), rows i < HO kept.
//               Lane 0 owns the packed (0, 32) column pair and separates / re-packs it (both are spectra of real rows).
//   O  outputs  thread = output row i of a PAIR of planes (p, q):  Q(f) = c~_p(i,f) + i*c~_q(i,f), extended to f > 32
//               by Hermitian symmetry, IFFT -> out_p(i,:) + i*out_q(i,:).
// Scale: the two factors of 2 and the two unnormalised inverse FFTs give 4 * 64 * 64 = 2^14 (exact power of two).
//
// Host-compilable: tests/host_fft_check.cpp runs these phases task by task on the CPU against a direct correlation.
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
    static constexpr int PITCH = 33;                // complex per spectral row: odd -> row-wise and column-wise accesses conflict-free
    static constexpr int XR_PLANE = 64 * PITCH;     // complex: row spectra (HP rows), then X^ (64 rows), then c~ (HO rows), in place
    static constexpr int KR_PLANE = KH * PITCH;
    static constexpr int RAW_FLOATS = G * (XPL + KPL), OUT_FLOATS = G * OPL;
    // task slots of phase R: complex rows (with a kernel row) first, real rows from the next warp boundary
    static constexpr int N_CPX = G * KH, CPX_PAD = (N_CPX + 31) / 32 * 32, N_REAL = G * (HP - KH), R_SLOTS = CPX_PAD + N_REAL;
    static constexpr int C_TASKS = G * 32, O_TASKS = (G / 2) * HO;
    static constexpr unsigned long long SMEM =
        (unsigned long long)(RAW_FLOATS + 2 * OUT_FLOATS) * 4 + (unsigned long long)G * (XR_PLANE + KR_PLANE) * 8 + 16;
    static_assert(HP <= 64 && WP <= 64, "padded input must fit the 64-point transform");
    static_assert(KH <= HP && KW <= WP && WO <= 33 + 31, "shape");
    static_assert(G % 4 == 0, "bulk copies need 16-byte multiples; planes are paired in phase O");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// ---- phase R -----------------------------------------------------------------------------------------------------------
// xrow: dense source row of x (already wrapped for the circular variant); krow: kernel row or nullptr-equivalent (HASK = false)
template <class Cfg, bool HASK>
HDN_HD void fftc_phase_row(const float *xrow, const float *krow, float2 *xr_row, float2 *kr_row) {
    using namespace fft;
    float re[64], im[64];
#pragma unroll
    for (int n = 0; n < 64; ++n) {
        if (n < Cfg::WP) {
            int q = n - Cfg::PW;  // compile-time: replicate padding of the columns
            q = q < 0 ? 0 : (q > Cfg::WX - 1 ? Cfg::WX - 1 : q);
            re[n] = xrow[q];
        } else {
            re[n] = 0.f;
        }
        im[n] = (HASK && n < Cfg::KW) ? krow[n] : 0.f;
    }
    fft64_nr<-1, Cfg::WP, !HASK>(re, im);
    if (HASK) {
        xr_row[0] = float2{2.f * re[POS(0)], 2.f * re[POS(32)]};
        kr_row[0] = float2{2.f * im[POS(0)], 2.f * im[POS(32)]};
#pragma unroll
        for (int f = 1; f < 32; ++f) {
            const float ar = re[POS(f)], ai = im[POS(f)], br = re[POS(64 - f)], bi = im[POS(64 - f)];
            xr_row[f] = float2{ar + br, ai - bi};
            kr_row[f] = float2{ai + bi, br - ar};
        }
    } else {
        xr_row[0] = float2{2.f * re[POS(0)], 2.f * re[POS(32)]};
#pragma unroll
        for (int f = 1; f < 32; ++f) xr_row[f] = float2{2.f * re[POS(f)], 2.f * im[POS(f)]};
    }
}

// slot -> task of phase R.  rawx / rawk / XR / KR point at the group's buffers.
template <class Cfg>
HDN_HD void fftc_phase_R(const float *rawx, const float *rawk, float2 *XR, float2 *KR, int slot) {
    int p, r;
    bool cpx;
    if (slot < Cfg::N_CPX) {
        p = slot / Cfg::KH; r = slot - p * Cfg::KH; cpx = true;
    } else if (slot >= Cfg::CPX_PAD) {
        const int t = slot - Cfg::CPX_PAD;
        p = t / (Cfg::HP - Cfg::KH); r = Cfg::KH + t - p * (Cfg::HP - Cfg::KH); cpx = false;
    } else {
        return;
    }
    int sr = r - Cfg::PH;
    if (Cfg::CIRC) {
        if (sr < 0) sr += Cfg::HX;
        else if (sr >= Cfg::HX) sr -= Cfg::HX;
    }
    const float *xrow = rawx + p * Cfg::XPL + sr * Cfg::WX;
    float2 *xr_row = XR + p * Cfg::XR_PLANE + r * Cfg::PITCH;
    if (cpx) fftc_phase_row<Cfg, true>(xrow, rawk + p * Cfg::KPL + r * Cfg::KW, xr_row, KR + p * Cfg::KR_PLANE + r * Cfg::PITCH);
    else fftc_phase_row<Cfg, false>(xrow, nullptr, xr_row, nullptr);
}

// ---- phase C -----------------------------------------------------------------------------------------------------------
template <class Cfg>
HDN_HD void fftc_phase_C(float2 *XR, const float2 *KR, int task) {
    using namespace fft;
    const int p = task >> 5, f = task & 31;
    float2 *xc = XR + p * Cfg::XR_PLANE + f;
    const float2 *kc = KR + p * Cfg::KR_PLANE + f;
    float re[64], im[64];
#pragma unroll
    for (int r = 0; r < 64; ++r) {
        if (r < Cfg::HP) {
            const float2 t = xc[r * Cfg::PITCH];
            re[r] = t.x; im[r] = t.y;
        } else {
            re[r] = im[r] = 0.f;
        }
    }
    fft64_nr<-1, Cfg::HP>(re, im);
#pragma unroll
    for (int fr = 0; fr < 64; ++fr) xc[fr * Cfg::PITCH] = float2{re[POS(fr)], im[POS(fr)]};  // X^(fr, f): own column, re-read below
#pragma unroll
    for (int r = 0; r < 64; ++r) {
        if (r < Cfg::KH) {
            const float2 t = kc[r * Cfg::PITCH];
            re[r] = t.x; im[r] = t.y;
        } else {
            re[r] = im[r] = 0.f;
        }
    }
    fft64_nr<-1, Cfg::KH>(re, im);
    if (f != 0) {
#pragma unroll
        for (int fr = 0; fr < 64; ++fr) {  // P = X^ * conj(K^), in place at POS(fr)
            const float2 x = xc[fr * Cfg::PITCH];
            const float kr = re[POS(fr)], ki = im[POS(fr)];
            re[POS(fr)] = x.x * kr + x.y * ki;
            im[POS(fr)] = x.y * kr - x.x * ki;
        }
    } else {
        // packed pair of real-row columns: W = V0 + i*V32 with V0, V32 Hermitian in fr.  For each (fr, -fr):
        //   4*X0 = A + conj B,  4*X32 = (A - conj B)/i   (A = Wx(fr), B = Wx(-fr));  likewise K from C = Wk(fr), D = Wk(-fr)
        //   Q(fr) = P0 + i*P32,  Q(-fr) = conj P0 + i*conj P32,   P = X * conj K   (the 1/4 restores the regular columns' scale)
#pragma unroll
        for (int fr = 0; fr <= 32; ++fr) {
            const int mf = (64 - fr) & 63;
            const float2 A = xc[fr * Cfg::PITCH], B = xc[mf * Cfg::PITCH];
            const float cr = re[POS(fr)], ci = im[POS(fr)], dr = re[POS(mf)], di = im[POS(mf)];
            const float ur = A.x + B.x, ui = A.y - B.y, vr = A.y + B.y, vi = B.x - A.x;
            const float sr = cr + dr, si = ci - di, tr = ci + di, ti = dr - cr;
            const float p0r = 0.25f * (ur * sr + ui * si), p0i = 0.25f * (ui * sr - ur * si);
            const float p1r = 0.25f * (vr * tr + vi * ti), p1i = 0.25f * (vi * tr - vr * ti);
            re[POS(fr)] = p0r - p1i; im[POS(fr)] = p0i + p1r;
            if (mf != fr) { re[POS(mf)] = p0r + p1i; im[POS(mf)] = p1r - p0i; }
        }
    }
    fft64_rn<+1>(re, im);
#pragma unroll
    for (int i = 0; i < Cfg::HO; ++i) xc[i * Cfg::PITCH] = float2{re[i], im[i]};
}

// ---- phase O -----------------------------------------------------------------------------------------------------------
template <class Cfg>
HDN_HD void fftc_phase_O(const float2 *XR, float *out, int task) {
    using namespace fft;
    const int m = task / Cfg::HO, i = task - m * Cfg::HO;
    const float2 *rp = XR + (2 * m) * Cfg::XR_PLANE + i * Cfg::PITCH, *rq = rp + Cfg::XR_PLANE;
    float re[64], im[64];
    {
        const float2 a = rp[0], c = rq[0];
        re[0] = a.x; im[0] = c.x; re[32] = a.y; im[32] = c.y;
    }
#pragma unroll
    for (int f = 1; f < 32; ++f) {
        const float2 a = rp[f], c = rq[f];
        re[f] = a.x - c.y; im[f] = a.y + c.x;
        re[64 - f] = a.x + c.y; im[64 - f] = c.x - a.y;
    }
    fft64_nr<+1>(re, im);
    constexpr float SCALE = 1.0f / 16384.0f;
    float *op = out + (2 * m) * Cfg::OPL + i * Cfg::WO, *oq = op + Cfg::OPL;
#pragma unroll
    for (int j = 0; j < Cfg::WO; ++j) {
        op[j] = re[POS(j)] * SCALE;
        oq[j] = im[POS(j)] * SCALE;
    }
}

}  // namespace hdn
