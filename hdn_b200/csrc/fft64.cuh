// fft64.cuh -- a 64-point complex FFT held entirely in one thread's registers (radix 8 x 8, fully unrolled, twiddles are
// compile-time immediates), the building block of the FFT correlation kernels (xcorr_fft.cu).
//
// Host-compilable (tests/host_fft_check.cpp builds it with g++ to check the arithmetic against a direct correlation):
// everything here is plain C++ on float arrays with compile-time indices.
//
//   X[f] = sum_n v[n] * exp(S * 2*pi*i * f*n / 64)        S = -1 forward, +1 inverse (unnormalised)
//
// With n = 8a + b and f = c + 8d:   exp(..fn/64) = W8^(ac) * W64^(bc) * W8^(bd).
//   fft64_nr: natural-order input (v[n]), output X[f] left at position POS(f) = 8*(f%8) + f/8   ("digit reversed")
//   fft64_rn: input X[f] at position POS(f), natural-order output -- the transposed flow graph; a forward nr followed by
//             an inverse rn therefore needs no reordering in between (point-wise products are done in place).
// NZ  = number of leading non-zero inputs of fft64_nr (the rest are never read: zero padding is pruned);
// REAL = the imaginary parts of the inputs are zero (never read).
#pragma once

#if defined(__CUDACC__)
#define HDN_HD __host__ __device__ __forceinline__
#else
#define HDN_HD inline
#endif

namespace hdn {
namespace fft {

HDN_HD constexpr int POS(int f) { return 8 * (f & 7) + (f >> 3); }

// cos(2*pi*k/64), k = 0..16 (quarter wave); everything else by symmetry, folded at compile time after unrolling
HDN_HD constexpr float cos64(int k) {
    constexpr float Q[17] = {1.0f,
                             0.99518472667219688624f,
                             0.98078528040323044913f,
                             0.95694033573220886494f,
                             0.92387953251128675613f,
                             0.88192126434835502971f,
                             0.83146961230254523708f,
                             0.77301045336273696081f,
                             0.70710678118654752440f,
                             0.63439328416364549822f,
                             0.55557023301960222474f,
                             0.47139673682599764856f,
                             0.38268343236508977173f,
                             0.29028467725446236764f,
                             0.19509032201612826785f,
                             0.09801714032956060199f,
                             0.0f};
    k &= 63;
    if (k > 32) k = 64 - k;
    return k > 16 ? -Q[32 - k] : Q[k];
}
HDN_HD constexpr float sin64(int k) { return cos64(k - 16); }

// (r, i) *= exp(S * 2*pi*i * K / 64)
template <int S, int K>
HDN_HD void twiddle(float &r, float &i) {
    constexpr int k = K & 63;
    if (k == 0) return;
    if (k == 16) { const float t = r; r = -S * i; i = S * t; return; }
    if (k == 32) { r = -r; i = -i; return; }
    if (k == 48) { const float t = r; r = S * i; i = -S * t; return; }
    constexpr float c = cos64(k), s = S * sin64(k);
    const float t = r * c - i * s;
    i = r * s + i * c;
    r = t;
}

// 8-point DFT, in place, natural order in and out.  NA = number of leading non-zero inputs (4 or 8).
template <int S, int NA = 8, bool REAL = false>
HDN_HD void dft8(float (&r)[8], float (&i)[8]) {
    constexpr float H = 0.70710678118654752440f;
    float a0r, a1r, a2r, a3r, a4r, a5r, a6r, a7r, a0i, a1i, a2i, a3i, a4i, a5i, a6i, a7i;
    if (NA <= 4) {  // x4..x7 = 0
        a0r = a1r = r[0]; a2r = a3r = r[2]; a4r = a5r = r[1]; a6r = a7r = r[3];
        a0i = a1i = i[0]; a2i = a3i = i[2]; a4i = a5i = i[1]; a6i = a7i = i[3];
    } else {
        a0r = r[0] + r[4]; a1r = r[0] - r[4]; a2r = r[2] + r[6]; a3r = r[2] - r[6];
        a4r = r[1] + r[5]; a5r = r[1] - r[5]; a6r = r[3] + r[7]; a7r = r[3] - r[7];
        a0i = i[0] + i[4]; a1i = i[0] - i[4]; a2i = i[2] + i[6]; a3i = i[2] - i[6];
        a4i = i[1] + i[5]; a5i = i[1] - i[5]; a6i = i[3] + i[7]; a7i = i[3] - i[7];
    }
    if (REAL) a0i = a1i = a2i = a3i = a4i = a5i = a6i = a7i = 0.f;
    // even outputs: DFT4 of (a0, a4, a2, a6);  W4 = S*i
    const float b0r = a0r + a2r, b0i = a0i + a2i, b1r = a0r - a2r, b1i = a0i - a2i;
    const float b2r = a4r + a6r, b2i = a4i + a6i, b3r = a4r - a6r, b3i = a4i - a6i;
    r[0] = b0r + b2r; i[0] = b0i + b2i;
    r[4] = b0r - b2r; i[4] = b0i - b2i;
    r[2] = b1r - S * b3i; i[2] = b1i + S * b3r;  // b1 + (S*i)*b3
    r[6] = b1r + S * b3i; i[6] = b1i - S * b3r;
    // odd outputs: DFT4 of (z0, z1, z2, z3) = (a1, a5*W8, a3*W4, a7*W8^3)
    const float z1r = H * (a5r - S * a5i), z1i = H * (a5i + S * a5r);    // a5 * (1 + S*i)/sqrt2
    const float z3r = H * (-a7r - S * a7i), z3i = H * (-a7i + S * a7r);  // a7 * (-1 + S*i)/sqrt2
    const float z2r = -S * a3i, z2i = S * a3r;                           // a3 * (S*i)
    const float c0r = a1r + z2r, c0i = a1i + z2i, c1r = a1r - z2r, c1i = a1i - z2i;
    const float c2r = z1r + z3r, c2i = z1i + z3i, c3r = z1r - z3r, c3i = z1i - z3i;
    r[1] = c0r + c2r; i[1] = c0i + c2i;
    r[5] = c0r - c2r; i[5] = c0i - c2i;
    r[3] = c1r - S * c3i; i[3] = c1i + S * c3r;
    r[7] = c1r + S * c3i; i[7] = c1i - S * c3r;
}

template <int S, int B, int C = 0>
struct TwiddleRow {  // (r[c], i[c]) *= W64^(S*B*c), c = C..7
    static HDN_HD void run(float (&r)[8], float (&i)[8]) {
        twiddle<S, B * C>(r[C], i[C]);
        TwiddleRow<S, B, C + 1>::run(r, i);
    }
};
template <int S, int B>
struct TwiddleRow<S, B, 8> {
    static HDN_HD void run(float (&)[8], float (&)[8]) {}
};

template <int S, int NZ, bool REAL, int B = 0>
struct Pass1 {  // nr, step 1+2: for each b, DFT8 over a (stride-8 inputs), twiddle by W64^(bc), result at [8c + b]
    static HDN_HD void run(float (&re)[64], float (&im)[64]) {
        constexpr int NA = (NZ - B + 7) / 8;  // inputs 8a + B < NZ  <=>  a < NA
        float r[8], i[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            r[a] = a < NA ? re[8 * a + B] : 0.f;
            i[a] = (a < NA && !REAL) ? im[8 * a + B] : 0.f;
        }
        dft8<S, (NA <= 4 ? 4 : 8), REAL>(r, i);
        TwiddleRow<S, B>::run(r, i);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            re[8 * c + B] = r[c];
            im[8 * c + B] = i[c];
        }
        Pass1<S, NZ, REAL, B + 1>::run(re, im);
    }
};
template <int S, int NZ, bool REAL>
struct Pass1<S, NZ, REAL, 8> {
    static HDN_HD void run(float (&)[64], float (&)[64]) {}
};

// natural-order input -> X[f] at POS(f)
template <int S, int NZ = 64, bool REAL = false>
HDN_HD void fft64_nr(float (&re)[64], float (&im)[64]) {
    Pass1<S, NZ, REAL>::run(re, im);
#pragma unroll
    for (int c = 0; c < 8; ++c) {  // step 3: DFT8 over b of the contiguous block [8c + b] -> [8c + d] = X[c + 8d]
        float r[8], i[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) { r[b] = re[8 * c + b]; i[b] = im[8 * c + b]; }
        dft8<S>(r, i);
#pragma unroll
        for (int d = 0; d < 8; ++d) { re[8 * c + d] = r[d]; im[8 * c + d] = i[d]; }
    }
}

template <int S, int C = 0>
struct PassT {  // rn, step 1'+2': for each c, DFT8 over d of the block [8c + d], twiddle by W64^(bc), result at [8c + b]
    static HDN_HD void run(float (&re)[64], float (&im)[64]) {
        float r[8], i[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) { r[d] = re[8 * C + d]; i[d] = im[8 * C + d]; }
        dft8<S>(r, i);
        TwiddleRow<S, C>::run(r, i);  // index here is b; the factor W64^(b*C) is symmetric in (b, c)
#pragma unroll
        for (int b = 0; b < 8; ++b) { re[8 * C + b] = r[b]; im[8 * C + b] = i[b]; }
        PassT<S, C + 1>::run(re, im);
    }
};
template <int S>
struct PassT<S, 8> {
    static HDN_HD void run(float (&)[64], float (&)[64]) {}
};

// X[f] at POS(f) -> natural-order output x[n] (outputs the caller never reads are dead code for the compiler)
template <int S>
HDN_HD void fft64_rn(float (&re)[64], float (&im)[64]) {
    PassT<S>::run(re, im);
#pragma unroll
    for (int b = 0; b < 8; ++b) {  // step 3': DFT8 over c (stride 8) -> x[8a + b]
        float r[8], i[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { r[c] = re[8 * c + b]; i[c] = im[8 * c + b]; }
        dft8<S>(r, i);
#pragma unroll
        for (int a = 0; a < 8; ++a) { re[8 * a + b] = r[a]; im[8 * a + b] = i[a]; }
    }
}

// ---- 64-point FFT split over two threads (decimation in frequency) ---------------------------------------------------------
// Half h in {0,1} of X = FFT64(a), S = -1:   X[2m + h] = FFT32( s )[m],   s[n] = a[n] + a[n+32]            (h = 0)
//                                                                          s[n] = (a[n] - a[n+32]) * W64^n  (h = 1)
// Both halves need all 64 inputs; each keeps 32 outputs.  half_butterfly leaves s in (re, im)[0..31].
template <int N = 1>
struct HalfTw {  // (re[N], im[N]) = (d[N]) * W64^(-N), N = 1..31
    static HDN_HD void run(float (&re)[64], float (&im)[64]) {
        twiddle<-1, N>(re[N], im[N]);
        HalfTw<N + 1>::run(re, im);
    }
};
template <>
struct HalfTw<32> {
    static HDN_HD void run(float (&)[64], float (&)[64]) {}
};

HDN_HD void half_butterfly(int h, float (&re)[64], float (&im)[64]) {
    if (h == 0) {
#pragma unroll
        for (int n = 0; n < 32; ++n) { re[n] += re[n + 32]; im[n] += im[n + 32]; }
    } else {
#pragma unroll
        for (int n = 0; n < 32; ++n) { re[n] -= re[n + 32]; im[n] -= im[n + 32]; }
        HalfTw<>::run(re, im);
    }
}

// 32-point forward FFT on (re, im)[0..31], natural order in, Y[m] left at POS32(m) = 8*(m%4) + m/4.
// n = 8a + b (a < 4, b < 8), m = c + 4d (c < 4, d < 8):  W32^(mn) = W4^(ac) * W32^(bc) * W8^(bd).
HDN_HD constexpr int POS32(int m) { return 8 * (m & 3) + (m >> 2); }

HDN_HD void dft4_fwd(float (&r)[4], float (&i)[4]) {  // W4 = -i
    const float b0r = r[0] + r[2], b0i = i[0] + i[2], b1r = r[0] - r[2], b1i = i[0] - i[2];
    const float b2r = r[1] + r[3], b2i = i[1] + i[3], b3r = r[1] - r[3], b3i = i[1] - i[3];
    r[0] = b0r + b2r; i[0] = b0i + b2i;
    r[2] = b0r - b2r; i[2] = b0i - b2i;
    r[1] = b1r + b3i; i[1] = b1i - b3r;  // b1 + (-i)*b3
    r[3] = b1r - b3i; i[3] = b1i + b3r;
}

template <int B = 0>
struct Pass32 {  // for each b: DFT4 over a (stride 8), twiddle by W32^(bc) = W64^(2bc), result at [8c + b]
    static HDN_HD void run(float (&re)[64], float (&im)[64]) {
        float r[4], i[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { r[a] = re[8 * a + B]; i[a] = im[8 * a + B]; }
        dft4_fwd(r, i);
        twiddle<-1, 2 * B * 1>(r[1], i[1]);
        twiddle<-1, 2 * B * 2>(r[2], i[2]);
        twiddle<-1, 2 * B * 3>(r[3], i[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) { re[8 * c + B] = r[c]; im[8 * c + B] = i[c]; }
        Pass32<B + 1>::run(re, im);
    }
};
template <>
struct Pass32<8> {
    static HDN_HD void run(float (&)[64], float (&)[64]) {}
};

HDN_HD void fft32_fwd(float (&re)[64], float (&im)[64]) {
    Pass32<>::run(re, im);
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // DFT8 over b of the block [8c + b] -> [8c + d] = Y[c + 4d]
        float r[8], i[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) { r[b] = re[8 * c + b]; i[b] = im[8 * c + b]; }
        dft8<-1>(r, i);
#pragma unroll
        for (int d = 0; d < 8; ++d) { re[8 * c + d] = r[d]; im[8 * c + d] = i[d]; }
    }
}

// X[f] of the half that owns parity f % 2, after half_butterfly + fft32_fwd:  X[f] = Y[f / 2]
HDN_HD constexpr int HPOS(int f) { return POS32(f >> 1); }

}  // namespace fft
}  // namespace hdn
