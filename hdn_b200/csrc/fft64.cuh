// fft64.cuh -- a 64-point complex FFT split over TWO threads and held entirely in their registers, the building block of the
// transform-domain correlation kernels (xcorr_fft.cu).
//
//   X[f] = sum_n a[n] * exp(-2*pi*i * f*n / 64)         (forward, unnormalised; inverses are taken as conj(FFT(conj .)))
//
// Decimation in frequency: half h in {0,1} owns the outputs X[2m + h] = FFT32(s)[m] with
//   s[n] = a[n] + a[n+32]  (h = 0),     s[n] = (a[n] - a[n+32]) * W64^n  (h = 1),      n < 32.
// The caller forms a[n] +- a[n+32] while loading (one FFMA per value with the sign as a warp-uniform operand), applies
// half_twiddle() when h = 1 and runs fft32_fwd(): a radix-4 x 8 FFT, fully unrolled, twiddles are compile-time immediates
// (n = 8a + b, m = c + 4d:  W32^(mn) = W4^(ac) * W32^(bc) * W8^(bd)).  Y[m] is left at POS32(m) = 8*(m%4) + m/4, i.e.
// X[f] of the owning half at HPOS(f).
//
// Host-compilable (tests/native/host_fft_check.cpp builds it with g++): plain C++ on float arrays with compile-time indices.
#pragma once

#if defined(__CUDACC__)
#define HDN_HD __host__ __device__ __forceinline__
#else
#define HDN_HD inline
#endif

namespace hdn {
namespace fft {

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

// (s[N]) *= W64^(-N), N = 1..31: the h = 1 half's twiddles
template <int N = 1>
struct HalfTw {
    static HDN_HD void run(float (&re)[32], float (&im)[32]) {
        twiddle<-1, N>(re[N], im[N]);
        HalfTw<N + 1>::run(re, im);
    }
};
template <>
struct HalfTw<32> {
    static HDN_HD void run(float (&)[32], float (&)[32]) {}
};
HDN_HD void half_twiddle(float (&re)[32], float (&im)[32]) { HalfTw<>::run(re, im); }

// 32-point forward FFT, natural order in, Y[m] left at POS32(m).
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
    static HDN_HD void run(float (&re)[32], float (&im)[32]) {
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
    static HDN_HD void run(float (&)[32], float (&)[32]) {}
};

HDN_HD void fft32_fwd(float (&re)[32], float (&im)[32]) {
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

// X[f] of the half that owns parity f % 2:  X[f] = Y[f / 2]
HDN_HD constexpr int HPOS(int f) { return POS32(f >> 1); }

}  // namespace fft
}  // namespace hdn
