// fft64.cuh -- a 64-point complex FFT split over TWO threads and held entirely in their registers, the building block of the
// transform-domain correlation kernels (xcorr_fft.cu).
//
//   X[f] = sum_n a[n] * exp(-2*pi*i * f*n / 64)         (forward, unnormalised; inverses are taken as conj(FFT(conj .)))
//
// Decimation in frequency: half h in {0,1} owns the outputs X[2m + h] = FFT32(s)[m] with
//   s[n] = a[n] + a[n+32]  (h = 0),     s[n] = (a[n] - a[n+32]) * W64^n  (h = 1),      n < 32.
// The caller forms a[n] +- a[n+32] while loading (one FFMA2 per complex value with the sign as a warp-uniform operand), applies
// half_twiddle() when h = 1 and runs fft32_fwd(): a radix-4 x 8 FFT on float2 (re, im) register pairs with the packed fp32x2
// instructions of sm_100 (a complex add or subtract is one issue slot), fully unrolled, twiddles are compile-time immediates
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

#if !defined(__CUDACC__)
struct float2 {
    float x, y;
};
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

// ---- packed fp32x2 arithmetic: a complex value is a float2 (re, im) in an aligned register pair; sm_100 adds / multiplies /
//      FMAs both halves with ONE instruction (add.f32x2, mul.f32x2, fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2), i.e. one issue slot
HDN_HD float2 f2(float x, float y) { return float2{x, y}; }
HDN_HD float2 fma2(float2 a, float2 b, float2 c) {  // a * b + c
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return float2{a.x * b.x + c.x, a.y * b.y + c.y};
#endif
}
HDN_HD float2 add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return float2{a.x + b.x, a.y + b.y};
#endif
}
HDN_HD float2 mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return float2{a.x * b.x, a.y * b.y};
#endif
}
HDN_HD float2 sub2(float2 a, float2 b) { return fma2(b, f2(-1.f, -1.f), a); }
HDN_HD float2 swp(float2 a) { return float2{a.y, a.x}; }
HDN_HD float2 add_mi(float2 a, float2 t) { return fma2(swp(t), f2(1.f, -1.f), a); }  // a + (-i) * t
HDN_HD float2 add_pi(float2 a, float2 t) { return fma2(swp(t), f2(-1.f, 1.f), a); }  // a + (+i) * t

// z * exp(-2*pi*i * K / 64)
template <int K>
HDN_HD float2 twiddle(float2 z) {
    constexpr int k = K & 63;
    if (k == 0) return z;
    if (k == 16) return float2{z.y, -z.x};
    if (k == 32) return float2{-z.x, -z.y};
    if (k == 48) return float2{-z.y, z.x};
    constexpr float c = cos64(k), s = sin64(k);  // (r + i*m)(c - i*s) = (r*c + m*s) + i*(m*c - r*s)
    return fma2(swp(z), f2(s, -s), mul2(z, f2(c, c)));
}

// (s[N]) *= W64^N, N = 1..31: the h = 1 half's twiddles
template <int N = 1>
struct HalfTw {
    static HDN_HD void run(float2 (&s)[32]) {
        s[N] = twiddle<N>(s[N]);
        HalfTw<N + 1>::run(s);
    }
};
template <>
struct HalfTw<32> {
    static HDN_HD void run(float2 (&)[32]) {}
};
HDN_HD void half_twiddle(float2 (&s)[32]) { HalfTw<>::run(s); }

// forward DFTs, in place, natural order in and out;  W4 = -i, W8 = (1 - i)/sqrt2, W8^3 = (-1 - i)/sqrt2
HDN_HD void dft4_fwd(float2 (&v)[4]) {
    const float2 b0 = add2(v[0], v[2]), b1 = sub2(v[0], v[2]), b2 = add2(v[1], v[3]), b3 = sub2(v[1], v[3]);
    v[0] = add2(b0, b2);
    v[2] = sub2(b0, b2);
    v[1] = add_mi(b1, b3);
    v[3] = add_pi(b1, b3);
}

HDN_HD void dft8_fwd(float2 (&v)[8]) {
    constexpr float H = 0.70710678118654752440f;
    const float2 a0 = add2(v[0], v[4]), a1 = sub2(v[0], v[4]), a2 = add2(v[2], v[6]), a3 = sub2(v[2], v[6]);
    const float2 a4 = add2(v[1], v[5]), a5 = sub2(v[1], v[5]), a6 = add2(v[3], v[7]), a7 = sub2(v[3], v[7]);
    // even outputs: DFT4 of (a0, a4, a2, a6)
    const float2 b0 = add2(a0, a2), b1 = sub2(a0, a2), b2 = add2(a4, a6), b3 = sub2(a4, a6);
    v[0] = add2(b0, b2);
    v[4] = sub2(b0, b2);
    v[2] = add_mi(b1, b3);
    v[6] = add_pi(b1, b3);
    // odd outputs: DFT4 of (a1, a5*W8, a3*W4, a7*W8^3)
    const float2 z1 = fma2(swp(a5), f2(H, -H), mul2(a5, f2(H, H)));    // H * (a5 + (-i)*a5)
    const float2 z3 = fma2(swp(a7), f2(H, -H), mul2(a7, f2(-H, -H)));  // H * ((-i)*a7 - a7)
    const float2 c0 = add_mi(a1, a3), c1 = add_pi(a1, a3);             // a1 +- (-i)*a3
    const float2 c2 = add2(z1, z3), c3 = sub2(z1, z3);
    v[1] = add2(c0, c2);
    v[5] = sub2(c0, c2);
    v[3] = add_mi(c1, c3);
    v[7] = add_pi(c1, c3);
}

// 32-point forward FFT, natural order in, Y[m] left at POS32(m).
HDN_HD constexpr int POS32(int m) { return 8 * (m & 3) + (m >> 2); }

template <int B = 0>
struct Pass32 {  // for each b: DFT4 over a (stride 8), twiddle by W32^(bc) = W64^(2bc), result at [8c + b]
    static HDN_HD void run(float2 (&s)[32]) {
        float2 v[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) v[a] = s[8 * a + B];
        dft4_fwd(v);
        s[B] = v[0];
        s[8 + B] = twiddle<2 * B * 1>(v[1]);
        s[16 + B] = twiddle<2 * B * 2>(v[2]);
        s[24 + B] = twiddle<2 * B * 3>(v[3]);
        Pass32<B + 1>::run(s);
    }
};
template <>
struct Pass32<8> {
    static HDN_HD void run(float2 (&)[32]) {}
};

HDN_HD void fft32_fwd(float2 (&s)[32]) {
    Pass32<>::run(s);
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // DFT8 over b of the block [8c + b] -> [8c + d] = Y[c + 4d]
        float2 v[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) v[b] = s[8 * c + b];
        dft8_fwd(v);
#pragma unroll
        for (int d = 0; d < 8; ++d) s[8 * c + d] = v[d];
    }
}

// X[f] of the half that owns parity f % 2:  X[f] = Y[f / 2]
HDN_HD constexpr int HPOS(int f) { return POS32(f >> 1); }

}  // namespace fft
}  // namespace hdn
