// conv_small.cu -- the few-channel convolutions of the path, which have no GEMM shape worth a tensor core (K7c):
//
//   the two 7x7 stride-2 stems          3 -> 64 (pad 0)  hdn/models/backbone/resnet_atrous.py:121-131
//                                       2 -> 64 (pad 3)  homo_estimator/Deep_homography/Oneline_DLTv1/backbone/resnet.py:142-160
//   PreShareFeature's 3x3 layers        1 -> 4 -> 8 -> 1 (pad 1)  Oneline_DLTv1/preprocess/input_feature_extractor.py:3-29
//
// each followed by an eval-mode BatchNorm and a ReLU, folded into the epilogue.  Direct fp32 FMA sum: a CTA owns a 32 x 8 tile of
// output pixels for COT output channels, stages the input window of all Cin channels and the [tap][channel] weights in shared memory
// and every thread accumulates two pixels x COT channels in registers (weights are warp-broadcast LDS.128, 2*COT FMAs per 2 + COT/4
// shared loads).  Row pitches of the staged window are chosen so that the 16 x 2 lanes of a warp hit 32 different banks.  With them
// no convolution of the path is left on cuDNN.
#include "common.cuh"

namespace hdn {

constexpr int CS_TX = 16, CS_TY = 8, CS_PX = 2;  // threads x, threads y, pixels per thread (ox = tx and tx + 16)
constexpr int CS_TW = CS_TX * CS_PX;              // 32 output columns per tile

template <int K, int S>
struct CSGeom {
    static constexpr int TH = (CS_TY - 1) * S + K;                   // staged rows
    static constexpr int TWr = (CS_TW - 1) * S + K;                  // staged columns in use
    // S = 1: lanes 16..31 sit one row below lanes 0..15 -> pitch = 16 (mod 32); S = 2: lanes step 2 words -> odd pitch
    static constexpr int PITCH = S == 1 ? ((TWr + 15) / 32) * 32 + 16 : (TWr | 1);
};

template <int COT, int K, int S>
__global__ void __launch_bounds__(CS_TX *CS_TY) conv_small_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                   const float *__restrict__ scale, const float *__restrict__ shift,
                                                                   float *__restrict__ out, int Cin, int Cout, int H, int W, int Ho, int Wo,
                                                                   int pad, int relu) {
    using G = CSGeom<K, S>;
    extern __shared__ __align__(16) float smem[];
    float *wsm = smem;                       // [Cin*K*K][COT]
    float *tile = smem + Cin * K * K * COT;  // [Cin][TH][PITCH]
    const int groups = Cout / COT;
    const int b = blockIdx.z / groups, co0 = (blockIdx.z % groups) * COT;
    const int tid = threadIdx.y * CS_TX + threadIdx.x;
    const int oy0 = blockIdx.y * CS_TY, ox0 = blockIdx.x * CS_TW;
    const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;

    const int taps = Cin * K * K;
    for (int i = tid; i < taps * COT; i += CS_TX * CS_TY) {
        const int t = i / COT, c = i % COT;  // w is [Cout][Cin][K][K]: tap index t = (ci, ky, kx)
        wsm[i] = w[(size_t)(co0 + c) * taps + t];
    }
    const float *xb = x + (size_t)b * Cin * H * W;
    for (int i = tid; i < Cin * G::TH * G::TWr; i += CS_TX * CS_TY) {
        const int c = i % G::TWr, r = (i / G::TWr) % G::TH, ci = i / (G::TWr * G::TH);
        const int iy = iy0 + r, ix = ix0 + c;
        tile[(ci * G::TH + r) * G::PITCH + c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xb + ((size_t)ci * H + iy) * W + ix) : 0.f;
    }
    __syncthreads();

    float acc[CS_PX][COT];
#pragma unroll
    for (int p = 0; p < CS_PX; ++p)
#pragma unroll
        for (int c = 0; c < COT; ++c) acc[p][c] = 0.f;

    const float *trow = tile + (threadIdx.y * S) * G::PITCH + threadIdx.x * S;
    for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const float *wp = wsm + ((ci * K + ky) * K + kx) * COT;
                const float *tp = trow + (ci * G::TH + ky) * G::PITCH + kx;
                const float v0 = tp[0], v1 = tp[CS_TX * S];
                if constexpr (COT % 4 == 0) {
#pragma unroll
                    for (int c = 0; c < COT; c += 4) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(wp + c);
                        acc[0][c] = fmaf(v0, w4.x, acc[0][c]), acc[0][c + 1] = fmaf(v0, w4.y, acc[0][c + 1]);
                        acc[0][c + 2] = fmaf(v0, w4.z, acc[0][c + 2]), acc[0][c + 3] = fmaf(v0, w4.w, acc[0][c + 3]);
                        acc[1][c] = fmaf(v1, w4.x, acc[1][c]), acc[1][c + 1] = fmaf(v1, w4.y, acc[1][c + 1]);
                        acc[1][c + 2] = fmaf(v1, w4.z, acc[1][c + 2]), acc[1][c + 3] = fmaf(v1, w4.w, acc[1][c + 3]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < COT; ++c) {
                        const float wv = wp[c];
                        acc[0][c] = fmaf(v0, wv, acc[0][c]), acc[1][c] = fmaf(v1, wv, acc[1][c]);
                    }
                }
            }
        }
    }

    const int oy = oy0 + threadIdx.y;
    if (oy >= Ho) return;
#pragma unroll
    for (int p = 0; p < CS_PX; ++p) {
        const int ox = ox0 + threadIdx.x + p * CS_TX;
        if (ox >= Wo) continue;
#pragma unroll
        for (int c = 0; c < COT; ++c) {
            const int co = co0 + c;
            float y = acc[p][c];
            if (scale) y *= __ldg(scale + co);
            if (shift) y += __ldg(shift + co);
            if (relu) y = fmaxf(y, 0.f);
            out[(((size_t)b * Cout + co) * Ho + oy) * Wo + ox] = y;  // 16 consecutive floats per half-warp and channel
        }
    }
}

template <int COT, int K, int S>
static int launch_small(const float *x, const float *w, const float *scale, const float *shift, float *out, int B, int Cin, int Cout, int H,
                        int W, int Ho, int Wo, int pad, int relu, cudaStream_t stream) {
    using G = CSGeom<K, S>;
    const size_t smem = sizeof(float) * ((size_t)Cin * K * K * COT + (size_t)Cin * G::TH * G::PITCH);
    if (smem > 48 * 1024) return HDN_ERR_UNSUPPORTED;
    const dim3 grid((Wo + CS_TW - 1) / CS_TW, (Ho + CS_TY - 1) / CS_TY, B * (Cout / COT));
    if (grid.z > 65535u) return HDN_ERR_SHAPE;
    conv_small_kernel<COT, K, S><<<grid, dim3(CS_TX, CS_TY), smem, stream>>>(x, w, scale, shift, out, Cin, Cout, H, W, Ho, Wo, pad, relu);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace hdn

extern "C" int hdn_conv_small_supported(int Cin, int Cout, int ksize, int stride) {
    if (Cin < 1 || Cin > 8 || Cout < 1) return 0;
    if (ksize == 7 && stride == 2) return Cout % 32 == 0;
    if (ksize == 3 && stride == 1) return Cout == 1 || Cout % 4 == 0;
    return 0;
}

extern "C" int hdn_conv_small_f32(const float *x, const float *w, const float *scale, const float *shift, float *out, int B, int Cin, int Cout,
                                  int H, int W, int ksize, int stride, int pad, int relu, hdn_stream_t stream_) {
    using namespace hdn;
    if (!x || !w || !out) return HDN_ERR_NULL;
    if (B < 1 || H < 1 || W < 1 || pad < 0 || pad > ksize / 2) return HDN_ERR_SHAPE;
    if (!hdn_conv_small_supported(Cin, Cout, ksize, stride)) return HDN_ERR_UNSUPPORTED;
    if ((((uintptr_t)x | (uintptr_t)w | (uintptr_t)out | (uintptr_t)scale | (uintptr_t)shift) & 3) != 0) return HDN_ERR_ALIGN;
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    if (Ho < 1 || Wo < 1) return HDN_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (ksize == 7) return launch_small<32, 7, 2>(x, w, scale, shift, out, B, Cin, Cout, H, W, Ho, Wo, pad, relu, stream);
    if (Cout % 8 == 0) return launch_small<8, 3, 1>(x, w, scale, shift, out, B, Cin, Cout, H, W, Ho, Wo, pad, relu, stream);
    if (Cout % 4 == 0) return launch_small<4, 3, 1>(x, w, scale, shift, out, B, Cin, Cout, H, W, Ho, Wo, pad, relu, stream);
    return launch_small<1, 3, 1>(x, w, scale, shift, out, B, Cin, Cout, H, W, Ho, Wo, pad, relu, stream);
}
