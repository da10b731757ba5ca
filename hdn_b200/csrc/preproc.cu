// preproc.cu -- device-side frame pre-processing of the tracker, BIT-COMPATIBLE with the OpenCV calls of the
// reference so that the crops -- and therefore every arg-max downstream -- are the ones the reference would see:
//
//   hdn_warp_perspective_u8     cv2.warpPerspective(img, inv(H_total), (w, h), borderMode=BORDER_REPLICATE)   hdn_tracker_proj_e2e.py:154
//   hdn_crop_resize_u8          SiameseTracker.get_subwindow: mean-padded square window + cv2.resize            base_tracker.py:61-136
//                               (+ get_search_info's gray normalisation, Oneline_DLTv1/tools/get_img_info.py:42-70)
//   hdn_warp_affine_cubic_u8    img_rot_around_center: cv2.warpAffine(flags=2, BORDER_REPLICATE)               hdn/utils/transform.py:69-100
//
// The reference does these on the host with single-threaded OpenCV (13 + 30 ms of full-frame warps and three crops per 1280x720
// frame [probed]) and uploads three crops; here the uint8 frame is uploaded ONCE and every pixel operation runs on the device.
// OpenCV's routines are fixed-point (5-bit sampling positions, 15-bit table weights that are nudged to sum to 2^15, an 11-bit
// two-pass resize with its own shifts; INTER_LINEAR at exactly 1/2 is INTER_AREA), restated in oracle/cv_port.py and pinned there
// bit for bit against cv2 itself; these kernels follow the same integer arithmetic and the same double / float rounding steps
// (explicit __dmul_rn / __dadd_rn: no FMA contraction, as in OpenCV's C loops).  All of it is HBM/L2-bound byte work: a thread
// produces one output pixel (3 channels), reads are gathers around a smooth map, writes are coalesced.
#include "common.cuh"

namespace hdn {

__device__ __forceinline__ int cv_clip(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }
__device__ __forceinline__ int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
__device__ __forceinline__ unsigned char sat_u8(int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
// saturate_cast<int>(double) = cvRound: round half to even, then clamp
__device__ __forceinline__ int cv_round_clamped(double v) {
    v = fmax(-2147483648.0, fmin(2147483647.0, v));
    return (int)__double2ll_rn(v);
}

struct M9 {
    double m[9];
};

// ---------------------------------------------------------------------------------------------------- warpPerspective
// WarpPerspectiveInvoker (imgwarp.cpp): 16-row x bw-column blocks; X0 / Y0 / W0 at the block's first column, + M * x1 inside.
__global__ void __launch_bounds__(256) warp_perspective_u8_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, int H,
                                                                   int W, int bw, M9 M) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int xb = (x / bw) * bw;
    const double fx0 = (double)xb, fy = (double)y, x1 = (double)(x - xb);
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M.m[0], fx0), __dmul_rn(M.m[1], fy)), M.m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M.m[3], fx0), __dmul_rn(M.m[4], fy)), M.m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M.m[6], fx0), __dmul_rn(M.m[7], fy)), M.m[8]);
    double Wd = __dadd_rn(W0, __dmul_rn(M.m[6], x1));
    Wd = Wd != 0.0 ? __ddiv_rn(32.0, Wd) : 0.0;
    const int X = cv_round_clamped(__dmul_rn(__dadd_rn(X0, __dmul_rn(M.m[0], x1)), Wd));
    const int Y = cv_round_clamped(__dmul_rn(__dadd_rn(Y0, __dmul_rn(M.m[3], x1)), Wd));
    const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5), ax = X & 31, ay = Y & 31;
    // BilinearTab_i: exact products except the (0, 0) entry, where 2^15 saturates and the correction lands on the last tap
    int w0 = (32 - ay) * (32 - ax) * 32, w1 = (32 - ay) * ax * 32, w2 = ay * (32 - ax) * 32, w3 = ay * ax * 32;
    if ((ax | ay) == 0) { w0 = 32767; w3 = 1; }
    const int sx0 = cv_clip(sx, 0, W), sx1 = cv_clip(sx + 1, 0, W), sy0 = cv_clip(sy, 0, H), sy1 = cv_clip(sy + 1, 0, H);
    const unsigned char *p00 = src + ((size_t)sy0 * W + sx0) * 3, *p01 = src + ((size_t)sy0 * W + sx1) * 3;
    const unsigned char *p10 = src + ((size_t)sy1 * W + sx0) * 3, *p11 = src + ((size_t)sy1 * W + sx1) * 3;
    unsigned char *d = dst + ((size_t)y * W + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int t = p00[c] * w0 + p01[c] * w1 + p10[c] * w2 + p11[c] * w3;
        d[c] = sat_u8((t + (1 << 14)) >> 15);
    }
}

// ---------------------------------------------------------------------------------------------------- warpAffine, bicubic
struct Aff6 {
    double m[6];  // inverse map (dst -> src), already inverted like cv::warpAffine does
};

__global__ void __launch_bounds__(256) warp_affine_cubic_u8_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, int H,
                                                                    int W, Aff6 A, const short *__restrict__ tab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    // adelta[x] = cvRound(M0 * x * 1024); X0 = cvRound((M1 * y + M2) * 1024) + 16   (AB_BITS = 10, INTER_BITS = 5)
    const int adelta = cv_round_clamped(__dmul_rn(__dmul_rn(A.m[0], (double)x), 1024.0));
    const int bdelta = cv_round_clamped(__dmul_rn(__dmul_rn(A.m[3], (double)x), 1024.0));
    const int X0 = cv_round_clamped(__dmul_rn(__dadd_rn(__dmul_rn(A.m[1], (double)y), A.m[2]), 1024.0)) + 16;
    const int Y0 = cv_round_clamped(__dmul_rn(__dadd_rn(__dmul_rn(A.m[4], (double)y), A.m[5]), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = sat_short(X >> 5) - 1, sy = sat_short(Y >> 5) - 1;
    const short *w = tab + (((Y & 31) * 32) + (X & 31)) * 16;
    int t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        const unsigned char *row = src + (size_t)cv_clip(sy + k1, 0, H) * W * 3;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            const unsigned char *p = row + cv_clip(sx + k2, 0, W) * 3;
            const int wk = w[k1 * 4 + k2];
            t0 += p[0] * wk;
            t1 += p[1] * wk;
            t2 += p[2] * wk;
        }
    }
    unsigned char *d = dst + ((size_t)y * W + x) * 3;
    d[0] = sat_u8((t0 + (1 << 14)) >> 15);
    d[1] = sat_u8((t1 + (1 << 14)) >> 15);
    d[2] = sat_u8((t2 + (1 << 14)) >> 15);
}

// ---------------------------------------------------------------------------------------------------- crop + resize
struct CropArgs {
    int H, W, x0, y0, n, S;  // frame size, window origin (frame coordinates, may be outside), window side, output side
    unsigned char fill[3];   // channel means cast to uint8 (what the reference's padded-frame assignment stores)
    int gray;                // 1: write get_search_info's normalised gray [S, S] instead of the 3-channel crop
    double mean[3], std[3];
};

__device__ __forceinline__ int win_px(const unsigned char *__restrict__ f, const CropArgs &a, int wy, int wx, int c) {
    const int y = a.y0 + wy, x = a.x0 + wx;
    return (y >= 0 && y < a.H && x >= 0 && x < a.W) ? f[((size_t)y * a.W + x) * 3 + c] : a.fill[c];
}

// resize.cpp, linear: position and weights of destination index d (float32 steps exactly as OpenCV takes them)
__device__ __forceinline__ void lin_coeff(int d, double scale, int &s, float &f) {
    f = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
    s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
}

__global__ void __launch_bounds__(256) crop_resize_u8_kernel(const unsigned char *__restrict__ frame, float *__restrict__ out, CropArgs a, double scale) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= a.S) return;
    int v[3];
    if (a.n == a.S) {  // base_tracker.py:117: no resize when the window already has the model size
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = win_px(frame, a, dy, dx, c);
    } else if (a.n == 2 * a.S) {  // exact 2x decimation: INTER_LINEAR is routed to INTER_AREA (2x2 box mean)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            v[c] = (win_px(frame, a, 2 * dy, 2 * dx, c) + win_px(frame, a, 2 * dy, 2 * dx + 1, c) + win_px(frame, a, 2 * dy + 1, 2 * dx, c) +
                    win_px(frame, a, 2 * dy + 1, 2 * dx + 1, c) + 2) >> 2;
    } else {
        int sx, sy;
        float fx, fy;
        lin_coeff(dx, scale, sx, fx);
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= a.n - 1) { fx = 0.f; sx = a.n - 1; }
        lin_coeff(dy, scale, sy, fy);
        const int a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
        const int b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = __float2int_rn(__fmul_rn(fy, 2048.f));
        const int sx1 = min(sx + 1, a.n - 1), y0 = cv_clip(sy, 0, a.n), y1 = cv_clip(sy + 1, 0, a.n);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = win_px(frame, a, y0, sx, c) * a0 + win_px(frame, a, y0, sx1, c) * a1;
            const int h1 = win_px(frame, a, y1, sx, c) * a0 + win_px(frame, a, y1, sx1, c) * a1;
            v[c] = sat_u8((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
        }
    }
    if (!a.gray) {
#pragma unroll
        for (int c = 0; c < 3; ++c) out[((size_t)c * a.S + dy) * a.S + dx] = (float)v[c];
    } else {  // ((v - mean) / std) per channel in float64, np.mean over the 3 channels = ((c0 + c1) + c2) / 3, then float32
        const double g0 = __ddiv_rn(__dadd_rn((double)v[0], -a.mean[0]), a.std[0]), g1 = __ddiv_rn(__dadd_rn((double)v[1], -a.mean[1]), a.std[1]);
        const double g2 = __ddiv_rn(__dadd_rn((double)v[2], -a.mean[2]), a.std[2]);
        out[(size_t)dy * a.S + dx] = (float)__ddiv_rn(__dadd_rn(__dadd_rn(g0, g1), g2), 3.0);
    }
}

// BicubicTab_i of initInterTab2D (A = -0.75): float32 1-D taps, 15-bit products, each 4x4 kernel forced to sum to 2^15
static void cubic_table_host(short *tab) {
    const float A = -0.75f;
    float t1[32][4];
    for (int i = 0; i < 32; ++i) {
        const float x = i * (1.f / 32);
        t1[i][0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        t1[i][1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        t1[i][2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        t1[i][3] = 1.f - t1[i][0] - t1[i][1] - t1[i][2];
    }
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            short *it = tab + (i * 32 + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < 4; ++k1)
                for (int k2 = 0; k2 < 4; ++k2) {
                    const float v = t1[i][k1] * t1[j][k2];
                    long r = lrintf(v * 32768.f);
                    r = r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
                    it[k1 * 4 + k2] = (short)r;
                    isum += (int)r;
                }
            if (isum != 32768) {
                const int diff = isum - 32768;
                int big = 10, small = 10;
                for (int k1 = 2; k1 < 4; ++k1)
                    for (int k2 = 2; k2 < 4; ++k2) {
                        const int k = k1 * 4 + k2;
                        if (it[k] < it[small]) small = k;
                        else if (it[k] > it[big]) big = k;
                    }
                if (diff < 0) it[big] = (short)(it[big] - diff);
                else it[small] = (short)(it[small] - diff);
            }
        }
}

}  // namespace hdn

using namespace hdn;

extern "C" int hdn_warp_perspective_u8(const uint8_t *src, uint8_t *dst, int H, int W, const double *Minv_host, hdn_stream_t stream) {
    if (!src || !dst || !Minv_host) return HDN_ERR_NULL;
    if (H < 1 || W < 1 || H > 65535) return HDN_ERR_SHAPE;
    if (src == dst) return HDN_ERR_UNSUPPORTED;
    M9 M;
    for (int i = 0; i < 9; ++i) M.m[i] = Minv_host[i];
    const int bh = H < 16 ? H : 16;
    int bw = 1024 / bh;
    if (bw > W) bw = W;
    warp_perspective_u8_kernel<<<dim3((W + 255) / 256, H), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, bw, M);
    count_launch();
    return launch_status();
}

extern "C" int hdn_cubic_table_host(int16_t *tab) {
    if (!tab) return HDN_ERR_NULL;
    cubic_table_host(tab);
    return HDN_OK;
}

extern "C" int hdn_warp_affine_cubic_u8(const uint8_t *src, uint8_t *dst, int H, int W, const double *M_host, const int16_t *tab_dev,
                                        hdn_stream_t stream) {
    if (!src || !dst || !M_host || !tab_dev) return HDN_ERR_NULL;
    if (H < 1 || W < 1 || H > 65535) return HDN_ERR_SHAPE;
    if (src == dst) return HDN_ERR_UNSUPPORTED;
    // cv::warpAffine inverts the forward 2x3 matrix like this (imgwarp.cpp), in double, in this operation order
    const double *M = M_host;
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    Aff6 A;
    const double A11 = M[4] * D, A22 = M[0] * D;
    A.m[0] = A11;
    A.m[1] = M[1] * (-D);
    A.m[3] = M[3] * (-D);
    A.m[4] = A22;
    A.m[2] = -A.m[0] * M[2] - A.m[1] * M[5];
    A.m[5] = -A.m[3] * M[2] - A.m[4] * M[5];
    warp_affine_cubic_u8_kernel<<<dim3((W + 255) / 256, H), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, A, tab_dev);
    count_launch();
    return launch_status();
}

extern "C" int hdn_crop_resize_u8(const uint8_t *frame, int H, int W, int x0, int y0, int n, const uint8_t *fill_host, int S, int gray,
                                  const double *mean_host, const double *std_host, float *out, hdn_stream_t stream) {
    if (!frame || !fill_host || !out) return HDN_ERR_NULL;
    if (H < 1 || W < 1 || n < 1 || S < 1 || S > 65535) return HDN_ERR_SHAPE;
    if (gray && (!mean_host || !std_host)) return HDN_ERR_NULL;
    CropArgs a{};
    a.H = H; a.W = W; a.x0 = x0; a.y0 = y0; a.n = n; a.S = S; a.gray = gray;
    for (int c = 0; c < 3; ++c) {
        a.fill[c] = fill_host[c];
        a.mean[c] = gray ? mean_host[c] : 0.0;
        a.std[c] = gray ? std_host[c] : 1.0;
    }
    const double inv_scale = (double)S / n, scale = 1. / inv_scale;  // resize.cpp: inv_scale_x = dsize.width / ssize.width; scale_x = 1. / inv_scale_x
    crop_resize_u8_kernel<<<dim3((S + 255) / 256, S), 256, 0, (cudaStream_t)stream>>>(frame, out, a, scale);
    count_launch();
    return launch_status();
}
