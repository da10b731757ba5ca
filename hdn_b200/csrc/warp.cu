// warp.cu -- K3 (log-polar resampling), K4 (projective bilinear warp), K5 (4-point DLT) and the
// fused K5+K4 launch, for sm_100a.
//
// K3 replaces hdn/models/logpolar.py:50-134 (STN_Polar: grid built on the host with
//    linspace/exp/cos/sin/meshgrid, uploaded every call, then F.grid_sample).  Here the grid is
//    analytic inside the kernel: nothing is uploaded, one launch.
// K4 replaces Oneline_DLTv1/utils.py:257-274 transform -> :70-254 transformer (~40 ATen launches).
// K5 replaces Oneline_DLTv1/utils.py:7-67 DLT_solve (torch.inverse of an 8x8 + ~15 cat/reshape).
//
// All coordinate arithmetic uses the explicit round-to-nearest intrinsics (__fmul_rn, ...) in the
// reference's operation order so the compiler cannot contract it into FMAs: the warp has jump
// discontinuities at the image border (weights come from clamped corners) and the arg-max path is
// bit-exact, so sample coordinates must not depend on contraction choices.
#include <cmath>
#include "common.cuh"

namespace hdn {

// ------------------------------------------------------------------------------------------ K3
// out[b,ch,i,j]: i = angle index, j = log-radius index.  One thread per (batch chunk, i, j): the sampling position depends on
// the batch item only through `polar`, so without it (the tracker always passes zeros, hdn_tracker_proj_e2e.py:194) the
// transcendental work, the IEEE divisions and the four neighbour offsets are computed once and shared by the BC items of the
// chunk and all channels -- and the BC * Ch * 4 gathers of a thread are independent loads in flight.
struct LpSample {
    int o00;
    bool xin, yin;
    float w00, w10, w01, w11;
};

__device__ __forceinline__ LpSample logpolar_sample_pos(int i, int j, float px, float py, float rot_delta, float mag, int H, int W, int S) {
    const float pi_f = 3.14159265358979323846f;
    const float fS = (float)S, fW = (float)W, fH = (float)H;
    const float half_h = (float)(H / 2), half_w = (float)(W / 2);
    // theta = i*2*pi/S + delta  (logpolar.py:66, evaluated left to right in fp32)
    const float theta = __fadd_rn(__fdiv_rn(__fmul_rn(__fmul_rn((float)i, 2.0f), pi_f), fS), rot_delta);
    float st, ct;
    sincosf(theta, &st, &ct);
    const float rho = __fsub_rn(expf(__fmul_rn(mag, (float)j)), 1.0f);  // :65
    // normalised grid (:113-117): x over size(2)//2, y over size(3)//2
    const float gx = __fdiv_rn(__fadd_rn(__fmul_rn(rho, ct), px), half_h);
    const float gy = __fdiv_rn(__fadd_rn(__fmul_rn(rho, st), py), half_w);
    // grid_sample: unnormalise (align_corners=False), clip to the border, bilinear
    float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), fW), 1.f), 2.f);
    float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), fH), 1.f), 2.f);
    ix = fminf(fmaxf(ix, 0.f), fW - 1.f);
    iy = fminf(fmaxf(iy, 0.f), fH - 1.f);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    LpSample s;
    s.xin = x0 + 1 <= W - 1;
    s.yin = y0 + 1 <= H - 1;
    const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.f), ix);
    const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.f), iy);
    s.w00 = __fmul_rn(wx0, wy0); s.w10 = __fmul_rn(wx1, wy0); s.w01 = __fmul_rn(wx0, wy1); s.w11 = __fmul_rn(wx1, wy1);
    s.o00 = y0 * W + x0;
    return s;
}

__device__ __forceinline__ float lp_px(const float *p) { return __ldg(p); }
__device__ __forceinline__ float lp_px(const unsigned char *p) { return (float)__ldg(p); }  // uint8 crops: exact in fp32

template <int BC, typename T>
__global__ void __launch_bounds__(256)
    logpolar_kernel(const T *__restrict__ img, const float *__restrict__ polar, float rot_delta, float *__restrict__ out, int B, int Ch,
                    int H, int W, int S, float mag) {
    const int chunks = (B + BC - 1) / BC;
    const long long total = (long long)chunks * S * S;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e % S);
        const int i = (int)((e / S) % S);
        const int b0 = (int)(e / ((long long)S * S)) * BC;
        LpSample s = logpolar_sample_pos(i, j, 0.f, 0.f, rot_delta, mag, H, W, S);
#pragma unroll
        for (int bb = 0; bb < BC; ++bb) {
            const int b = b0 + bb;
            if (b >= B) break;
            if (polar) s = logpolar_sample_pos(i, j, __ldg(polar + 2 * b), __ldg(polar + 2 * b + 1), rot_delta, mag, H, W, S);
            const T *ip = img + (long long)b * Ch * H * W + s.o00;
            float *op = out + ((long long)b * Ch * S + i) * S + j;
            for (int c = 0; c < Ch; ++c) {
                float v = __fmul_rn(lp_px(ip), s.w00);
                if (s.xin) v = __fadd_rn(v, __fmul_rn(lp_px(ip + 1), s.w10));
                if (s.yin) v = __fadd_rn(v, __fmul_rn(lp_px(ip + W), s.w01));
                if (s.xin && s.yin) v = __fadd_rn(v, __fmul_rn(lp_px(ip + W + 1), s.w11));
                *op = v;
                ip += (long long)H * W;
                op += (long long)S * S;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ K5
// 8 lanes own one 8x9 augmented system (lane r = row r), Gauss-Jordan with partial pivoting in
// fp64 through warp shuffles.  A and b are FORMED in fp32 exactly as the reference forms them
// (utils.py:42-60), so this solves the reference's own matrix; only its LU rounding differs.
struct Mat9 {
    float v[9];
};

__device__ __forceinline__ void dlt_build_row(const float *__restrict__ src, const float *__restrict__ off, int r, double (&a)[9]) {
    const int order[4] = {0, 1, 3, 2};  // utils.py:18-26 re-orders the quad to [p0,p1,p3,p2]
    const int q = order[r >> 1];
    const float x = src[2 * q], y = src[2 * q + 1];
    const float u = __fadd_rn(x, off[2 * q]), v = __fadd_rn(y, off[2 * q + 1]);
    if ((r & 1) == 0) {
        a[0] = x; a[1] = y; a[2] = 1.0; a[3] = 0.0; a[4] = 0.0; a[5] = 0.0;
        a[6] = -(double)__fmul_rn(u, x); a[7] = -(double)__fmul_rn(u, y); a[8] = u;
    } else {
        a[0] = 0.0; a[1] = 0.0; a[2] = 0.0; a[3] = x; a[4] = y; a[5] = 1.0;
        a[6] = -(double)__fmul_rn(v, x); a[7] = -(double)__fmul_rn(v, y); a[8] = v;
    }
}

// Called by 8 consecutive lanes (r = lane & 7) with their row in a[]; returns h[r] = solution entry r.
__device__ __forceinline__ double dlt_solve8(double (&a)[9], int r, unsigned mask) {
#pragma unroll
    for (int col = 0; col < 8; ++col) {
        // pivot = first row >= col with the largest |a[col]|
        double best = (r >= col) ? fabs(a[col]) : -1.0;
        int bi = r;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const double ob = __shfl_xor_sync(mask, best, d, 8);
            const int oi = __shfl_xor_sync(mask, bi, d, 8);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const int piv = bi;
        double prow[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const double vp = __shfl_sync(mask, a[c], piv, 8);
            const double vc = __shfl_sync(mask, a[c], col, 8);
            if (r == col) a[c] = vp;
            else if (r == piv) a[c] = vc;
            prow[c] = vp;  // the pivot row now sits in lane col
        }
        const double inv = __drcp_rn(prow[col]);
        if (r != col) {
            const double f = __dmul_rn(a[col], inv);
#pragma unroll
            for (int c = 0; c < 9; ++c)
                if (c >= col) a[c] = __dsub_rn(a[c], __dmul_rn(f, prow[c]));
        }
    }
    double diag = a[0];
#pragma unroll
    for (int c = 1; c < 8; ++c)
        if (r == c) diag = a[c];
    return __ddiv_rn(a[8], diag);
}

__global__ void __launch_bounds__(128) dlt4_kernel(const float *__restrict__ src, const float *__restrict__ off, float *__restrict__ Hm, int B) {
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // system index
    const int r = threadIdx.x & 7;
    const bool live = gid < B;
    const int b = live ? gid : B - 1;  // idle groups redo the last system (shuffles need full participation)
    double a[9];
    dlt_build_row(src + 8 * b, off + 8 * b, r, a);
    const double h = dlt_solve8(a, r, 0xffffffffu);
    if (live) {
        Hm[9 * b + r] = (float)h;
        if (r == 0) Hm[9 * b + 8] = 1.0f;
    }
}

// ------------------------------------------------------------------------------------------ K4
__device__ __forceinline__ float linspace_m1_p1(int i, int n) {  // torch.linspace(-1, 1, n)[i] in fp32
    const float step = __fdiv_rn(2.0f, (float)(n - 1));
    return i < n / 2 ? __fadd_rn(-1.0f, __fmul_rn(step, (float)i)) : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
}

__device__ __forceinline__ float dot3_rn(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// theta = (Minv @ Hm) @ M, element e of 9 (utils.py:262)
__device__ __forceinline__ float theta_elem(const float *__restrict__ Hm, const Mat9 &M, const Mat9 &Mi, int e) {
    const int r = e / 3, q = e % 3;
    float t0[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) t0[c] = dot3_rn(Mi.v[r * 3 + 0], Hm[0 * 3 + c], Mi.v[r * 3 + 1], Hm[1 * 3 + c], Mi.v[r * 3 + 2], Hm[2 * 3 + c]);
    return dot3_rn(t0[0], M.v[0 * 3 + q], t0[1], M.v[1 * 3 + q], t0[2], M.v[2 * 3 + q]);
}

__device__ __forceinline__ void warp_pixel(const float *__restrict__ img, float *__restrict__ out, const float *th, int b, int pix, int Ch,
                                           int H, int W) {
    const int i = pix / W, j = pix - i * W;
    const float xt = linspace_m1_p1(j, W), yt = linspace_m1_p1(i, H);
    const float xs = __fadd_rn(__fadd_rn(__fmul_rn(th[0], xt), __fmul_rn(th[1], yt)), th[2]);
    const float ys = __fadd_rn(__fadd_rn(__fmul_rn(th[3], xt), __fmul_rn(th[4], yt)), th[5]);
    float ts = __fadd_rn(__fadd_rn(__fmul_rn(th[6], xt), __fmul_rn(th[7], yt)), th[8]);
    if (!(fabsf(ts) >= 1e-7f)) ts = __fadd_rn(ts, 1e-6f);  // utils.py:235-238
    const float x = __fdiv_rn(__fmul_rn(__fadd_rn(__fdiv_rn(xs, ts), 1.0f), (float)W), 2.0f);  // :127
    const float y = __fdiv_rn(__fmul_rn(__fadd_rn(__fdiv_rn(ys, ts), 1.0f), (float)H), 2.0f);  // :128
    float fx = floorf(x), fy = floorf(y);
    fx = fminf(fmaxf(fx, -4.0f), (float)W + 4.0f);  // saturate before the int cast (NaN -> -4 -> clamps to 0)
    fy = fminf(fmaxf(fy, -4.0f), (float)H + 4.0f);
    if (fx != fx) fx = 0.f;
    if (fy != fy) fy = 0.f;
    int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);  // :136-139
    y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
    // weights from the CLAMPED corners (:179-187): out-of-range samples cancel to ~0, no border replicate
    const float dx1 = __fsub_rn((float)x1, x), dx0 = __fsub_rn(x, (float)x0);
    const float dy1 = __fsub_rn((float)y1, y), dy0 = __fsub_rn(y, (float)y0);
    const float wa = __fmul_rn(dx1, dy1), wb = __fmul_rn(dx1, dy0), wc = __fmul_rn(dx0, dy1), wd = __fmul_rn(dx0, dy0);
    const float *ip = img + (long long)b * Ch * H * W;
    float *op = out + (long long)b * Ch * H * W + pix;
    for (int c = 0; c < Ch; ++c) {
        float v = __fmul_rn(wa, __ldg(ip + y0 * W + x0));
        v = __fadd_rn(v, __fmul_rn(wb, __ldg(ip + y1 * W + x0)));
        v = __fadd_rn(v, __fmul_rn(wc, __ldg(ip + y0 * W + x1)));
        v = __fadd_rn(v, __fmul_rn(wd, __ldg(ip + y1 * W + x1)));
        *op = v;
        ip += H * W;
        op += H * W;
    }
}

// grid = (tiles over H*W, B).  FUSED: the block first solves its item's DLT (8 lanes) and block 0 of
// each item publishes H; otherwise H is read from memory.
template <bool FUSED>
__global__ void __launch_bounds__(256)
    homo_warp_kernel(const float *__restrict__ src, const float *__restrict__ off, const float *__restrict__ img, float *__restrict__ Hm,
                     Mat9 M, Mat9 Mi, float *__restrict__ out, int Ch, int H, int W) {
    __shared__ float sH[9];
    __shared__ float sth[9];
    const int b = blockIdx.y;
    if (FUSED) {
        if (threadIdx.x < 32) {
            const int r = threadIdx.x & 7;
            double a[9];
            dlt_build_row(src + 8 * b, off + 8 * b, r, a);
            const double h = dlt_solve8(a, r, 0xffffffffu);
            if (threadIdx.x < 8) sH[r] = (float)h;
            if (threadIdx.x == 0) sH[8] = 1.0f;
        }
    } else {
        if (threadIdx.x < 9) sH[threadIdx.x] = Hm[9 * b + threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        sth[threadIdx.x] = theta_elem(sH, M, Mi, threadIdx.x);
        if (FUSED && blockIdx.x == 0) Hm[9 * b + threadIdx.x] = sH[threadIdx.x];
    }
    __syncthreads();
    float th[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) th[e] = sth[e];
    const int npix = H * W;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) warp_pixel(img, out, th, b, pix, Ch, H, W);
}

static void default_M(int W, int H, Mat9 &M, Mat9 &Mi) {
    const float a = (float)(W / 2.0), c = (float)(H / 2.0);
    const float m[9] = {a, 0, a, 0, c, c, 0, 0, 1};
    const float mi[9] = {1.0f / a, 0, -1.0f, 0, 1.0f / c, -1.0f, 0, 0, 1};
    for (int i = 0; i < 9; ++i) { M.v[i] = m[i]; Mi.v[i] = mi[i]; }
}

}  // namespace hdn

using namespace hdn;

template <typename T>
static int logpolar_launch(const T *img, const float *polar, float rot_delta, float *out, int B, int Ch, int H, int W, int S, hdn_stream_t stream) {
    if (!img || !out) return HDN_ERR_NULL;
    if (B < 1 || Ch < 1 || H < 2 || W < 2 || S < 2) return HDN_ERR_SHAPE;
    const float mag = (float)(log((double)S / 2.0) / (double)S);  // logpolar.py:63
    // batch items per thread: as many as still leave a few waves of CTAs
    const long long cap = (long long)sm_count() * 8;
    auto launch = [&](auto kern, int bc) {
        const long long total = (long long)((B + bc - 1) / bc) * S * S;
        long long blocks = (total + 255) / 256;
        if (blocks > cap) blocks = cap;
        kern<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, polar, rot_delta, out, B, Ch, H, W, S, mag);
    };
    if ((long long)B * S * S >= 4 * cap * 256) launch(logpolar_kernel<4, T>, 4);
    else launch(logpolar_kernel<1, T>, 1);
    count_launch();
    return launch_status();
}

extern "C" int hdn_logpolar_f32(const float *img, const float *polar, float rot_delta, float *out, int B, int Ch, int H, int W, int S,
                                hdn_stream_t stream) {
    return logpolar_launch(img, polar, rot_delta, out, B, Ch, H, W, S, stream);
}

extern "C" int hdn_logpolar_u8(const uint8_t *img, const float *polar, float rot_delta, float *out, int B, int Ch, int H, int W, int S,
                               hdn_stream_t stream) {
    return logpolar_launch(img, polar, rot_delta, out, B, Ch, H, W, S, stream);
}

extern "C" int hdn_dlt4_f32(const float *src4, const float *off4, float *Hm, int B, hdn_stream_t stream) {
    if (!src4 || !off4 || !Hm) return HDN_ERR_NULL;
    if (B < 1) return HDN_ERR_SHAPE;
    const int blocks = (B * 8 + 127) / 128;
    dlt4_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(src4, off4, Hm, B);
    count_launch();
    return launch_status();
}

static int warp_launch(bool fused, const float *src4, const float *off4, const float *img, const float *M_host, const float *Minv_host,
                       float *Hm, float *out, int B, int Ch, int H, int W, cudaStream_t st) {
    if (!img || !Hm || !out || (fused && (!src4 || !off4))) return HDN_ERR_NULL;
    if ((M_host == nullptr) != (Minv_host == nullptr)) return HDN_ERR_NULL;
    if (B < 1 || Ch < 1 || H < 2 || W < 2) return HDN_ERR_SHAPE;
    if (B > 65535) return HDN_ERR_UNSUPPORTED;
    Mat9 M, Mi;
    if (M_host) {
        for (int i = 0; i < 9; ++i) { M.v[i] = M_host[i]; Mi.v[i] = Minv_host[i]; }
    } else {
        default_M(W, H, M, Mi);
    }
    int tiles = (H * W + 255) / 256;
    // keep the whole launch around a few waves of CTAs
    const int cap = (sm_count() * 8 + B - 1) / B;
    if (tiles > cap) tiles = cap < 1 ? 1 : cap;
    dim3 grid(tiles, B);
    if (fused)
        homo_warp_kernel<true><<<grid, 256, 0, st>>>(src4, off4, img, Hm, M, Mi, out, Ch, H, W);
    else
        homo_warp_kernel<false><<<grid, 256, 0, st>>>(nullptr, nullptr, img, Hm, M, Mi, out, Ch, H, W);
    count_launch();
    return launch_status();
}

extern "C" int hdn_homo_warp_f32(const float *img, const float *Hm, const float *M_host, const float *Minv_host, float *out, int B, int Ch,
                                 int H, int W, hdn_stream_t stream) {
    return warp_launch(false, nullptr, nullptr, img, M_host, Minv_host, const_cast<float *>(Hm), out, B, Ch, H, W, (cudaStream_t)stream);
}

extern "C" int hdn_dlt_warp_f32(const float *src4, const float *off4, const float *img, const float *M_host, const float *Minv_host,
                                float *Hm, float *out, int B, int Ch, int H, int W, hdn_stream_t stream) {
    return warp_launch(true, src4, off4, img, M_host, Minv_host, Hm, out, B, Ch, H, W, (cudaStream_t)stream);
}
