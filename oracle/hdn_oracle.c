/*
 * hdn_oracle.c -- CPU restatement of the reference's hot-path operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (hdn_b200/) may link,
 * import or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and there only as the
 * checker.  Scalar single-thread C, fp32, compiled with -ffp-contract=off so
 * that every multiply/add below rounds exactly as written.
 *
 * Parity pin: every function here is checked against golden vectors produced
 * by running the UNMODIFIED reference (zhanxinrui/HDN @ 52cbb00) in this
 * container (oracle/gen_golden.py -> tests/golden/ops_*.npz; test in
 * tests/test_oracle_golden.py).  The reference ships no tests or fixtures of
 * its own for this path.
 *
 * Reference files are cited relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* --------------------------------------------------------------------------
 * K1 / K2: depth-wise cross-correlation.
 *   K1  hdn/core/xcorr.py:37-46  xcorr_depthwise:
 *         out[b,c,i,j] = sum_{u,v} x[b,c,i+u,j+v] * k[b,c,u,v]   (valid, no flip)
 *   K2  hdn/core/xcorr.py:48-61  xcorr_depthwise_circular:
 *         rows padded circularly by Hx/2 on both sides (:55), THEN columns padded
 *         by replicate with Wx/2 on both sides (:56; x.size(3) is still the
 *         unpadded width there), then K1.
 *         xp[i',j'] = x[(i' - Hx/2) mod Hx, clamp(j' - Wx/2, 0, Wx-1)]
 * k_bstride: element stride between batch items of k (0 = one template shared
 * by the whole batch, C*Hk*Wk = dense).
 * -------------------------------------------------------------------------- */
void orc_xcorr_dw(const float *x, const float *k, float *out, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular,
                  long long k_bstride)
{
    const int ph = circular ? Hx / 2 : 0, pw = circular ? Wx / 2 : 0;
    const int Ho = Hx + 2 * ph - Hk + 1, Wo = Wx + 2 * pw - Wk + 1;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *xp = x + ((size_t)b * C + c) * Hx * Wx;
            const float *kp = k + (size_t)b * k_bstride + (size_t)c * Hk * Wk;
            float *op = out + ((size_t)b * C + c) * Ho * Wo;
            for (int i = 0; i < Ho; ++i)
                for (int j = 0; j < Wo; ++j) {
                    float acc = 0.f;
                    for (int u = 0; u < Hk; ++u) {
                        int r = i + u - ph;
                        if (circular) r = ((r % Hx) + Hx) % Hx;
                        for (int v = 0; v < Wk; ++v) {
                            int cc = j + v - pw;
                            if (cc < 0) cc = 0;
                            if (cc > Wx - 1) cc = Wx - 1;
                            acc += xp[r * Wx + cc] * kp[u * Wk + v];
                        }
                    }
                    op[i * Wo + j] = acc;
                }
        }
}

/* --------------------------------------------------------------------------
 * K3: log-polar resampling.  hdn/models/logpolar.py:50-134 (STN_Polar).
 *   S = INSTANCE_SIZE//2 (:56);  mag = ln(S/2)/S (:63);
 *   rho_j = exp(mag*j) - 1 (:65);  theta_i = i*2*pi/S + delta_rot (:66);
 *   rows = angle, cols = log-radius (meshgrid([theta, rho]) :67);
 *   gx = (rho*cos(theta) + polar_x) / (H//2),  gy = (rho*sin(theta) + polar_y) / (W//2)  (:113-117;
 *   the reference really divides x by size(2) and y by size(3));
 *   F.grid_sample(bilinear, padding_mode='border', align_corners=False) (:124).
 * -------------------------------------------------------------------------- */
static float clipf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

void orc_logpolar(const float *img, const float *polar, float delta_rot, float *out, int B, int Ch, int H, int W, int S)
{
    const float mag = (float)(log((double)S / 2.0) / (double)S);
    const float pi_f = (float)3.14159265358979323846;
    for (int b = 0; b < B; ++b) {
        const float px = polar ? polar[2 * b] : 0.f, py = polar ? polar[2 * b + 1] : 0.f;
        for (int i = 0; i < S; ++i) {
            const float theta = ((float)i * 2.0f) * pi_f / (float)S + delta_rot;
            const float ct = cosf(theta), st = sinf(theta);
            for (int j = 0; j < S; ++j) {
                const float rho = expf(mag * (float)j) - 1.0f;
                const float gx = (rho * ct + px) / (float)(H / 2);
                const float gy = (rho * st + py) / (float)(W / 2);
                /* grid_sampler unnormalize, align_corners=False, then border clip */
                float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
                float iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
                ix = clipf(ix, 0.f, (float)(W - 1));
                iy = clipf(iy, 0.f, (float)(H - 1));
                const float fx = floorf(ix), fy = floorf(iy);
                const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
                const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
                const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
                for (int c = 0; c < Ch; ++c) {
                    const float *ip = img + ((size_t)b * Ch + c) * H * W;
                    float v = 0.f;
                    v += ip[y0 * W + x0] * (wx0 * wy0);
                    if (x1 <= W - 1) v += ip[y0 * W + x1] * (wx1 * wy0);
                    if (y1 <= H - 1) v += ip[y1 * W + x0] * (wx0 * wy1);
                    if (x1 <= W - 1 && y1 <= H - 1) v += ip[y1 * W + x1] * (wx1 * wy1);
                    out[(((size_t)b * Ch + c) * S + i) * S + j] = v;
                }
            }
        }
    }
}

/* --------------------------------------------------------------------------
 * K5: 4-point DLT.  homo_estimator/Deep_homography/Oneline_DLTv1/utils.py:7-67.
 *   src, off: [B,8] = (x0,y0,x1,y1,x2,y2,x3,y3).  For 8-vectors divide=1 and the
 *   index list (:18-26) re-orders the points to [p0,p1,p3,p2]; dst = src + off (:42).
 *   Per point (x,y)->(u,v):  [x y 1 0 0 0 -u*x -u*y | u],  [0 0 0 x y 1 -v*x -v*y | v]  (:52-60),
 *   h8 = inverse(A) b (:62-63), H = [h8,1] (:65).
 * A and b are formed in fp32 exactly as the reference does; the 8x8 system is
 * then solved in double with partial pivoting (the reference's fp32 LU inverse
 * differs from this by its own rounding only; tolerance in the tests).
 * -------------------------------------------------------------------------- */
void orc_dlt4(const float *src, const float *off, float *Hout, int B)
{
    static const int order[4] = {0, 1, 3, 2};
    for (int b = 0; b < B; ++b) {
        double A[8][9];
        for (int p = 0; p < 4; ++p) {
            const int q = order[p];
            const float x = src[b * 8 + 2 * q], y = src[b * 8 + 2 * q + 1];
            const float u = x + off[b * 8 + 2 * q], v = y + off[b * 8 + 2 * q + 1];
            const float ux = u * x, uy = u * y, vx = v * x, vy = v * y;
            double *r0 = A[2 * p], *r1 = A[2 * p + 1];
            r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -ux; r0[7] = -uy; r0[8] = u;
            r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -vx; r1[7] = -vy; r1[8] = v;
        }
        for (int col = 0; col < 8; ++col) {
            int piv = col;
            for (int r = col + 1; r < 8; ++r)
                if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
            if (piv != col)
                for (int c = 0; c < 9; ++c) { double t = A[col][c]; A[col][c] = A[piv][c]; A[piv][c] = t; }
            const double inv = 1.0 / A[col][col];
            for (int r = 0; r < 8; ++r) {
                if (r == col) continue;
                const double f = A[r][col] * inv;
                for (int c = col; c < 9; ++c) A[r][c] -= f * A[col][c];
            }
        }
        for (int r = 0; r < 8; ++r) Hout[b * 9 + r] = (float)(A[r][8] / A[r][r]);
        Hout[b * 9 + 8] = 1.0f;
    }
}

/* --------------------------------------------------------------------------
 * K4: projective bilinear warp.  Oneline_DLTv1/utils.py:257-274 transform ->
 * :70-254 transformer (_meshgrid :192, _transform :215, _interpolate :114).
 *   theta = (Minv @ H) @ M (:262), M = [[hw,0,hw],[0,hh,hh],[0,0,1]]
 *   (hw = hh = 63.5 hard-coded by the caller, model_builder...py:196-199);
 *   target grid x_t = linspace(-1,1,W)[j], y_t = linspace(-1,1,H)[i] (:195-198);
 *   (xs,ys,ts) = theta (x_t,y_t,1);  ts += 1e-6 where |ts| < 1e-7 (:235-238);
 *   x = (xs/ts + 1)*W/2, y = (ys/ts + 1)*H/2 (:127-128);
 *   x0 = floor(x), x1 = x0+1, both clamped to [0,W-1] (:131-139); weights are
 *   taken from the CLAMPED corners (:179-187), so out-of-range samples give 0
 *   and there is no border replicate;  out = wa*Ia + wb*Ib + wc*Ic + wd*Id (:188).
 *   The trailing gather by patch_indices (:268-273) is the identity permutation
 *   for the caller's arange(H*W) indices (get_img_info.py:92) and is omitted.
 * Minv is passed in (the caller computes torch.inverse(M)); Hm is [B,9].
 * torch.linspace(-1,1,n) in fp32: step = 2/(n-1); element i is start + step*i
 * for i < n/2 and end - step*(n-1-i) otherwise (ATen RangeFactories.cpp).
 * -------------------------------------------------------------------------- */
static float linspace_m1_p1(int i, int n)
{
    const float step = (1.0f - (-1.0f)) / (float)(n - 1);
    const int halfway = n / 2;
    return i < halfway ? -1.0f + step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

static void mat3_mul(const float *a, const float *b, float *c)
{
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) {
            float s = a[r * 3 + 0] * b[0 * 3 + q];
            s += a[r * 3 + 1] * b[1 * 3 + q];
            s += a[r * 3 + 2] * b[2 * 3 + q];
            c[r * 3 + q] = s;
        }
}

void orc_homo_warp(const float *img, const float *Hm, const float *M, const float *Minv, float *out, int B, int Ch, int H, int W)
{
    for (int b = 0; b < B; ++b) {
        float t0[9], th[9];
        mat3_mul(Minv, Hm + 9 * b, t0);
        mat3_mul(t0, M, th);
        for (int i = 0; i < H; ++i) {
            const float yt = linspace_m1_p1(i, H);
            for (int j = 0; j < W; ++j) {
                const float xt = linspace_m1_p1(j, W);
                float xs = th[0] * xt; xs += th[1] * yt; xs += th[2];
                float ys = th[3] * xt; ys += th[4] * yt; ys += th[5];
                float ts = th[6] * xt; ts += th[7] * yt; ts += th[8];
                if (!(fabsf(ts) >= 1e-7f)) ts = ts + 1e-6f;
                float x = (xs / ts + 1.0f) * (float)W / 2.0f;
                float y = (ys / ts + 1.0f) * (float)H / 2.0f;
                /* floor(...).int() then clamp; saturate first so the int cast is defined */
                float fx = floorf(x), fy = floorf(y);
                fx = clipf(fx, -4.0f, (float)W + 4.0f);
                fy = clipf(fy, -4.0f, (float)H + 4.0f);
                if (fx != fx) fx = 0.f;
                if (fy != fy) fy = 0.f;
                int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
                x0 = x0 < 0 ? 0 : (x0 > W - 1 ? W - 1 : x0);
                x1 = x1 < 0 ? 0 : (x1 > W - 1 ? W - 1 : x1);
                y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
                y1 = y1 < 0 ? 0 : (y1 > H - 1 ? H - 1 : y1);
                const float wa = ((float)x1 - x) * ((float)y1 - y);
                const float wb = ((float)x1 - x) * (y - (float)y0);
                const float wc = (x - (float)x0) * ((float)y1 - y);
                const float wd = (x - (float)x0) * (y - (float)y0);
                for (int c = 0; c < Ch; ++c) {
                    const float *ip = img + ((size_t)b * Ch + c) * H * W;
                    float v = wa * ip[y0 * W + x0];
                    v += wb * ip[y1 * W + x0];
                    v += wc * ip[y0 * W + x1];
                    v += wd * ip[y1 * W + x1];
                    out[(((size_t)b * Ch + c) * H + i) * W + j] = v;
                }
            }
        }
    }
}

/* --------------------------------------------------------------------------
 * K6: score conversion, window blend, arg-max, offset gather.
 *   hdn/tracker/hdn_tracker.py:82-89 _convert_score: 2-way softmax, p(fg) (fp32).
 *   hdn/tracker/hdn_tracker_proj_e2e.py:172-174:
 *       pscore = score*(1-w) + window*w   -- score is float32, (1-w) a Python
 *       float (product stays float32 under NumPy scalar rules), window float64,
 *       so the sum and the arg-max are in float64;  np.argmax = first maximum.
 *   lp branch (:199-206): no window (window == NULL), arg-max of the fp32 score.
 *   loc gather: loc[b,:,idx] (base_tracker.py:54-59 / hdn_tracker.py:51-67 then
 *   only read column idx).
 * Outputs: idx[B] int64, pscore[B] float64 (value the <0.05 / <0.25 gates test),
 * score[B] float32 ('best_score'), gathered[B,L] float32.
 * -------------------------------------------------------------------------- */
void orc_score_argmax(const float *cls, const float *loc, const double *window, double win_influence, int64_t *idx, double *pscore,
                      float *score, float *gathered, int B, int L, int N)
{
    const int n = N * N;
    const float one_minus_w = (float)(1.0 - win_influence);
    for (int b = 0; b < B; ++b) {
        const float *c0 = cls + (size_t)b * 2 * n, *c1 = c0 + n;
        int best = 0;
        double bestv = -INFINITY;
        float bests = 0.f;
        for (int p = 0; p < n; ++p) {
            const float m = c0[p] > c1[p] ? c0[p] : c1[p];
            const float e0 = expf(c0[p] - m), e1 = expf(c1[p] - m);
            const float s = e1 / (e0 + e1);
            double ps;
            if (window) {
                const float t = s * one_minus_w;
                ps = (double)t + window[p] * win_influence;
            } else {
                ps = (double)s;
            }
            /* np.argmax: first maximum; a NaN counts as the maximum and the first NaN wins */
            if ((ps > bestv && bestv == bestv) || (ps != ps && bestv == bestv)) { bestv = ps; best = p; bests = s; }
        }
        idx[b] = best;
        pscore[b] = bestv;
        score[b] = bests;
        for (int l = 0; l < L; ++l) gathered[b * L + l] = loc[((size_t)b * L + l) * n + best];
    }
}

int orc_abi_version(void) { return 1; }
