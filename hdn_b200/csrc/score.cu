// score.cu -- K6: 2-way softmax -> window blend -> first-max arg-max -> offset gather, on device.
//
// Replaces the per-frame device->host->NumPy epilogue of the reference:
//   hdn/tracker/hdn_tracker.py:82-89        _convert_score (softmax over the 2 cls channels, p(fg))
//   hdn/tracker/hdn_tracker_proj_e2e.py:172-174  pscore = score*(1-w) + window*w ; np.argmax
//   hdn/tracker/base_tracker.py:54-59 / hdn_tracker.py:51-67  only column idx of loc is consumed
// so one 8-byte index, two scalars and L floats cross PCIe instead of the whole score/loc maps.
//
// Bit-exactness of the index: the reference multiplies the float32 score by the Python float (1-w)
// in float32, adds the float64 window term in float64 and takes the FIRST maximum.  The kernel
// does exactly that arithmetic (fp32 product, fp64 sum) and resolves ties towards the smaller index.
#include "common.cuh"

namespace hdn {

struct Best {
    double v;
    int i;
    float s;
};

// np.argmax semantics (the reference's arg-max, hdn_tracker_proj_e2e.py:174): the FIRST maximum, and a NaN counts as the maximum
// (NumPy propagates it: np.argmax([1, nan, 3]) == 1), so a frame with non-finite scores yields the index NumPy would, never an
// out-of-range one.
__device__ __forceinline__ Best better(const Best &a, const Best &b) {
    const bool an = a.v != a.v, bn = b.v != b.v;
    if (an || bn) return (an && bn) ? (b.i < a.i ? b : a) : (bn ? b : a);
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

__global__ void __launch_bounds__(256)
    score_argmax_kernel(const float *__restrict__ cls, const float *__restrict__ loc, const double *__restrict__ window, double w_infl,
                        float one_minus_w, long long *__restrict__ idx, double *__restrict__ pscore, float *__restrict__ score,
                        float *__restrict__ gathered, int L, int n) {
    const int b = blockIdx.x;
    const float *c0 = cls + (long long)b * 2 * n, *c1 = c0 + n;
    Best best{-INFINITY, 0x7fffffff, 0.f};
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const float a0 = __ldg(c0 + p), a1 = __ldg(c1 + p);
        const float m = fmaxf(a0, a1);
        const float e0 = expf(__fsub_rn(a0, m)), e1 = expf(__fsub_rn(a1, m));
        const float s = __fdiv_rn(e1, __fadd_rn(e0, e1));
        double ps;
        if (window) ps = __dadd_rn((double)__fmul_rn(s, one_minus_w), __dmul_rn(__ldg(window + p), w_infl));
        else ps = (double)s;
        const Best cand{ps, p, s};
        best = better(best, cand);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Best o;
        o.v = __shfl_xor_sync(0xffffffffu, best.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
        o.s = __shfl_xor_sync(0xffffffffu, best.s, off);
        best = better(best, o);
    }
    __shared__ Best sb[8];
    __shared__ int s_idx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sb[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        Best r = sb[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = better(r, sb[w]);
        idx[b] = min(max(r.i, 0), n - 1);
        pscore[b] = r.v;
        score[b] = r.s;
        s_idx = min(max(r.i, 0), n - 1);  // defensive: the gather below must stay inside loc whatever the scores were
    }
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x) gathered[b * L + l] = __ldg(loc + ((long long)b * L + l) * n + s_idx);
}

// ---- fused head epilogue: level-weighted combination of the per-level partial maps + the arg-max above ---------------------------
// MultiBAN.forward (hdn/models/head/ban.py:102-127) ends with
//     cls = sum_l softmax(cls_weight)[l] * cls_l          loc = sum_l softmax(loc_weight)[l] * (loc_l * loc_scale[l])
// where cls_l / loc_l = second 1x1 convolution (+ bias) of level l.  hdn_head_project_multi_f32 leaves that convolution as `ntile`
// partial sums per level (one per 128-channel tile, no bias); this kernel adds them in a fixed order, applies bias, loc_scale and
// the level weights with the reference's operation order (separately rounded multiply / add) and runs K6 on the result, so the
// per-level maps never exist in HBM and only the combined [2 | L, N, N] maps (optional) and the arg-max leave the kernel.
struct HeadLevels {
    const float *cls[4], *loc[4], *cls_bias[4], *loc_bias[4];
    float cls_w[4], loc_scale[4], loc_w[4];
    int nlev, ntile;
};

__device__ __forceinline__ float head_combine(const float *const (&parts)[4], const float *const (&bias)[4], const float (&w)[4], const float *scale,
                                              int nlev, int ntile, int B, int b, int nch, int ch, int n, int p) {
    float acc = 0.f;
    for (int l = 0; l < nlev; ++l) {
        float v = 0.f;
        for (int t = 0; t < ntile; ++t) v = __fadd_rn(v, __ldg(parts[l] + (((size_t)t * B + b) * nch + ch) * n + p));
        v = __fadd_rn(v, __ldg(bias[l] + ch));
        if (scale) v = __fmul_rn(v, scale[l]);
        acc = __fadd_rn(acc, __fmul_rn(v, w[l]));
    }
    return acc;
}

__global__ void __launch_bounds__(1024)
    head_score_kernel(const __grid_constant__ HeadLevels h, float *__restrict__ cls_out, float *__restrict__ loc_out, const double *__restrict__ window,
                      double w_infl, float one_minus_w, long long *__restrict__ idx, double *__restrict__ pscore, float *__restrict__ score,
                      float *__restrict__ gathered, int B, int L, int n) {
    const int b = blockIdx.x;
    Best best{-INFINITY, 0x7fffffff, 0.f};
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const float a0 = head_combine(h.cls, h.cls_bias, h.cls_w, nullptr, h.nlev, h.ntile, B, b, 2, 0, n, p);
        const float a1 = head_combine(h.cls, h.cls_bias, h.cls_w, nullptr, h.nlev, h.ntile, B, b, 2, 1, n, p);
        if (cls_out) {
            cls_out[((size_t)b * 2 + 0) * n + p] = a0;
            cls_out[((size_t)b * 2 + 1) * n + p] = a1;
        }
        if (loc_out)
            for (int l = 0; l < L; ++l) loc_out[((size_t)b * L + l) * n + p] = head_combine(h.loc, h.loc_bias, h.loc_w, h.loc_scale, h.nlev, h.ntile, B, b, L, l, n, p);
        const float m = fmaxf(a0, a1);
        const float e0 = expf(__fsub_rn(a0, m)), e1 = expf(__fsub_rn(a1, m));
        const float s = __fdiv_rn(e1, __fadd_rn(e0, e1));
        double ps;
        if (window) ps = __dadd_rn((double)__fmul_rn(s, one_minus_w), __dmul_rn(__ldg(window + p), w_infl));
        else ps = (double)s;
        best = better(best, Best{ps, p, s});
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Best o;
        o.v = __shfl_xor_sync(0xffffffffu, best.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
        o.s = __shfl_xor_sync(0xffffffffu, best.s, off);
        best = better(best, o);
    }
    __shared__ Best sb[32];
    __shared__ int s_idx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sb[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        Best r = sb[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = better(r, sb[w]);
        s_idx = min(max(r.i, 0), n - 1);
        idx[b] = s_idx;
        pscore[b] = r.v;
        score[b] = r.s;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x)
        gathered[b * L + l] = head_combine(h.loc, h.loc_bias, h.loc_w, h.loc_scale, h.nlev, h.ntile, B, b, L, l, n, s_idx);
}

}  // namespace hdn

using namespace hdn;

extern "C" int hdn_head_score_f32(int nlev, int ntile, const float *const *cls_parts_host, const float *const *loc_parts_host,
                                  const float *const *cls_bias_host, const float *const *loc_bias_host, const float *cls_w_host,
                                  const float *loc_scale_host, const float *loc_w_host, float *cls_out, float *loc_out, const double *window,
                                  double win_influence, int64_t *idx, double *pscore, float *score, float *gathered, int B, int L, int N,
                                  hdn_stream_t stream) {
    if (!cls_parts_host || !loc_parts_host || !cls_bias_host || !loc_bias_host || !cls_w_host || !loc_scale_host || !loc_w_host) return HDN_ERR_NULL;
    if (!idx || !pscore || !score || !gathered) return HDN_ERR_NULL;
    if (nlev < 1 || nlev > 4 || ntile < 1 || B < 1 || L < 1 || N < 1) return HDN_ERR_SHAPE;
    HeadLevels h{};
    for (int l = 0; l < nlev; ++l) {
        if (!cls_parts_host[l] || !loc_parts_host[l] || !cls_bias_host[l] || !loc_bias_host[l]) return HDN_ERR_NULL;
        h.cls[l] = cls_parts_host[l];
        h.loc[l] = loc_parts_host[l];
        h.cls_bias[l] = cls_bias_host[l];
        h.loc_bias[l] = loc_bias_host[l];
        h.cls_w[l] = cls_w_host[l];
        h.loc_scale[l] = loc_scale_host[l];
        h.loc_w[l] = loc_w_host[l];
    }
    h.nlev = nlev;
    h.ntile = ntile;
    // one block per pair; 1024 threads: the block is latency-bound on its ~20 independent loads per pixel, so width buys time
    head_score_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(h, cls_out, loc_out, window, win_influence, (float)(1.0 - win_influence),
                                                          reinterpret_cast<long long *>(idx), pscore, score, gathered, B, L, N * N);
    count_launch();
    return launch_status();
}

extern "C" int hdn_score_argmax_f32(const float *cls, const float *loc, const double *window, double win_influence, int64_t *idx,
                                    double *pscore, float *score, float *gathered, int B, int L, int N, hdn_stream_t stream) {
    if (!cls || !loc || !idx || !pscore || !score || !gathered) return HDN_ERR_NULL;
    if (B < 1 || L < 1 || N < 1) return HDN_ERR_SHAPE;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    score_argmax_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(cls, loc, window, win_influence, (float)(1.0 - win_influence),
                                                            reinterpret_cast<long long *>(idx), pscore, score, gathered, L, N * N);
    count_launch();
    return launch_status();
}
