/*
 * hdn_b200.h -- C ABI of libhdn_b200.so: the B200 (sm_100a) kernels behind HDN's
 * per-frame homography hot path.
 *
 * The reference (zhanxinrui/HDN) has no FFI layer: its operator boundary is a
 * set of Python callables on torch tensors.  Each entry point below replaces
 * one of them; the Python drop-ins that bind these symbols with the
 * reference's own names and signatures live in hdn_b200/ops.py and
 * hdn_b200/compat/ (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers to contiguous NCHW fp32 unless
 *     the parameter name ends in _host.  The caller owns every buffer.
 *   - Every launcher is asynchronous on the caller's stream (a cudaStream_t
 *     passed as void*; NULL = legacy default stream) and re-entrant: launches go
 *     to the CURRENT device of the calling thread; the only library state is
 *     per-device kernel configuration and the per-thread algorithm choice of
 *     hdn_xcorr_set_algo.
 *   - Return value: 0 on success, a negative hdn_status for a rejected
 *     argument (nothing was launched), or a positive cudaError_t if the launch
 *     itself failed.  No function throws or exits.  The only message ever
 *     printed (once per process, to stderr, silenced by HDN_B200_QUIET) says
 *     that a correlation shape has no tiled kernel and runs the slow generic one.
 */
#ifndef HDN_B200_H
#define HDN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDN_ABI_VERSION 1
#define HDN_MAX_PROBLEMS 8

typedef void *hdn_stream_t;

enum hdn_status {
    HDN_OK = 0,
    HDN_ERR_NULL = -1,        /* a required pointer is NULL */
    HDN_ERR_SHAPE = -2,       /* non-positive size, kernel larger than (padded) input, ... */
    HDN_ERR_ALIGN = -3,       /* pointer not 4-byte aligned */
    HDN_ERR_UNSUPPORTED = -4, /* argument combination not implemented */
    HDN_ERR_DEVICE = -5       /* no sm_100 device / driver failure */
};

int hdn_abi_version(void);
const char *hdn_status_string(int status);
/* SM count and compute capability of the current device. */
int hdn_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t hdn_launch_count(void);

/* K1 + K2.  Replaces hdn/core/xcorr.py:37-46 xcorr_depthwise (circular = 0) and
 * hdn/core/xcorr.py:48-61 xcorr_depthwise_circular (circular = 1; rows wrap by
 * Hx/2, columns replicate by Wx/2, no padded copy is materialised).
 *   x   [B,C,Hx,Wx]      k [B,C,Hk,Wk] with batch stride k_batch_stride elements
 *   out [B,C,Ho,Wo]      (0 = one template shared by the batch, C*Hk*Wk = dense)
 *   Ho = Hx + 2*(circular ? Hx/2 : 0) - Hk + 1, likewise Wo.
 *   out[b,c,i,j] = sum_{u,v} xpad[b,c,i+u,j+v] * k[b,c,u,v]   (no flip, stride 1) */
int hdn_xcorr_dw_f32(const float *x, const float *k, float *out, int B, int C, int Hx, int Wx, int Hk, int Wk, int circular,
                     int64_t k_batch_stride, hdn_stream_t stream);

/* The same operator for n <= HDN_MAX_PROBLEMS problems of identical shape in ONE launch
 * (the 3 levels x {cls,loc} correlations of MultiBAN.forward, hdn/models/head/ban.py:102-127,
 * and of MultiCircBAN.forward, ban_lp.py:65-92).  x_host/k_host/out_host are HOST arrays of n device pointers. */
int hdn_xcorr_dw_multi_f32(int n, const float *const *x_host, const float *const *k_host, float *const *out_host, int B, int C, int Hx,
                           int Wx, int Hk, int Wk, int circular, int64_t k_batch_stride, hdn_stream_t stream);

/* Algorithm of K1/K2 for the shapes that have both kernels (29x29 and 15x15 templates, where the direct sum is FMA-bound at
 * 30-140 flop/B): HDN_XCORR_DIRECT = the direct register-tiled sum; HDN_XCORR_FFT / HDN_XCORR_AUTO (default) = the transform-domain
 * kernel (row FFTs + per-frequency column correlation + inverse row FFTs, xcorr_fft.cu), faster on a B200 for all of them.
 * HDN_XCORR_FFT_PHASED = the same arithmetic in the barrier-separated three-phase kernel (kept for A/B measurements; AUTO / FFT run the
 * software-pipelined one: inverse FFTs of a group next to the forward FFTs of the next).
 * Per calling thread; all give the reference's result to ~3e-7 of max|out|.  hdn_xcorr_uses_fft: 1 if that shape (dense or shared
 * template, 16-byte aligned pointers) now takes the FFT kernel -- the same decision hdn_xcorr_dw*_f32 makes. */
enum hdn_xcorr_algo { HDN_XCORR_AUTO = 0, HDN_XCORR_DIRECT = 1, HDN_XCORR_FFT = 2, HDN_XCORR_FFT_PHASED = 3, HDN_XCORR_FFT_PIPE = 4, HDN_XCORR_FFT_WS = 5 };
int hdn_xcorr_set_algo(int algo);
int hdn_xcorr_uses_fft(int C, int Hx, int Wx, int Hk, int Wk, int circular, int64_t k_batch_stride);

/* Shared template with cached row spectra (29x29 templates at 256/512 crops: 61x61 plain, 29x29 circular).  A template that serves
 * many pairs -- config 3's broadcast template, the tracker's per-sequence template (hdn/tracker/hdn_tracker_proj_e2e.py:86-118 computes
 * it once in init) -- has its 29 row transforms taken ONCE: hdn_xcorr_template_spectra_f32 writes K'(u,f) of n templates [C,Hk,Wk]
 * into spectra buffers of hdn_xcorr_spectra_floats(...) floats each (0 = the shape has no such kernel), and
 * hdn_xcorr_dw_multi_spec_f32 is hdn_xcorr_dw_multi_f32 with k_batch_stride == 0 reading those spectra instead of the templates
 * (same result to fp32 rounding: ~3e-7 of max|out|).  All pointers 16-byte aligned, C % 4 == 0. */
int64_t hdn_xcorr_spectra_floats(int C, int Hx, int Wx, int Hk, int Wk, int circular);
int hdn_xcorr_template_spectra_f32(int n, const float *const *k, float *const *spectra, int C, int Hx, int Wx, int Hk, int Wk, int circular,
                                   hdn_stream_t stream);
int hdn_xcorr_dw_multi_spec_f32(int n, const float *const *x, const float *const *spectra, float *const *out, int B, int C, int Hx, int Wx,
                                int Hk, int Wk, int circular, hdn_stream_t stream);

/* 1 if (shape, 16-byte aligned pointers) takes the TMA-staged kernel, 0 if it takes the generic one-thread-per-output kernel. */
int hdn_xcorr_is_staged(int C, int Hx, int Wx, int Hk, int Wk, int circular, int64_t k_batch_stride);
/* How many correlation launches of this process fell to the generic kernel (shape outside the tiled table, unaligned pointers or a
 * strided template): correct but slow; the first one is also reported on stderr. */
int64_t hdn_xcorr_generic_launches(void);

/* K3.  Replaces hdn/models/logpolar.py:50-134 STN_Polar.forward: analytic log-polar
 * grid (rows = angle, cols = log-radius) + bilinear border sampling, align_corners=False.
 *   img [B,Ch,H,W] -> out [B,Ch,S,S];  S = INSTANCE_SIZE//2;  polar [B,2] or NULL (= zeros);
 *   rot_delta = delta[1] of the reference call. */
int hdn_logpolar_f32(const float *img, const float *polar, float rot_delta, float *out, int B, int Ch, int H, int W, int S,
                     hdn_stream_t stream);
/* The same for a uint8 crop [B,Ch,H,W] (the crop get_subwindow resizes is 8-bit; the reference widens it to fp32 on the host
 * before the upload, base_tracker.py:127-131): identical results, a quarter of the bytes across the bus. */
int hdn_logpolar_u8(const uint8_t *img, const float *polar, float rot_delta, float *out, int B, int Ch, int H, int W, int S,
                    hdn_stream_t stream);

/* K5.  Replaces Oneline_DLTv1/utils.py:7-67 DLT_solve for 8-vectors (one quad per item).
 *   src4, off4 [B,8] (x0,y0,...,x3,y3) -> Hm [B,9] row-major 3x3 with Hm[8] = 1. */
int hdn_dlt4_f32(const float *src4, const float *off4, float *Hm, int B, hdn_stream_t stream);

/* K4.  Replaces Oneline_DLTv1/utils.py:257-274 transform (-> transformer :70-254) for the
 * identity patch_indices the tracker passes (get_img_info.py:92).
 *   img [B,Ch,H,W], Hm [B,9] -> out [B,Ch,H,W];  theta = (Minv Hm) M.
 *   M_host / Minv_host: HOST float[9] (the reference passes M = [[63.5,0,63.5],[0,63.5,63.5],[0,0,1]]
 *   and torch.inverse(M), model_builder_e2e_unconstrained_v2.py:196-205); NULL = built from W/2, H/2. */
int hdn_homo_warp_f32(const float *img, const float *Hm, const float *M_host, const float *Minv_host, float *out, int B, int Ch, int H,
                      int W, hdn_stream_t stream);

/* K5 + K4 in one launch: offsets -> H -> warped image (what ModelBuilder.track_proj does back to back,
 * model_builder_e2e_unconstrained_v2.py:195-210).  Hm [B,9] is also written. */
int hdn_dlt_warp_f32(const float *src4, const float *off4, const float *img, const float *M_host, const float *Minv_host, float *Hm,
                     float *out, int B, int Ch, int H, int W, hdn_stream_t stream);

/* K6.  Replaces the device->host->NumPy epilogue: hdn/tracker/hdn_tracker.py:82-89 _convert_score,
 * hdn_tracker_proj_e2e.py:172-174 window blend + np.argmax, and the column read of
 * base_tracker.py:54-59 _convert_c / hdn_tracker.py:51-67 _convert_logpolar_simi.
 *   cls [B,2,N,N], loc [B,L,N,N], window [N*N] float64 or NULL (lp branch: no window)
 *   idx [B] int64 (first maximum), pscore [B] float64 (the value the 0.05 / 0.25 gates test),
 *   score [B] fp32 softmax p(fg) at idx ('best_score'), gathered [B,L] = loc[b,:,idx]. */
int hdn_score_argmax_f32(const float *cls, const float *loc, const double *window, double win_influence, int64_t *idx, double *pscore,
                         float *score, float *gathered, int B, int L, int N, hdn_stream_t stream);

/* Few-channel convolutions (K7c): the 7x7 stride-2 stems (3 -> 64 pad 0, hdn/models/backbone/resnet_atrous.py:121-131; 2 -> 64 pad 3,
 * Oneline_DLTv1/backbone/resnet.py:142-160) and PreShareFeature's 3x3 stride-1 layers (1 -> 4 -> 8 -> 1, pad 1,
 * Oneline_DLTv1/preprocess/input_feature_extractor.py:3-29) + eval-mode BatchNorm + optional ReLU as a direct fp32 FMA sum.
 *   x [B,Cin,H,W];  w [Cout,Cin,k,k] (the module's own layout);  scale/shift [Cout] or NULL;  out [B,Cout,Ho,Wo],
 *   Ho = (H + 2*pad - k) / stride + 1.  Supported: Cin <= 8 and (k,stride) = (7,2) with Cout % 32 == 0, or (3,1) with Cout == 1 or
 *   Cout % 4 == 0;  0 <= pad <= k/2. */
int hdn_conv_small_supported(int Cin, int Cout, int ksize, int stride);
int hdn_conv_small_f32(const float *x, const float *w, const float *scale, const float *shift, float *out, int B, int Cin, int Cout, int H,
                       int W, int ksize, int stride, int pad, int relu, hdn_stream_t stream);

/* Backbone / neck / head convolution: stride-1 1x1 or 3x3 convolution (valid = 0: padding =
 * dilation, same-size output; valid = 1: no padding, output H-2d x W-2d) + eval-mode
 * BatchNorm + optional residual add + optional ReLU, as a 3xTF32 (fp32-accurate) implicit GEMM on tcgen05 / TMEM.
 * Replaces nn.Conv2d -> nn.BatchNorm2d (-> += residual) (-> ReLU) of hdn/models/backbone/resnet_atrous.py:62-110
 * and hdn/models/neck/neck.py:11-29.
 *   x [B,Cin,H,W];  wpk = the weight packed ONCE by hdn_conv_pack_weight_f32;  scale/shift [Cout] or NULL
 *   (y = conv*scale + shift, the folded BatchNorm);  residual [B,Cout,Ho,Wo] or NULL;  out [B,Cout,Ho,Wo].
 * Requires hdn_conv_gemm_supported(): Cin % 32 == 0, Cout % 64 == 0, ksize in {1,3}.
 * hdn_conv_pack_weight_f32: wt [Cout, Ktot = ksize*ksize*Cin] tap-major (weight.permute(0,2,3,1)) -> packed [2*R*Ktot] floats,
 *   R = Cout rounded up to 128 (a 64-wide layer runs in a zero-padded 128-row tile): TF32 hi / lo halves, tiled per 128-channel x
 *   32-deep block in the tensor core's shared-memory layout.
 * hdn_conv_gemm_ex_f32: the same kernel with explicit stride (1 | 2) and zero padding (0 <= pad <= dilation * (ksize / 2)) -- the
 *   stride-2 3x3 / 1x1 layers of the ResNet stages (hdn/models/backbone/resnet_atrous.py:62-110,169-173; the homography estimator's
 *   ResNet-34, Oneline_DLTv1/backbone/resnet.py:65-194).  out [B,Cout,Ho,Wo], Ho = (H + 2*pad - dilation*(ksize-1) - 1) / stride + 1. */
int hdn_conv_gemm_supported(int Cin, int Cout, int ksize, int dilation);
int hdn_conv_gemm_ex_f32(const float *x, const float *wpk, const float *scale, const float *shift, const float *residual, float *out, int B,
                         int Cin, int Cout, int H, int W, int ksize, int stride, int pad, int dilation, int relu, hdn_stream_t stream);
/* Small problems (a 15x15 or 31x31 map at tracking batch sizes fills a fraction of the 148 SMs) split K over a thread-block
 * cluster of 2 / 4 / 8 CTAs whose fp32 partial tiles are added in rank order through distributed shared memory (deterministic).
 * enable = 0 switches that off (A/B runs); default on. */
int hdn_conv_gemm_set_splitk(int enable);
/* Consecutive convolution launches of a stream are chained by programmatic dependent launch: the next kernel's prologue (barriers,
 * TMEM allocation, first weight records) overlaps the tail of the previous one; its activation reads and all its writes wait for
 * the previous kernel's completion (griddepcontrol.wait).  enable = 0: plain stream order (A/B runs); default on. */
int hdn_conv_gemm_set_pdl(int enable);
/* Large launches (>= 2 CTAs per SM of 128 x 128 tiles) run conv_gemm_ts.cu: the split activations go from registers straight into
 * tensor memory and the MMAs take their A operand from there, so shared memory carries the weight records only.  enable = 0: the
 * shared-memory-operand kernel for every launch (A/B runs). */
int hdn_conv_gemm_set_ts(int enable);
/* 3x3 'valid' layers with W <= 63 (the heads' conv_search / conv_kernel) run a kernel that stages the activations once per
 * 32-channel block and feeds the nine taps as shifted windows of that tile (conv_shift.cu).  mode = 1 (default): on; 2: on, and
 * clusters of two neighbouring pixel tiles share every weight record through TMA multicast (built and correct, measured slower:
 * an A/B switch); 0: the generic implicit GEMM instead.  Same fp32-accurate 3xTF32 arithmetic in all three. */
int hdn_conv_gemm_set_shift(int mode);
int hdn_conv_pack_weight_f32(const float *wt, float *packed, int Cout, int Ktot, hdn_stream_t stream);
int hdn_conv_gemm_f32(const float *x, const float *wpk, const float *scale, const float *shift, const float *residual, float *out, int B,
                      int Cin, int Cout, int H, int W, int ksize, int dilation, int valid, int relu, hdn_stream_t stream);

/* n <= HDN_MAX_PROBLEMS convolutions of one shape in ONE launch (blockIdx.z = problem x image): the 3 levels x {cls, loc}
 * `conv_search` / `conv_kernel` layers of MultiBAN / MultiCircBAN (hdn/models/head/ban.py:56-61, ban_lp.py:19-24), each
 * 3x3 + BatchNorm + ReLU on a neck feature map.  *_host: HOST arrays of n device pointers (scale_host / shift_host may be NULL). */
int hdn_conv_gemm_multi_f32(int n, const float *const *x_host, const float *const *wpk_host, const float *const *scale_host,
                            const float *const *shift_host, float *const *out_host, int B, int Cin, int Cout, int H, int W, int ksize,
                            int dilation, int valid, int relu, hdn_stream_t stream);

/* Fused tail of DepthwiseXCorr (hdn/models/head/ban.py:62-66, :77): head = 1x1 (C->C, no bias) + BatchNorm + ReLU + 1x1 (C->L, bias)
 * applied to the correlation features, for n <= HDN_MAX_PROBLEMS branches in one launch.  The hidden C-channel map stays on chip:
 * each CTA multiplies its 128-channel x 64-pixel tile by the matching slice of the second convolution and stores PARTIAL sums.
 *   x_host[i]   [B,C,H,W] correlation features of branch i        wpk_host[i] packed first 1x1 (hdn_conv_pack_weight_f32)
 *   scale/shift [C] folded BatchNorm                               w2_host[i]  [L,C] second 1x1 weight (row-major, device)
 *   part_host[i] [C/128, B, L, H*W] partial sums (bias NOT added): hdn_head_score_f32 finishes them. */
int hdn_head_project_multi_f32(int n, const float *const *x_host, const float *const *wpk_host, const float *const *scale_host,
                               const float *const *shift_host, const float *const *w2_host, float *const *part_host, int B, int C, int H,
                               int W, int L, hdn_stream_t stream);

/* End of MultiBAN.forward (ban.py:102-127) fused with K6: per level l the partial sums are added (fixed order) + bias,
 * loc is scaled by loc_scale[l], the levels are combined with cls_w / loc_w (= softmax(cls_weight), softmax(loc_weight), HOST
 * floats), then hdn_score_argmax_f32's epilogue runs on the combined maps.
 *   cls_parts_host[l] [ntile,B,2,N*N], loc_parts_host[l] [ntile,B,L,N*N], cls_bias_host[l] [2], loc_bias_host[l] [L]  (device)
 *   cls_out [B,2,N,N] / loc_out [B,L,N,N]: the combined maps `track_new` returns; NULL = not stored. */
int hdn_head_score_f32(int nlev, int ntile, const float *const *cls_parts_host, const float *const *loc_parts_host,
                       const float *const *cls_bias_host, const float *const *loc_bias_host, const float *cls_w_host,
                       const float *loc_scale_host, const float *loc_w_host, float *cls_out, float *loc_out, const double *window,
                       double win_influence, int64_t *idx, double *pscore, float *score, float *gathered, int B, int L, int N,
                       hdn_stream_t stream);

/* ---- device-side frame pre-processing: bit-compatible with the OpenCV calls of the reference ------------------
 * Frames are uint8 HWC (BGR, 3 channels, dense) in DEVICE memory; small parameters are HOST pointers. */

/* cv2.warpPerspective(img, M, (W, H), borderMode=BORDER_REPLICATE) of hdn/tracker/hdn_tracker_proj_e2e.py:154 (bilinear, 5-bit
 * positions, 15-bit weights).  Minv_host: the 3x3 DESTINATION -> SOURCE map, i.e. cv2.invert(M) (row-major doubles): OpenCV inverts
 * the matrix it is given before sampling, the caller passes that inverse.  dst != src. */
int hdn_warp_perspective_u8(const uint8_t *src, uint8_t *dst, int H, int W, const double *Minv_host, hdn_stream_t stream);

/* cv2.warpAffine(img, M, (W, H), flags=2 (INTER_CUBIC), borderMode=BORDER_REPLICATE) of img_rot_around_center,
 * hdn/utils/transform.py:69-100.  M_host: the FORWARD 2x3 matrix the reference passes (it is inverted here exactly as cv::warpAffine
 * does).  tab_dev: the 32x32x16 int16 bicubic weight table in device memory; fill a host copy with hdn_cubic_table_host. */
int hdn_cubic_table_host(int16_t *tab /* [32*32*16] */);
int hdn_warp_affine_cubic_u8(const uint8_t *src, uint8_t *dst, int H, int W, const double *M_host, const int16_t *tab_dev, hdn_stream_t stream);

/* SiameseTracker.get_subwindow (hdn/tracker/base_tracker.py:61-136) on the device: the n x n window whose top-left pixel is
 * (x0, y0) in frame coordinates (it may stick out of the frame: outside pixels take fill_host[3], the channel means cast to uint8)
 * resized to S x S like cv2.resize (INTER_LINEAR; exact 2x decimation = INTER_AREA; n == S = copy) and converted to float32.
 *   gray = 0: out [3,S,S] planar BGR (what the reference uploads as x_crop)
 *   gray = 1: out [S,S] = get_search_info's normalised gray image, mean over channels of (v - mean_host[c]) / std_host[c] in float64
 *             (Oneline_DLTv1/tools/get_img_info.py:42-70). */
int hdn_crop_resize_u8(const uint8_t *frame, int H, int W, int x0, int y0, int n, const uint8_t *fill_host, int S, int gray,
                       const double *mean_host, const double *std_host, float *out, hdn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HDN_B200_H */
