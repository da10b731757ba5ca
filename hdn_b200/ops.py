"""Drop-in operators: the reference's names and signatures on top of libhdn_b200's C ABI.

Each callable replaces the reference function cited in its docstring; torch is used only for
device memory and the current stream.  CPU tensors are rejected -- there is no CPU fallback.
"""
import ctypes
import os
import math

import torch

from . import _lib

_vp = ctypes.c_void_p


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _dev(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: hdn_b200 kernels run on CUDA only (no CPU fallback)" % (name, t.device))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ptr(t):
    return _vp(t.data_ptr())


def _out(out, shape, like, name="out"):
    """A caller-supplied result buffer goes to the kernel as a raw pointer: insist on exactly what the kernel will write."""
    if out is None:
        return torch.empty(shape, device=like.device, dtype=torch.float32)
    if not isinstance(out, torch.Tensor) or tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or out.device != like.device \
            or not out.is_contiguous():
        raise ValueError("%s must be a contiguous float32 tensor of shape %s on %s" % (name, tuple(shape), like.device))
    return out


class _on_device:
    """Launch on the device that owns the tensors (the C ABI launches on the calling thread's CURRENT device)."""

    def __init__(self, t):
        self.ctx = torch.cuda.device(t.device) if t.device.index is not None and t.device.index != torch.cuda.current_device() else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def xcorr_out_hw(Hx, Wx, Hk, Wk, circular):
    ph, pw = (Hx // 2, Wx // 2) if circular else (0, 0)
    return Hx + 2 * ph - Hk + 1, Wx + 2 * pw - Wk + 1


def _xcorr(x, kernel, circular, out=None):
    x, kernel = _dev(x, "x"), _dev(kernel, "kernel")
    if x.dim() != 4 or kernel.dim() != 4:
        raise RuntimeError("xcorr expects 4-D NCHW tensors")
    B, C, Hx, Wx = x.shape
    Bk, Ck, Hk, Wk = kernel.shape
    if Ck != C or (Bk != B and Bk != 1):
        raise RuntimeError("xcorr: x %s and kernel %s disagree in batch/channels" % (tuple(x.shape), tuple(kernel.shape)))
    Ho, Wo = xcorr_out_hw(Hx, Wx, Hk, Wk, circular)
    if Ho < 1 or Wo < 1:
        raise RuntimeError("xcorr: kernel %dx%d larger than (padded) input %dx%d" % (Hk, Wk, Hx, Wx))
    if kernel.device != x.device:
        raise RuntimeError("xcorr: x is on %s but kernel on %s" % (x.device, kernel.device))
    out = _out(out, (B, C, Ho, Wo), x)
    kbs = 0 if (Bk == 1 and B > 1) else C * Hk * Wk
    with _on_device(x):
        st = _lib.lib().hdn_xcorr_dw_f32(_ptr(x), _ptr(kernel), _ptr(out), B, C, Hx, Wx, Hk, Wk, int(circular), kbs, _stream())
    _lib.check(st, "hdn_xcorr_dw_f32")
    return out


XCORR_ALGOS = {"auto": 0, "direct": 1, "fft": 2, "fft_phased": 3, "fft_pipe": 4, "fft_ws": 5}


def set_xcorr_algo(name):
    """'auto' (default) / 'fft': the 29x29- and 15x15-template shapes run the transform-domain kernel (xcorr_fft.cu);
    'direct': every shape runs the direct register-tiled sum (xcorr.cu).  Process-wide."""
    _lib.check(_lib.lib().hdn_xcorr_set_algo(XCORR_ALGOS[name]), "hdn_xcorr_set_algo")


def xcorr_depthwise(x, kernel, out=None):
    """hdn/core/xcorr.py:37-46.  x [B,C,Hx,Wx], kernel [B,C,h,w] (or [1,C,h,w] = template shared by the batch)."""
    return _xcorr(x, kernel, False, out)


def xcorr_depthwise_circular(x, kernel, out=None):
    """hdn/core/xcorr.py:48-61: rows wrap by H//2, columns replicate by W//2, then depth-wise correlation."""
    return _xcorr(x, kernel, True, out)


def xcorr_depthwise_multi(xs, kernels, circular=False, outs=None):
    """n <= 8 same-shape correlations in ONE launch (the 3 levels x {cls,loc} loop of MultiBAN.forward,
    hdn/models/head/ban.py:102-127 / ban_lp.py:65-92)."""
    n = len(xs)
    xs = [_dev(x, "x") for x in xs]
    kernels = [_dev(k, "kernel") for k in kernels]
    B, C, Hx, Wx = xs[0].shape
    Bk, _, Hk, Wk = kernels[0].shape
    for x, k in zip(xs, kernels):
        if tuple(x.shape) != (B, C, Hx, Wx) or tuple(k.shape) != (Bk, C, Hk, Wk):
            raise RuntimeError("xcorr_depthwise_multi: all problems must share one shape")
    Ho, Wo = xcorr_out_hw(Hx, Wx, Hk, Wk, circular)
    outs = [_out(None if outs is None else outs[i], (B, C, Ho, Wo), xs[0], "outs[%d]" % i) for i in range(n)]
    arr = _vp * n
    kbs = 0 if (Bk == 1 and B > 1) else C * Hk * Wk
    with _on_device(xs[0]):
        st = _lib.lib().hdn_xcorr_dw_multi_f32(n, arr(*[x.data_ptr() for x in xs]), arr(*[k.data_ptr() for k in kernels]),
                                               arr(*[o.data_ptr() for o in outs]), B, C, Hx, Wx, Hk, Wk, int(circular), kbs, _stream())
    _lib.check(st, "hdn_xcorr_dw_multi_f32")
    return outs


TEMPLATE_SPECTRA = os.environ.get("HDN_B200_TEMPLATE_SPECTRA", "1") != "0"  # cached row spectra for shared templates (A/B switch)


def xcorr_template_spectra(kernels, Hx, Wx, circular=False):
    """Row spectra of n <= 8 SHARED templates [1,C,Hk,Wk] (one launch), for xcorr_depthwise_multi(..., spectra=...): a template that
    serves many pairs / frames has its row transforms taken once (hdn_xcorr_template_spectra_f32).  -> list of 1-D float32 tensors, or
    None when the shape has no cached-spectra kernel (everything but the 29x29 templates at 256/512 crops) or the switch is off."""
    kernels = [_dev(k, "kernel") for k in kernels]
    Bk, C, Hk, Wk = kernels[0].shape
    if Bk != 1 or any(tuple(k.shape) != (1, C, Hk, Wk) for k in kernels):
        raise RuntimeError("xcorr_template_spectra: templates must be [1,C,Hk,Wk] and share one shape")
    nf = int(_lib.lib().hdn_xcorr_spectra_floats(C, Hx, Wx, Hk, Wk, int(circular))) if TEMPLATE_SPECTRA else 0
    if nf == 0 or any(k.data_ptr() % 16 for k in kernels):
        return None
    spectra = [torch.empty(nf, device=kernels[0].device, dtype=torch.float32) for _ in kernels]
    n = len(kernels)
    arr = _vp * n
    with _on_device(kernels[0]):
        st = _lib.lib().hdn_xcorr_template_spectra_f32(n, arr(*[k.data_ptr() for k in kernels]), arr(*[t.data_ptr() for t in spectra]), C, Hx, Wx,
                                                       Hk, Wk, int(circular), _stream())
    _lib.check(st, "hdn_xcorr_template_spectra_f32")
    return spectra


def xcorr_depthwise_multi_spec(xs, spectra, Hk, Wk, circular=False, outs=None):
    """xcorr_depthwise_multi for a shared template given by its cached row spectra (xcorr_template_spectra)."""
    n = len(xs)
    xs = [_dev(x, "x") for x in xs]
    B, C, Hx, Wx = xs[0].shape
    if any(tuple(x.shape) != (B, C, Hx, Wx) for x in xs) or len(spectra) != n:
        raise RuntimeError("xcorr_depthwise_multi_spec: all problems must share one shape")
    Ho, Wo = xcorr_out_hw(Hx, Wx, Hk, Wk, circular)
    outs = [_out(None if outs is None else outs[i], (B, C, Ho, Wo), xs[0], "outs[%d]" % i) for i in range(n)]
    arr = _vp * n
    with _on_device(xs[0]):
        st = _lib.lib().hdn_xcorr_dw_multi_spec_f32(n, arr(*[x.data_ptr() for x in xs]), arr(*[t.data_ptr() for t in spectra]),
                                                    arr(*[o.data_ptr() for o in outs]), B, C, Hx, Wx, Hk, Wk, int(circular), _stream())
    _lib.check(st, "hdn_xcorr_dw_multi_spec_f32")
    return outs


def logpolar_sample(x, polar=None, rot_delta=0.0, out_size=None, out=None):
    """Functional form of STN_Polar.forward (hdn/models/logpolar.py:120-134)."""
    x = _dev(x, "x")
    B, Ch, H, W = x.shape
    S = int(out_size)
    if polar is not None:
        polar = _dev(polar, "polar")
        if tuple(polar.shape) != (B, 2):
            raise RuntimeError("polar must be [B,2]")
    out = _out(out, (B, Ch, S, S), x)
    st = _lib.lib().hdn_logpolar_f32(_ptr(x), _ptr(polar) if polar is not None else None, float(rot_delta), _ptr(out), B, Ch, H, W, S,
                                     _stream())
    _lib.check(st, "hdn_logpolar_f32")
    return out


class STN_Polar(torch.nn.Module):
    """hdn/models/logpolar.py:50-134.  forward(x, polar, delta=[0,0]) -> (x_lp, grid).

    The sampling grid is analytic inside the kernel, so nothing is built on the host or uploaded.
    `grid` (which no caller on the tracking path reads) is returned only when `return_grid` is set,
    and is then produced by the same formulas with torch ops on the device."""

    def __init__(self, image_sz, return_grid=False):
        super().__init__()
        self._orignal_sz = [image_sz // 2, image_sz // 2]
        self.return_grid = return_grid

    def forward(self, x, polar, delta=[0, 0]):
        S = self._orignal_sz[0]
        out = logpolar_sample(x, polar, float(delta[1]), S)
        grid = self._grid(x, polar, float(delta[1])) if self.return_grid else None
        return out, grid

    def _grid(self, x, polar, rot):
        S = self._orignal_sz[0]
        dev = x.device
        j = torch.arange(S, device=dev, dtype=torch.float32)
        rho = torch.exp(math.log(S / 2) / S * j) - 1.0
        th = j * 2.0 * math.pi / S + rot
        gx = rho[None, :] * torch.cos(th)[:, None]
        gy = rho[None, :] * torch.sin(th)[:, None]
        gx = gx[None] + polar[:, 0].reshape(-1, 1, 1)
        gy = gy[None] + polar[:, 1].reshape(-1, 1, 1)
        return torch.stack((gx / (x.size(2) // 2), gy / (x.size(3) // 2)), 3)


def DLT_solve(src_p, off_set):
    """Oneline_DLTv1/utils.py:7-67 for 8-vectors: src_p, off_set [B,8] -> H [B,1,3,3]."""
    src_p, off_set = _dev(src_p, "src_p"), _dev(off_set, "off_set")
    if src_p.dim() != 2 or src_p.shape[1] != 8 or tuple(off_set.shape) != tuple(src_p.shape):
        raise RuntimeError("DLT_solve: only the 4-point (8-vector) form is implemented; got %s / %s" % (tuple(src_p.shape), tuple(off_set.shape)))
    B = src_p.shape[0]
    H = torch.empty((B, 1, 3, 3), device=src_p.device, dtype=torch.float32)
    _lib.check(_lib.lib().hdn_dlt4_f32(_ptr(src_p), _ptr(off_set), _ptr(H), B, _stream()), "hdn_dlt4_f32")
    return H


def _m9(t):
    """[...,3,3] tensor (batch-expanded constant) -> ctypes float[9] from its first matrix (one tiny D2H if on device)."""
    m = t.reshape(-1, 3, 3)[0].detach().to("cpu", torch.float32).reshape(9).tolist()
    return (ctypes.c_float * 9)(*m)


_M_CACHE = {}


def _m9_cached(t):
    """M / M^-1 are per-model constants: read them from the device once per tensor OBJECT.  The entry holds a weak reference to
    the tensor it was read from -- a new tensor that merely reuses a freed address (the reference builds M afresh every call)
    misses and is read again, so a stale matrix can never be returned."""
    import weakref
    key = (t.data_ptr(), t._version, tuple(t.shape), str(t.device))
    hit = _M_CACHE.get(key)
    if hit is not None and hit[0]() is t:
        return hit[1]
    if len(_M_CACHE) > 64:
        _M_CACHE.clear()
    v = _m9(t)
    _M_CACHE[key] = (weakref.ref(t), v)
    return v


def homo_warp(I1, H_mat, M=None, M_inv=None, out=None):
    """Projective warp core of `transform` (identity patch_indices). M / M_inv: python lists of 9 floats or None."""
    I1, H_mat = _dev(I1, "I1"), _dev(H_mat, "H_mat")
    B, Ch, H, W = I1.shape
    if H_mat.numel() != B * 9:
        raise RuntimeError("H_mat must hold one 3x3 per batch item")
    out = _out(out, tuple(I1.shape), I1)
    mp = (ctypes.c_float * 9)(*M) if M is not None and not isinstance(M, ctypes.Array) else M
    mip = (ctypes.c_float * 9)(*M_inv) if M_inv is not None and not isinstance(M_inv, ctypes.Array) else M_inv
    st = _lib.lib().hdn_homo_warp_f32(_ptr(I1), _ptr(H_mat), mp, mip, _ptr(out), B, Ch, H, W, _stream())
    _lib.check(st, "hdn_homo_warp_f32")
    return out


def dlt_warp(src_p, off_set, I1, M=None, M_inv=None):
    """K5 + K4 in one launch (what track_proj does back to back, model_builder...py:195-210). -> (H [B,3,3], warped)."""
    src_p, off_set, I1 = _dev(src_p, "src_p"), _dev(off_set, "off_set"), _dev(I1, "I1")
    B, Ch, H, W = I1.shape
    Hm = torch.empty((B, 3, 3), device=I1.device, dtype=torch.float32)
    out = torch.empty_like(I1)
    mp = (ctypes.c_float * 9)(*M) if M is not None else None
    mip = (ctypes.c_float * 9)(*M_inv) if M_inv is not None else None
    st = _lib.lib().hdn_dlt_warp_f32(_ptr(src_p), _ptr(off_set), _ptr(I1), mp, mip, _ptr(Hm), _ptr(out), B, Ch, H, W, _stream())
    _lib.check(st, "hdn_dlt_warp_f32")
    return Hm, out


def transform(patch_size_h, patch_size_w, M_tile_inv, H_mat, M_tile, I1, patch_indices, batch_indices_tensor):
    """Oneline_DLTv1/utils.py:257-274.  Same arguments and result ([B,C,ph,pw]).

    `patch_indices=None` (or a tensor tagged `_hdn_identity`) declares the arange(H*W) indices the tracker always
    passes (get_img_info.py:92): the trailing gather is then the identity and is skipped.  Any other index tensor
    is honoured with a device-side gather."""
    B, Ch, H, W = I1.shape
    warped = homo_warp(I1, H_mat, _m9_cached(M_tile), _m9_cached(M_tile_inv))
    if patch_indices is None or getattr(patch_indices, "_hdn_identity", False):
        if patch_size_h * patch_size_w != H * W:
            raise RuntimeError("identity patch_indices need patch size == image size")
        return warped.reshape(B, Ch, patch_size_h, patch_size_w)
    flat = warped.permute(0, 2, 3, 1).reshape(-1, Ch)
    pix = patch_indices.reshape(-1).long().to(flat.device) + batch_indices_tensor.to(flat.device)
    return flat.index_select(0, pix).reshape(B, patch_size_h, patch_size_w, Ch).permute(0, 3, 1, 2)


def score_argmax(cls, loc, window=None, win_influence=0.0):
    """K6: fused _convert_score + window blend + np.argmax + column read (hdn_tracker.py:82-89,
    hdn_tracker_proj_e2e.py:172-174, base_tracker.py:54-59).
    cls [B,2,N,N], loc [B,L,N,N], window float64 [N*N] on device or None.
    -> idx int64 [B], pscore float64 [B], score float32 [B], gathered float32 [B,L] (all on device)."""
    cls, loc = _dev(cls, "cls"), _dev(loc, "loc")
    B, two, N, N2 = cls.shape
    if two != 2 or N != N2 or loc.shape[0] != B or tuple(loc.shape[2:]) != (N, N):
        raise RuntimeError("score_argmax: cls must be [B,2,N,N] and loc [B,L,N,N]")
    L = loc.shape[1]
    if window is not None:
        window = _dev(window, "window", torch.float64)
        if window.numel() != N * N:
            raise RuntimeError("window must have N*N entries")
    dev = cls.device
    idx = torch.empty(B, device=dev, dtype=torch.int64)
    ps = torch.empty(B, device=dev, dtype=torch.float64)
    sc = torch.empty(B, device=dev, dtype=torch.float32)
    g = torch.empty((B, L), device=dev, dtype=torch.float32)
    st = _lib.lib().hdn_score_argmax_f32(_ptr(cls), _ptr(loc), _ptr(window) if window is not None else None, float(win_influence), _ptr(idx),
                                         _ptr(ps), _ptr(sc), _ptr(g), B, L, N, _stream())
    _lib.check(st, "hdn_score_argmax_f32")
    return idx, ps, sc, g


def score_argmax_packed(cls, loc, window=None, win_influence=0.0):
    """K6 writing all four results into ONE device byte buffer: [idx i64 x B | pscore f64 x B | score f32 x B | loc[:,idx] f32 x B*L].
    Graph-capturable (no host sync).  Decode the host copy with `unpack_scores`."""
    cls, loc = _dev(cls, "cls"), _dev(loc, "loc")
    B, _, N, _ = cls.shape
    L = loc.shape[1]
    if window is not None:
        window = _dev(window, "window", torch.float64)
        if window.numel() != N * N:
            raise RuntimeError("window must have N*N entries")
    o_ps, o_sc, o_g, total = 8 * B, 16 * B, 20 * B, 20 * B + 4 * B * L
    buf = torch.empty(total + (-total) % 8, device=cls.device, dtype=torch.uint8)
    base = buf.data_ptr()
    st = _lib.lib().hdn_score_argmax_f32(_ptr(cls), _ptr(loc), _ptr(window) if window is not None else None, float(win_influence), _vp(base),
                                         _vp(base + o_ps), _vp(base + o_sc), _vp(base + o_g), B, L, N, _stream())
    _lib.check(st, "hdn_score_argmax_f32")
    return buf


def unpack_scores(host, B, L):
    """host: NumPy uint8 copy of a `score_argmax_packed` buffer -> (idx i64[B], pscore f64[B], score f32[B], gathered f32[B,L])."""
    o_ps, o_sc, o_g, total = 8 * B, 16 * B, 20 * B, 20 * B + 4 * B * L
    return (host[:o_ps].view("<i8"), host[o_ps:o_sc].view("<f8"), host[o_sc:o_g].view("<f4"), host[o_g:total].view("<f4").reshape(B, L))


def score_argmax_host(cls, loc, window=None, win_influence=0.0):
    """K6 + ONE packed device->host copy (8+8+4+4L bytes per item instead of the whole score / loc maps; the
    reference moves both maps to the host and does this in NumPy).  -> NumPy (idx, pscore, score, gathered)."""
    buf = score_argmax_packed(cls, loc, window, win_influence)
    return unpack_scores(buf.cpu().numpy(), cls.shape[0], loc.shape[1])


def conv_gemm(x, wpk, scale=None, shift=None, residual=None, ksize=1, dilation=1, relu=False, out=None, valid=False, stride=1, padding=None,
              cout=None):
    """1x1 / 3x3 convolution + folded BatchNorm (+ residual) (+ ReLU) on tcgen05 (3xTF32, fp32-accurate).
    x [B,Cin,H,W]; wpk = pack_conv_weight(weight) (done once per layer); scale/shift [Cout].
    Geometry: stride 1 | 2; padding None = 'same' (dilation * (ksize // 2)) or 0 when `valid`; else 0 <= padding <= dilation * (ksize // 2).
    cout: the layer's output channels when they are not a multiple of 128 (the packed weight is padded to 128-row tiles)."""
    x, wpk = _dev(x, "x"), _dev(wpk, "wpk")
    B, Cin, H, W = x.shape
    rows = wpk.numel() // (2 * ksize * ksize * Cin)
    if wpk.numel() != 2 * rows * ksize * ksize * Cin or rows % 128:
        raise RuntimeError("conv_gemm: packed weight of %d floats does not match Cin=%d, ksize=%d" % (wpk.numel(), Cin, ksize))
    Cout = rows if cout is None else int(cout)
    if not (rows - 128 < Cout <= rows):
        raise RuntimeError("conv_gemm: cout=%d does not fit the packed weight (%d rows)" % (Cout, rows))
    pad = (0 if valid else dilation * (ksize // 2)) if padding is None else int(padding)
    Ho = (H + 2 * pad - dilation * (ksize - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dilation * (ksize - 1) - 1) // stride + 1
    out = _out(out, (B, Cout, Ho, Wo), x)
    opt = lambda t, n: _ptr(_dev(t, n)) if t is not None else None  # noqa: E731
    if residual is not None and tuple(residual.shape) != tuple(out.shape):
        raise RuntimeError("conv_gemm: residual shape mismatch")
    with _on_device(x):
        st = _lib.lib().hdn_conv_gemm_ex_f32(_ptr(x), _ptr(wpk), opt(scale, "scale"), opt(shift, "shift"), opt(residual, "residual"), _ptr(out), B, Cin,
                                             Cout, H, W, ksize, int(stride), pad, dilation, int(bool(relu)), _stream())
    _lib.check(st, "hdn_conv_gemm_ex_f32")
    return out


def conv_small(x, weight, scale=None, shift=None, stride=1, padding=0, relu=False, out=None):
    """Few-channel convolution + folded BatchNorm (+ ReLU) as a direct fp32 sum (hdn_conv_small_f32): the 7x7 stride-2 stems and
    PreShareFeature's 3x3 layers.  x [B,Cin<=8,H,W]; weight [Cout,Cin,k,k] as the module holds it."""
    x, weight = _dev(x, "x"), _dev(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, Cw, k, k2 = weight.shape
    if Cw != Cin or k != k2:
        raise RuntimeError("conv_small: weight %s does not match %d input channels" % (tuple(weight.shape), Cin))
    Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
    out = _out(out, (B, Cout, Ho, Wo), x)
    opt = lambda t, n: _ptr(_dev(t, n)) if t is not None else None  # noqa: E731
    with _on_device(x):
        st = _lib.lib().hdn_conv_small_f32(_ptr(x), _ptr(weight), opt(scale, "scale"), opt(shift, "shift"), _ptr(out), B, Cin, Cout, H, W, k,
                                           int(stride), int(padding), int(bool(relu)), _stream())
    _lib.check(st, "hdn_conv_small_f32")
    return out


def conv_small_supported(Cin, Cout, ksize, stride):
    return bool(_lib.lib().hdn_conv_small_supported(Cin, Cout, ksize, stride))


def _ptr_array(tensors):
    return (_vp * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])


def conv_gemm_multi(xs, wpks, scales, shifts, ksize=3, dilation=1, relu=True, valid=True, outs=None):
    """n <= 8 same-shape convolutions (+ folded BatchNorm, ReLU) in ONE launch: the `conv_search` layers of the 3 levels x {cls, loc}
    branches of MultiBAN / MultiCircBAN (hdn/models/head/ban.py:56-61).  xs[i] [B,Cin,H,W]; wpks[i] = pack_conv_weight(weight_i)."""
    n = len(xs)
    xs = [_dev(x, "x") for x in xs]
    B, Cin, H, W = xs[0].shape
    if any(tuple(x.shape) != (B, Cin, H, W) for x in xs):
        raise RuntimeError("conv_gemm_multi: all problems must share one shape")
    Cout = wpks[0].numel() // (2 * ksize * ksize * Cin)
    shrink = 2 * dilation if (valid and ksize == 3) else 0
    outs = [_out(None if outs is None else outs[i], (B, Cout, H - shrink, W - shrink), xs[0], "outs[%d]" % i) for i in range(n)]
    with _on_device(xs[0]):
        st = _lib.lib().hdn_conv_gemm_multi_f32(n, _ptr_array(xs), _ptr_array(wpks), _ptr_array(scales), _ptr_array(shifts), _ptr_array(outs), B, Cin,
                                                Cout, H, W, ksize, dilation, int(bool(valid)), int(bool(relu)), _stream())
    _lib.check(st, "hdn_conv_gemm_multi_f32")
    return outs


def head_project_multi(feats, wpks, scales, shifts, w2s, parts=None):
    """Fused tail of DepthwiseXCorr for n branches in one launch: 1x1 (C->C) + BatchNorm + ReLU + 1x1 (C->L) on the correlation
    features; the hidden map stays on chip.  w2s[i] [L,C].  -> parts[i] [C/128, B, L, H*W] partial sums (no bias; see head_score)."""
    n = len(feats)
    feats = [_dev(f, "feature") for f in feats]
    B, C, H, W = feats[0].shape
    L = w2s[0].shape[0]
    if any(tuple(f.shape) != (B, C, H, W) for f in feats) or any(tuple(w.shape) != (L, C) for w in w2s):
        raise RuntimeError("head_project_multi: all branches must share one shape")
    w2s = [_dev(w, "w2") for w in w2s]
    parts = [_out(None if parts is None else parts[i], (C // 128, B, L, H * W), feats[0], "parts[%d]" % i) for i in range(n)]
    with _on_device(feats[0]):
        st = _lib.lib().hdn_head_project_multi_f32(n, _ptr_array(feats), _ptr_array(wpks), _ptr_array(scales), _ptr_array(shifts), _ptr_array(w2s),
                                                   _ptr_array(parts), B, C, H, W, L, _stream())
    _lib.check(st, "hdn_head_project_multi_f32")
    return parts


def head_score(cls_parts, loc_parts, cls_bias, loc_bias, cls_w, loc_scale, loc_w, N, window=None, win_influence=0.0, want_maps=True,
               maps=None, packed=None):
    """End of MultiBAN.forward (ban.py:102-127) + K6 in one launch.  cls_parts[l] [ntile,B,2,N*N], loc_parts[l] [ntile,B,L,N*N] from
    head_project_multi; cls_bias[l] [2], loc_bias[l] [L] device tensors; cls_w / loc_scale / loc_w: python floats per level.
    -> (cls [B,2,N,N] | None, loc [B,L,N,N] | None, packed uint8 buffer in score_argmax_packed's layout)."""
    nlev = len(cls_parts)
    ntile, B, two, n = cls_parts[0].shape
    L = loc_parts[0].shape[2]
    if two != 2 or n != N * N:
        raise RuntimeError("head_score: cls parts must be [ntile,B,2,N*N]")
    dev = cls_parts[0].device
    if window is not None:
        window = _dev(window, "window", torch.float64)
        if window.numel() != N * N:
            raise RuntimeError("window must have N*N entries")
    if want_maps:
        cls, loc = maps if maps is not None else (torch.empty((B, 2, N, N), device=dev), torch.empty((B, L, N, N), device=dev))
    else:
        cls = loc = None
    o_ps, o_sc, o_g, total = 8 * B, 16 * B, 20 * B, 20 * B + 4 * B * L
    buf = packed if packed is not None else torch.empty(total + (-total) % 8, device=dev, dtype=torch.uint8)
    base = buf.data_ptr()
    fl = ctypes.c_float * nlev
    with _on_device(cls_parts[0]):
        st = _lib.lib().hdn_head_score_f32(nlev, ntile, _ptr_array(cls_parts), _ptr_array(loc_parts), _ptr_array(cls_bias), _ptr_array(loc_bias),
                                           fl(*cls_w), fl(*loc_scale), fl(*loc_w), _ptr(cls) if cls is not None else None,
                                           _ptr(loc) if loc is not None else None, _ptr(window) if window is not None else None,
                                           float(win_influence), _vp(base), _vp(base + o_ps), _vp(base + o_sc), _vp(base + o_g), B, L, N, _stream())
    _lib.check(st, "hdn_head_score_f32")
    return cls, loc, buf


def set_conv_splitk(enable=True):
    """Split-K over a thread-block cluster for small convolutions (default on); off = one CTA per output tile (A/B runs)."""
    _lib.check(_lib.lib().hdn_conv_gemm_set_splitk(int(bool(enable))), "hdn_conv_gemm_set_splitk")


def set_conv_ts(enable=True):
    """Large convolution launches with the activations in tensor memory (conv_gemm_ts.cu); False = shared-memory operands (A/B runs)."""
    _lib.check(_lib.lib().hdn_conv_gemm_set_ts(int(enable)), "hdn_conv_gemm_set_ts")  # 2: small launches as well (split-K clusters)


def set_conv_pdl(enable=True):
    """Programmatic dependent launch between consecutive tcgen05 convolutions (default on); False = plain stream order (A/B runs)."""
    _lib.check(_lib.lib().hdn_conv_gemm_set_pdl(int(bool(enable))), "hdn_conv_gemm_set_pdl")


def set_conv_shift(mode=True):
    """3x3 'valid' layers on the shifted-window kernel (conv_shift.cu).  True / 1 (default): on; 2: on, with weight multicast
    across 2-CTA clusters (correct, measured slower: A/B switch); False / 0: the generic implicit GEMM."""
    mode = 1 if mode is True else int(mode)
    _lib.check(_lib.lib().hdn_conv_gemm_set_shift(mode), "hdn_conv_gemm_set_shift")


def conv_gemm_supported(Cin, Cout, ksize, dilation=1):
    return bool(_lib.lib().hdn_conv_gemm_supported(Cin, Cout, ksize, dilation))


def pack_conv_weight(weight):
    """[Cout,Cin,k,k] conv weight -> the tensor-core record format of hdn_conv_pack_weight_f32 (2*R*k*k*Cin floats, R = Cout rounded
    up to 128): tap-major K (K index = tap*Cin + ci), TF32 hi / lo halves, tiled per 128-channel x 32-deep K block."""
    wt = _dev(weight.detach().permute(0, 2, 3, 1).reshape(weight.shape[0], -1), "weight")
    rows = (wt.shape[0] + 127) // 128 * 128
    packed = torch.empty(2 * rows * wt.shape[1], device=wt.device, dtype=torch.float32)
    _lib.check(_lib.lib().hdn_conv_pack_weight_f32(_ptr(wt), _ptr(packed), wt.shape[0], wt.shape[1], _stream()), "hdn_conv_pack_weight_f32")
    return packed
