"""ctypes front-end of oracle/hdn_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg -- never by hdn_b200/ (the product path).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhdn_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_d = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "hdn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hdn_oracle.c")
        if not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(_SO) < os.path.getmtime(src)):
            build()
        L = ctypes.CDLL(_SO)
        ci, cll = ctypes.c_int, ctypes.c_longlong
        L.orc_xcorr_dw.argtypes = [_f, _f, _f, ci, ci, ci, ci, ci, ci, ci, cll]
        L.orc_logpolar.argtypes = [_f, _f, ctypes.c_float, _f, ci, ci, ci, ci, ci]
        L.orc_dlt4.argtypes = [_f, _f, _f, ci]
        L.orc_homo_warp.argtypes = [_f, _f, _f, _f, _f, ci, ci, ci, ci]
        L.orc_score_argmax.argtypes = [_f, _f, _d, ctypes.c_double, _i64, _d, _f, _f, ci, ci, ci]
        for fn in (L.orc_xcorr_dw, L.orc_logpolar, L.orc_dlt4, L.orc_homo_warp, L.orc_score_argmax):
            fn.restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, typ=_f):
    return a.ctypes.data_as(typ)


def xcorr_out_shape(Hx, Wx, Hk, Wk, circular):
    ph, pw = (Hx // 2, Wx // 2) if circular else (0, 0)
    return Hx + 2 * ph - Hk + 1, Wx + 2 * pw - Wk + 1


def xcorr_dw(x, k, circular=False):
    """x [B,C,Hx,Wx]; k [B,C,Hk,Wk] or [1,C,Hk,Wk] (template shared by the batch)."""
    x, k = _f32(x), _f32(k)
    B, C, Hx, Wx = x.shape
    Hk, Wk = k.shape[2:]
    Ho, Wo = xcorr_out_shape(Hx, Wx, Hk, Wk, circular)
    out = np.empty((B, C, Ho, Wo), np.float32)
    kb = 0 if (k.shape[0] == 1 and B > 1) else C * Hk * Wk
    lib().orc_xcorr_dw(_p(x), _p(k), _p(out), B, C, Hx, Wx, Hk, Wk, int(bool(circular)), kb)
    return out


def logpolar(img, polar, delta_rot, S):
    img = _f32(img)
    B, Ch, H, W = img.shape
    out = np.empty((B, Ch, S, S), np.float32)
    pol = _f32(polar) if polar is not None else None
    lib().orc_logpolar(_p(img), _p(pol) if pol is not None else None, float(np.float32(delta_rot)), _p(out), B, Ch, H, W, S)
    return out


def dlt4(src, off):
    src, off = _f32(src), _f32(off)
    B = src.shape[0]
    H = np.empty((B, 3, 3), np.float32)
    lib().orc_dlt4(_p(src), _p(off), _p(H), B)
    return H


def homo_warp(img, Hm, M=None, Minv=None):
    img, Hm = _f32(img), _f32(Hm).reshape(-1, 9)
    B, Ch, H, W = img.shape
    if M is None:
        M, Minv = default_M(W, H)
    M, Minv = _f32(M), _f32(Minv)
    out = np.empty_like(img)
    lib().orc_homo_warp(_p(img), _p(Hm), _p(M), _p(Minv), _p(out), B, Ch, H, W)
    return out


def default_M(W, H):
    """M and its inverse as the caller builds them (model_builder...py:196-205): fp32, torch.inverse.
    For the diagonal-plus-shift M the exact inverse is [[1/a,0,-1],[0,1/b,-1],[0,0,1]]."""
    a, b = np.float32(W / 2.0), np.float32(H / 2.0)
    M = np.asarray([[a, 0, a], [0, b, b], [0, 0, 1]], np.float32)
    Minv = np.asarray([[np.float32(1) / a, 0, -1], [0, np.float32(1) / b, -1], [0, 0, 1]], np.float32)
    return M, Minv


def score_argmax(cls, loc, window=None, win_influence=0.0):
    cls, loc = _f32(cls), _f32(loc)
    B, two, N, _ = cls.shape
    assert two == 2
    L = loc.shape[1]
    idx = np.empty(B, np.int64)
    ps = np.empty(B, np.float64)
    sc = np.empty(B, np.float32)
    g = np.empty((B, L), np.float32)
    w = np.ascontiguousarray(window, np.float64) if window is not None else None
    lib().orc_score_argmax(_p(cls), _p(loc), _p(w, _d) if w is not None else None, float(win_influence), _p(idx, _i64), _p(ps, _d),
                           _p(sc), _p(g), B, L, N)
    return idx, ps, sc, g
