"""ctypes binding of libhdn_b200.so (the C ABI in include/hdn_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a call returns a
non-zero status this module raises.  (The oracle under oracle/ is test infrastructure and is
never imported from here.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("HDN_B200_LIB") or os.path.join(HERE, "libhdn_b200.so")  # override: A/B builds of the same ABI (dev)

# name -> (restype, argtypes); mirrors include/hdn_b200.h one to one.
_vp, _ci, _i64, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double
SIGNATURES = {
    "hdn_abi_version": (_ci, []),
    "hdn_status_string": (ctypes.c_char_p, [_ci]),
    "hdn_device_info": (_ci, [ctypes.POINTER(_ci)] * 3),
    "hdn_launch_count": (_i64, []),
    "hdn_xcorr_dw_f32": (_ci, [_vp, _vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _i64, _vp]),
    "hdn_xcorr_dw_multi_f32": (_ci, [_ci, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _ci, _ci, _ci, _ci, _ci, _ci, _ci,
                                     _i64, _vp]),
    "hdn_xcorr_set_algo": (_ci, [_ci]),
    "hdn_xcorr_uses_fft": (_ci, [_ci, _ci, _ci, _ci, _ci, _ci, _i64]),
    "hdn_xcorr_is_staged": (_ci, [_ci, _ci, _ci, _ci, _ci, _ci, _i64]),
    "hdn_xcorr_generic_launches": (_i64, []),
    "hdn_logpolar_f32": (_ci, [_vp, _vp, _f, _vp, _ci, _ci, _ci, _ci, _ci, _vp]),
    "hdn_logpolar_u8": (_ci, [_vp, _vp, _f, _vp, _ci, _ci, _ci, _ci, _ci, _vp]),
    "hdn_dlt4_f32": (_ci, [_vp, _vp, _vp, _ci, _vp]),
    "hdn_homo_warp_f32": (_ci, [_vp, _vp, ctypes.POINTER(_f), ctypes.POINTER(_f), _vp, _ci, _ci, _ci, _ci, _vp]),
    "hdn_dlt_warp_f32": (_ci, [_vp, _vp, _vp, ctypes.POINTER(_f), ctypes.POINTER(_f), _vp, _vp, _ci, _ci, _ci, _ci, _vp]),
    "hdn_conv_gemm_supported": (_ci, [_ci, _ci, _ci, _ci]),
    "hdn_conv_gemm_set_splitk": (_ci, [_ci]),
    "hdn_conv_gemm_set_pdl": (_ci, [_ci]),
    "hdn_conv_gemm_set_ts": (_ci, [_ci]),
    "hdn_conv_gemm_set_shift": (_ci, [_ci]),
    "hdn_conv_pack_weight_f32": (_ci, [_vp, _vp, _ci, _ci, _vp]),
    "hdn_conv_gemm_f32": (_ci, [_vp, _vp, _vp, _vp, _vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _vp]),
    "hdn_score_argmax_f32": (_ci, [_vp, _vp, _vp, _d, _vp, _vp, _vp, _vp, _ci, _ci, _ci, _vp]),
    "hdn_warp_perspective_u8": (_ci, [_vp, _vp, _ci, _ci, ctypes.POINTER(_d), _vp]),
    "hdn_cubic_table_host": (_ci, [_vp]),
    "hdn_warp_affine_cubic_u8": (_ci, [_vp, _vp, _ci, _ci, ctypes.POINTER(_d), _vp, _vp]),
    "hdn_crop_resize_u8": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, ctypes.POINTER(ctypes.c_uint8), _ci, _ci, ctypes.POINTER(_d), ctypes.POINTER(_d), _vp, _vp]),
    "hdn_xcorr_spectra_floats": (ctypes.c_int64, [_ci] * 6),
    "hdn_xcorr_template_spectra_f32": (_ci, [_ci, ctypes.POINTER(_vp), ctypes.POINTER(_vp)] + [_ci] * 6 + [_vp]),
    "hdn_xcorr_dw_multi_spec_f32": (_ci, [_ci] + [ctypes.POINTER(_vp)] * 3 + [_ci] * 7 + [_vp]),
    "hdn_conv_small_supported": (_ci, [_ci, _ci, _ci, _ci]),
    "hdn_conv_small_f32": (_ci, [_vp] * 5 + [_ci] * 9 + [_vp]),
    "hdn_conv_gemm_ex_f32": (_ci, [_vp, _vp, _vp, _vp, _vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _vp]),
    "hdn_conv_gemm_multi_f32": (_ci, [_ci] + [ctypes.POINTER(_vp)] * 5 + [_ci] * 9 + [_vp]),
    "hdn_head_project_multi_f32": (_ci, [_ci] + [ctypes.POINTER(_vp)] * 6 + [_ci] * 5 + [_vp]),
    "hdn_head_score_f32": (_ci, [_ci, _ci] + [ctypes.POINTER(_vp)] * 4 + [ctypes.POINTER(_f)] * 3 + [_vp, _vp, _vp, _d, _vp, _vp, _vp, _vp, _ci, _ci, _ci, _vp]),
}

_lib = None


class HdnError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise HdnError("libhdn_b200.so is not built (%s). Run `python -m hdn_b200.build`; there is no CPU fallback." % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the binding drift apart
            fn.restype, fn.argtypes = res, args
        if "HDN_B200_CONV_TS" in os.environ:  # A/B switch of the large-launch convolution kernel (see hdn_conv_gemm_set_ts)
            L.hdn_conv_gemm_set_ts(int(os.environ["HDN_B200_CONV_TS"]))
        _lib = L
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().hdn_status_string(status).decode()
        if status < 0:
            raise ValueError("%s: %s" % (what, msg))
        raise HdnError("%s: CUDA error %d (%s)" % (what, status, msg))


def launch_count():
    return int(lib().hdn_launch_count())
