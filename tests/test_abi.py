"""CPU-only: the C-ABI library loads, exports every symbol include/hdn_b200.h declares, and rejects bad
arguments before touching a device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from hdn_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hdn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hdn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "ctypes binding lacks " + n
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_status_strings(lib):
    assert lib.hdn_abi_version() == 1
    assert lib.hdn_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(lib.hdn_status_string(code)) > 3
    assert lib.hdn_launch_count() >= 0


def test_argument_validation_without_device(lib):
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert lib.hdn_xcorr_dw_f32(null, one, one, 1, 4, 8, 8, 3, 3, 0, 36, null) == -1
    assert lib.hdn_xcorr_dw_f32(one, one, one, 1, 4, 8, 8, 9, 3, 0, 108, null) == -2   # kernel larger than input
    assert lib.hdn_xcorr_dw_f32(one, one, one, 0, 4, 8, 8, 3, 3, 0, 36, null) == -2
    assert lib.hdn_xcorr_dw_f32(ctypes.c_void_p(18), one, one, 1, 4, 8, 8, 3, 3, 0, 36, null) == -3
    assert lib.hdn_xcorr_dw_f32(one, one, one, 1, 4, 8, 8, 3, 3, 0, 7, null) == -2     # batch stride smaller than a template
    assert lib.hdn_logpolar_f32(null, null, 0.0, one, 1, 1, 8, 8, 4, null) == -1
    assert lib.hdn_logpolar_f32(one, null, 0.0, one, 1, 1, 8, 8, 1, null) == -2
    assert lib.hdn_dlt4_f32(one, one, null, 1, null) == -1
    assert lib.hdn_dlt4_f32(one, one, one, 0, null) == -2
    assert lib.hdn_homo_warp_f32(one, null, None, None, one, 1, 1, 8, 8, null) == -1
    assert lib.hdn_score_argmax_f32(one, one, null, 0.0, one, one, one, null, 1, 2, 5, null) == -1
    with pytest.raises(ValueError):
        _lib.check(-2, "x")


def test_xcorr_algorithm_selector(lib):
    """AUTO / FFT: the 29x29- and 15x15-template shapes take the transform-domain kernel; DIRECT: none does; bad values are rejected."""
    shapes = ((61, 29, 0), (29, 29, 1), (39, 15, 0))
    for algo, expect in ((0, 1), (2, 1), (1, 0), (0, 1)):
        assert lib.hdn_xcorr_set_algo(algo) == 0
        for Hx, Hk, circ in shapes:
            assert lib.hdn_xcorr_uses_fft(256, Hx, Hx, Hk, Hk, circ, 256 * Hk * Hk) == expect
            assert lib.hdn_xcorr_uses_fft(256, Hx, Hx, Hk, Hk, circ, 0) == expect          # shared template
    assert lib.hdn_xcorr_uses_fft(256, 29, 29, 5, 5, 0, 256 * 25) == 0                     # HBM-bound native shape: direct sum
    assert lib.hdn_xcorr_uses_fft(6, 61, 61, 29, 29, 0, 6 * 841) == 0                      # C % 4 != 0: planes would straddle a window
    assert lib.hdn_xcorr_set_algo(7) == -4


def test_round2_entry_points_validate_without_device(lib):
    """The convolution, spectra and pre-processing entry points reject bad arguments before touching the device."""
    null, one = ctypes.c_void_p(0), ctypes.c_void_p(16)
    # few-channel convolution: supported shapes, NULL, padding beyond k/2, too many input channels
    assert lib.hdn_conv_small_supported(3, 64, 7, 2) == 1 and lib.hdn_conv_small_supported(2, 64, 7, 2) == 1
    assert lib.hdn_conv_small_supported(1, 4, 3, 1) == 1 and lib.hdn_conv_small_supported(8, 1, 3, 1) == 1
    assert lib.hdn_conv_small_supported(16, 4, 3, 1) == 0 and lib.hdn_conv_small_supported(3, 48, 7, 2) == 0 and lib.hdn_conv_small_supported(3, 64, 5, 1) == 0
    assert lib.hdn_conv_small_f32(null, one, None, None, one, 1, 3, 64, 31, 31, 7, 2, 0, 1, null) == -1
    assert lib.hdn_conv_small_f32(one, one, None, None, one, 1, 3, 64, 31, 31, 7, 2, 4, 1, null) == -2
    assert lib.hdn_conv_small_f32(one, one, None, None, one, 1, 16, 4, 31, 31, 3, 1, 1, 1, null) == -4
    assert lib.hdn_conv_small_f32(one, one, None, None, one, 1, 3, 64, 5, 5, 7, 2, 0, 1, null) == -2       # no output pixel
    # tcgen05 convolution: geometry of hdn_conv_gemm_ex_f32
    assert lib.hdn_conv_gemm_supported(256, 256, 3, 2) == 1 and lib.hdn_conv_gemm_supported(64, 64, 3, 1) == 1
    assert lib.hdn_conv_gemm_supported(48, 128, 1, 1) == 0 and lib.hdn_conv_gemm_supported(256, 2, 1, 1) == 0
    assert lib.hdn_conv_gemm_ex_f32(one, one, None, None, None, one, 1, 256, 256, 15, 15, 3, 3, 1, 1, 1, null) == -4   # stride 3
    assert lib.hdn_conv_gemm_ex_f32(one, one, None, None, None, one, 1, 256, 256, 15, 15, 3, 1, 3, 2, 1, null) == -4   # padding > dilation * (k / 2)
    assert lib.hdn_conv_gemm_ex_f32(null, one, None, None, None, one, 1, 256, 256, 15, 15, 3, 1, 1, 1, 1, null) == -1
    for mode in (0, 1, 2, 7, -3):  # clamped to {0, 1, 2}
        assert lib.hdn_conv_gemm_set_ts(mode) == 0
    assert lib.hdn_conv_gemm_set_ts(2) == 0 and lib.hdn_conv_gemm_set_pdl(1) == 0 and lib.hdn_conv_gemm_set_splitk(1) == 0
    # cached template spectra: only the 29x29-template shapes at 256/512 crops have the kernels
    assert lib.hdn_xcorr_spectra_floats(256, 61, 61, 29, 29, 0) == 256 * 29 * 33 * 2
    assert lib.hdn_xcorr_spectra_floats(256, 29, 29, 29, 29, 1) == 256 * 29 * 33 * 2
    assert lib.hdn_xcorr_spectra_floats(256, 29, 29, 5, 5, 0) == 0 and lib.hdn_xcorr_spectra_floats(6, 61, 61, 29, 29, 0) == 0
    arr = (ctypes.c_void_p * 1)
    assert lib.hdn_xcorr_template_spectra_f32(1, arr(16), arr(32), 256, 29, 29, 5, 5, 0, null) == -4
    assert lib.hdn_xcorr_template_spectra_f32(1, arr(16), arr(0), 256, 61, 61, 29, 29, 0, null) == -1
    assert lib.hdn_xcorr_template_spectra_f32(1, arr(20), arr(32), 256, 61, 61, 29, 29, 0, null) == -3         # 16-byte alignment
    assert lib.hdn_xcorr_dw_multi_spec_f32(1, arr(16), arr(32), arr(48), 0, 256, 61, 61, 29, 29, 0, null) == -2
    assert lib.hdn_xcorr_dw_multi_spec_f32(9, arr(16), arr(32), arr(48), 1, 256, 61, 61, 29, 29, 0, null) == -4
    # pre-processing
    assert lib.hdn_warp_perspective_u8(null, one, 8, 8, (ctypes.c_double * 9)(*([1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0])), null) == -1


def test_ops_reject_cpu_tensors():
    import torch
    from hdn_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.xcorr_depthwise(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.DLT_solve(torch.zeros(1, 8), torch.zeros(1, 8))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libhdn_b200.so")
    with pytest.raises(_lib.HdnError, match="no CPU fallback"):
        _lib.lib()


def test_network_shapes_take_the_staged_kernel(lib):
    """The five shapes the network produces (C = 256) must hit the TMA-staged kernels, not the generic fallback."""
    for Hx, Hk, circ in ((29, 5, 0), (13, 13, 1), (61, 29, 0), (29, 29, 1), (39, 15, 0)):
        assert lib.hdn_xcorr_is_staged(256, Hx, Hx, Hk, Hk, circ, 256 * Hk * Hk) == 1, (Hx, Hk, circ)
        assert lib.hdn_xcorr_is_staged(256, Hx, Hx, Hk, Hk, circ, 0) == 1            # shared template
    assert lib.hdn_xcorr_is_staged(256, 31, 31, 5, 5, 0, 256 * 25) == 0               # not in the table
    assert lib.hdn_xcorr_is_staged(20, 29, 29, 5, 5, 0, 20 * 25) == 0                 # C not a multiple of the group
