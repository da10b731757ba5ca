"""CPU-only: the arithmetic of the FFT correlation kernel (hdn_b200/csrc/fft64.cuh, xcorr_fft.cuh), compiled for the host.

tests/native/host_fft_check.cpp runs the kernel's five phases (rows / three column passes / outputs) task by task, in the order the
kernel's barriers impose, on one group of planes per shape, and compares with a direct double-precision correlation
(hdn/core/xcorr.py:37-61 semantics, including the circular-row / replicate-column padding of K2).  The GPU parity tests
(-m gpu) check the same code as it runs on the device.  The *_spectra cases run the shared-template path: the template's row spectra taken once
(phase R on an all-zero x), then the KSPEC configuration's phases with those spectra given.
"""
import os
import subprocess

from conftest import ROOT


def test_fft_correlation_arithmetic_on_host(tmp_path):
    exe = str(tmp_path / "host_fft_check")
    src = os.path.join(ROOT, "tests", "native", "host_fft_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = [l.split() for l in res.stdout.strip().splitlines()]
    assert [l[0] for l in lines] == ["k1_256", "k2_256", "win15", "ragged", "ragged_circ", "k1_256_spectra", "k2_256_spectra", "k1_256_spectra_pipe"]
    for name, _, err, _, ref, _, order_free in lines:
        assert float(err) <= 2e-6 * float(ref), (name, err, ref)  # fp32-accurate: ~3e-7 of max|out|
        assert order_free == "1", name  # the result does not depend on the task order inside a phase (no intra-phase race)
