#!/bin/bash
# build experimental variants of the library (dev tool) and time one conv shape with each
set -e
cd "$(dirname "$0")/../.."
SRC="hdn_b200/csrc/api.cu hdn_b200/csrc/xcorr.cu hdn_b200/csrc/warp.cu hdn_b200/csrc/score.cu hdn_b200/csrc/conv_gemm.cu"
for v in BASE NOFENCE NOMMA ONEMMA; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -DHDN_EXP_$v -o /tmp/lib_$v.so $SRC
done
python - <<'PY'
import ctypes, time, torch
from hdn_b200 import _lib, ops
for v in ["BASE", "NOFENCE", "NOMMA", "ONEMMA"]:
    _lib._lib = None; _lib.SO_PATH = "/tmp/lib_%s.so" % v
    x = torch.randn(1, 512, 31, 31, device="cuda"); w = torch.randn(512, 512, 3, 3, device="cuda") * 0.02
    wt = ops.pack_conv_weight(w)
    for _ in range(3): ops.conv_gemm(x, wt, ksize=3, dilation=4)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): ops.conv_gemm(x, wt, ksize=3, dilation=4)
    torch.cuda.synchronize()
    print(v, "%.3f ms" % ((time.perf_counter() - t) / 20 * 1e3), flush=True)
PY
