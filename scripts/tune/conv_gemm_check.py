"""Correctness + timing of the tcgen05 3xTF32 implicit-GEMM convolution against cuDNN fp32 (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from hdn_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def t_ms(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t) / n


cases = [(1, 64, 128, 9, 9, 1, 1), (1, 512, 512, 31, 31, 3, 4), (1, 1024, 256, 31, 31, 1, 1), (1, 512, 2048, 31, 31, 1, 1), (2, 256, 256, 15, 15, 3, 2),
         (1, 256, 1024, 31, 31, 1, 1), (8, 512, 512, 31, 31, 3, 4), (1, 1024, 2048, 31, 31, 3, 2), (1, 2048, 256, 31, 31, 1, 1), (1, 128, 512, 63, 63, 1, 1)]
for (B, Cin, Cout, H, W, k, d) in cases:
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Cin + Cout)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = 1 + 0.1 * torch.randn(Cout, device="cuda", generator=g)
    shift = 0.1 * torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(B, Cout, H, W, device="cuda", generator=g)
    ref = F.relu(F.conv2d(x, w, padding=d * (k // 2), dilation=d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    ref64 = F.relu(F.conv2d(x.double(), w.double(), padding=d * (k // 2), dilation=d) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1) + res.double())
    wt = ops.pack_conv_weight(w)
    got = ops.conv_gemm(x, wt, scale, shift, res, ksize=k, dilation=d, relu=True)
    torch.cuda.synchronize()
    den = float(ref64.abs().max())
    e_ours = float((got.double() - ref64).abs().max()) / den
    e_cudnn = float((ref.double() - ref64).abs().max()) / den
    a = t_ms(lambda: F.conv2d(x, w, padding=d * (k // 2), dilation=d))
    b = t_ms(lambda: ops.conv_gemm(x, wt, scale, shift, res, ksize=k, dilation=d, relu=True))
    fl = 2.0 * B * Cout * Cin * k * k * H * W
    print("B=%d %4d->%4d %2dx%2d k%d d%d: err vs fp64 ours %.2e cudnn %.2e | cuDNN conv only %.3f ms | tcgen05 conv+bn+res+relu %.3f ms (%.1f TFLOP/s)" % (
        B, Cin, Cout, H, W, k, d, e_ours, e_cudnn, a, b, fl / b / 1e9), flush=True)
