"""Dilated 3x3 conv at B=1: cuDNN vs 9 shifted cuBLAS GEMMs on the padded, flattened plane (dev tool)."""
import time
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def t_ms(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t) / n


def shifted_gemm_conv(x, wt, d):
    """x [B,Cin,H,W], wt [9,Cout,Cin] (tap-major), dilation = padding = d, stride 1 -> [B,Cout,H,W]."""
    B, Cin, H, W = x.shape
    Wp = W + 2 * d
    xp = F.pad(x, (d, d, d, d)).reshape(B, Cin, -1)
    L = (H - 1) * Wp + W
    out = None
    for ky in range(3):
        for kx in range(3):
            off = ky * d * Wp + kx * d
            xs = xp[:, :, off:off + L]
            w = wt[ky * 3 + kx]
            out = torch.matmul(w, xs) if out is None else torch.baddbmm(out, w.expand(B, -1, -1), xs)
    full = torch.empty((B, wt.shape[1], H * Wp), device=x.device, dtype=x.dtype)
    full[:, :, :L] = out
    return full.view(B, -1, H, Wp)[:, :, :, :W]


for (B, C, Co, S, d) in [(1, 512, 512, 31, 4), (1, 512, 512, 31, 2), (1, 256, 256, 31, 2), (8, 512, 512, 31, 4), (1, 1024, 2048, 31, 2), (1, 512, 512, 15, 4)]:
    x = torch.randn(B, C, S, S, device="cuda")
    w = torch.randn(Co, C, 3, 3, device="cuda") * 0.02
    wt = w.permute(2, 3, 0, 1).reshape(9, Co, C).contiguous()
    ref = F.conv2d(x, w, padding=d, dilation=d)
    got = shifted_gemm_conv(x, wt, d)
    err = float((got - ref).abs().max() / ref.abs().max())
    a = t_ms(lambda: F.conv2d(x, w, padding=d, dilation=d))
    b = t_ms(lambda: shifted_gemm_conv(x, wt, d))
    print("B=%d %4d->%4d %dx%d dil %d: cuDNN %.3f ms | 9 shifted GEMMs %.3f ms | max rel diff %.2e" % (B, C, Co, S, S, d, a, b, err), flush=True)
