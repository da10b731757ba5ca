"""dev: time / profile the large-launch convolution kernels on representative batch-64 backbone layers (HDN_B200_CONV_TS=0|1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hdn_b200 import ops
LAYERS = [  # B, Cin, Cout, HW, k, dil   (512-crop backbone at batch 32)
    (32, 1024, 256, 63, 1, 1), (32, 256, 256, 63, 3, 2), (32, 256, 1024, 63, 1, 1), (32, 512, 512, 63, 3, 4), (32, 1024, 2048, 63, 3, 2), (32, 64, 64, 127, 3, 1),
]
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
for i, (B, Cin, Cout, HW, k, d) in enumerate(LAYERS):
    if only >= 0 and i != only:
        continue
    x = torch.randn(B, Cin, HW, HW, device="cuda"); w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.02
    wt = ops.pack_conv_weight(w)
    sc = torch.ones(Cout, device="cuda"); sh = torch.zeros(Cout, device="cuda")
    for _ in range(2): ops.conv_gemm(x, wt, sc, sh, ksize=k, dilation=d, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 1 if only >= 0 else 5
    e0.record()
    for _ in range(n): ops.conv_gemm(x, wt, sc, sh, ksize=k, dilation=d, relu=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * Cout * Cin * k * k * HW * HW
    print("B%d %4d->%4d %dx%d k%d d%d : %.3f ms  %.1f TFLOP/s fp32-equivalent (%.0f%% of the 3xTF32 tensor peak)" % (B, Cin, Cout, HW, HW, k, d, ms, fl / ms / 1e9, 100 * 3 * fl / ms / 1e9 / 1150))
