"""One-hot probes of the tcgen05 conv kernel: where does W[m*,k*] * X[k*,n*] land? (dev tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hdn_b200 import ops
Cin, Cout, H, W = 32, 128, 8, 8
for (m, k, n) in [(0, 0, 0), (1, 0, 0), (8, 0, 0), (0, 1, 0), (0, 4, 0), (0, 8, 0), (0, 0, 1), (0, 0, 4), (0, 0, 5), (5, 9, 7), (77, 21, 35), (127, 31, 63)]:
    x = torch.zeros(1, Cin, H, W, device="cuda"); w = torch.zeros(Cout, Cin, 1, 1, device="cuda")
    x.view(Cin, -1)[k, n] = 1.0
    w[m, k, 0, 0] = 1.0
    out = ops.conv_gemm(x, ops.pack_conv_weight(w), ksize=1).view(Cout, -1)
    nz = out.nonzero().tolist()
    print("probe (m=%d,k=%d,n=%d) -> nonzeros %s vals %s" % (m, k, n, nz[:6], [round(float(out[i, j]), 3) for i, j in nz[:6]]))
# full-K probe: all-ones weights row m, x one-hot (k, n): out[m, n] should be 1 for every k
w = torch.zeros(Cout, Cin, 1, 1, device="cuda"); w[3] = 1.0
for k in (0, 3, 4, 7, 8, 15, 16, 31):
    x = torch.zeros(1, Cin, H, W, device="cuda"); x.view(Cin, -1)[k, 10] = 2.0
    out = ops.conv_gemm(x, ops.pack_conv_weight(w), ksize=1).view(Cout, -1)
    print("k=%d: out[3,10]=%.3f, sum=%.3f nonzeros %s" % (k, float(out[3, 10]), float(out.sum()), out.nonzero().tolist()[:4]))
