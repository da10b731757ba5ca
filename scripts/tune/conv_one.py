import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hdn_b200 import ops
x = torch.randn(1, 512, 31, 31, device="cuda"); w = torch.randn(512, 512, 3, 3, device="cuda") * 0.02
wt = ops.pack_conv_weight(w)
for _ in range(3): ops.conv_gemm(x, wt, ksize=3, dilation=4)
torch.cuda.synchronize()
