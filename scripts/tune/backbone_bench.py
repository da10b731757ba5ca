"""cuDNN fp32 (TF32 off) backbone time at small batch: default vs cudnn.benchmark vs channels_last (dev tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hdn_b200 import compat, synthetic

compat.activate()
from hdn.core.config import cfg
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cfg.merge_from_file(os.path.join(ROOT, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml"))
from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder


def t_ms(fn, n=10):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t) / n


model = synthetic.fill_weights(ModelBuilder()).cuda().eval()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for B in (1, 8):
    x = torch.rand(B, 3, 255, 255, device="cuda") * 255
    with torch.no_grad():
        torch.backends.cudnn.benchmark = False
        ref = [f.clone() for f in model.backbone(x)]
        a = t_ms(lambda: model.backbone(x))
        torch.backends.cudnn.benchmark = True
        b = t_ms(lambda: model.backbone(x))
        out = model.backbone(x)
        err = max(float((o - r).abs().max() / r.abs().max()) for o, r in zip(out, ref))
        xcl = x.contiguous(memory_format=torch.channels_last)
        mcl = model.backbone.to(memory_format=torch.channels_last)
        c = t_ms(lambda: mcl(xcl))
        out2 = mcl(xcl)
        err2 = max(float((o - r).abs().max() / r.abs().max()) for o, r in zip(out2, ref))
        model.backbone.to(memory_format=torch.contiguous_format)
    print("B=%d backbone fp32: default %.2f ms | cudnn.benchmark %.2f ms (max rel diff %.2e) | +channels_last %.2f ms (%.2e)" % (B, a, b, err, c, err2), flush=True)
