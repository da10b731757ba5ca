"""Time the three network stages at several batch sizes (dev tool): eager vs CUDA graph, TF32 off vs on."""
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
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t) / n


for tf32 in (0, 1):
    os.environ["HDN_B200_TF32"] = str(tf32)
    model = synthetic.fill_weights(ModelBuilder()).cuda().eval()
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    for B in (1, 4, 8, 16, 32):
        z = torch.rand(B, 6, 127, 127, device="cuda") * 255
        x = torch.rand(B, 3, 255, 255, device="cuda") * 255
        pair = torch.randn(B, 2, 127, 127, device="cuda")
        h4p = torch.tensor([[0.0, 0, 0, 127, 127, 127, 127, 0]], device="cuda").repeat(B, 1)
        model.enable_graphs(False)
        model.template(z)
        with torch.no_grad():
            bb = t_ms(lambda: model.backbone(x))
        e1 = t_ms(lambda: model.track_new_scored(x))
        e2 = t_ms(lambda: model.track_new_lp_scored(x))
        e3 = t_ms(lambda: model.track_proj_packed(pair, h4p))
        model.enable_graphs(True)
        g1 = t_ms(lambda: model.track_new_scored(x))
        g2 = t_ms(lambda: model.track_new_lp_scored(x))
        g3 = t_ms(lambda: model.track_proj_packed(pair, h4p))
        print("tf32=%d B=%2d  backbone255 %.2f | stage1 eager %.2f graph %.2f | stage2 eager %.2f graph %.2f | stage3 eager %.2f graph %.2f  -> %.2f ms/frame-equivalent (graph)" % (
            tf32, B, bb, e1, g1, e2, g2, e3, g3, (g1 + g2 + g3) / B), flush=True)
