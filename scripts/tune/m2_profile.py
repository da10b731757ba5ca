"""dev: torch-profiler kernel table of one full forward (M2) at batch B, 256/512 or 127/255 crops."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from hdn_b200 import compat, synthetic  # noqa: E402

compat.activate()
from hdn.core.config import cfg  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "256/512"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg.merge_from_file(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml"))
ex, inst = (256, 512) if wl == "256/512" else (127, 255)
cfg.TRACK.INSTANCE_SIZE, cfg.TRACK.EXEMPLAR_SIZE = inst, ex
cfg.CUDA = True
from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder  # noqa: E402

model = synthetic.fill_weights(ModelBuilder()).cuda().eval()
g = torch.Generator(device="cuda").manual_seed(5)
z = torch.rand((B, 6, ex, ex), device="cuda", generator=g) * 255.0
x = torch.rand((B, 3, inst, inst), device="cuda", generator=g) * 255.0
pair = torch.randn((B, 2, 127, 127), device="cuda", generator=g)
h4p = torch.tensor([[0.0, 0.0, 0.0, 127.0, 127.0, 127.0, 127.0, 0.0]], device="cuda").repeat(B, 1)
with torch.no_grad():
    model.template(z)

    def frame():
        model._stage1_packed(x, cfg.TRACK.WINDOW_INFLUENCE)
        model._stage2_packed(x)
        model._stage3_packed(pair, h4p)
    for _ in range(2):
        frame()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        frame()
        torch.cuda.synchronize()
print("M2 %s batch %d" % (wl, B))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
