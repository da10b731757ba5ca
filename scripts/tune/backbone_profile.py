import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from hdn_b200 import compat, synthetic
compat.activate()
from hdn.core.config import cfg
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cfg.merge_from_file(os.path.join(ROOT, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml"))
from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
model = synthetic.fill_weights(ModelBuilder()).cuda().eval()
torch.backends.cudnn.allow_tf32 = False
x = torch.rand(1, 3, 255, 255, device="cuda") * 255
with torch.no_grad():
    pass
with torch.no_grad():
    for _ in range(3): model.backbone(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        model.backbone(x)
        torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=4, max_name_column_width=40, max_shapes_column_width=70))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=110))
