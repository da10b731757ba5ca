"""dev: per-rank host->device bandwidth with all ranks copying at once (torchrun), to name the limiter of the end-to-end path at N GPUs.
Each rank copies a 512 MB pinned buffer to its GPU 8 times after a barrier; prints per-rank GB/s, the aggregate, and the NUMA node the
GPU reports."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from hdn_b200 import shard  # noqa: E402

rank, local_rank, world = shard.init()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
host = torch.empty(128 * 1024 * 1024, dtype=torch.float32).pin_memory()
dst = torch.empty_like(host, device=dev)
dst.copy_(host, non_blocking=True)
torch.cuda.synchronize()
shard.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    dst.copy_(host, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = 8 * host.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
numa = "?"
try:
    import subprocess
    bus = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
    out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    p = "/sys/bus/pci/devices/%s/numa_node" % out.lower().replace("00000000:", "0000:")
    numa = open(p).read().strip() if os.path.exists(p) else "?"
except Exception:
    pass
t = torch.tensor([gbs], dtype=torch.float64, device=dev)
allv = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    import torch.distributed as dist
    dist.all_gather(allv, t)
else:
    allv = [t]
if rank == 0:
    vals = [float(v.item()) for v in allv]
    print(json.dumps({"world": world, "h2d_gbs_per_rank": vals, "aggregate_gbs": sum(vals), "rank0_gpu_numa_node": numa,
                      "host_threads": len(os.sched_getaffinity(0))}), flush=True)
shard.barrier()
