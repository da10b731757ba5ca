"""Worker of tests/test_gpu_multi.py, launched by torchrun (one rank per GPU, NCCL): config 3 on the CUDA path.
Every rank runs its block of the batch through HeadEngine (fused BAN heads from neck features + K3 + K5/K4 + K6) with the template
pack broadcast from rank 0 by NCCL, gathers offsets + H, and writes its results for the parent test to compare with a 1-rank run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdn_b200 import head_engine as he, shard  # noqa: E402


def main(out_dir, B_total):
    rank, local_rank, world = shard.init()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    host = he.make_inputs("127/255", B_total, seed=7, shared_template=True)  # the same global batch on every rank ...
    lo, hi = shard.block_range(B_total, rank, world)                           # ... of which this rank owns a contiguous block
    up = lambda v: [t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)  # noqa: E731
    zf, zf_lp = up(host["zf"]), up(host["zf_lp"])
    if rank != 0:  # only rank 0 knows the template: everybody else receives it through the ONE broadcast
        for t in zf + zf_lp:
            t.zero_()
    shard.broadcast_template_pack(zf + zf_lp, src=0)
    # chunk = 1: every launch then has the same size whatever the rank count.  (The convolution dispatch -- split-K cluster size, kernel
    # choice -- depends on the launch size, so only equal-sized launches are comparable bit for bit.)
    eng = he.HeadEngine("127/255", hi - lo, dev, chunk=1)
    eng.set_template(zf, zf_lp)
    eng.bind({k: ([t[lo:hi].to(dev) for t in host[k]] if isinstance(host[k], list) else host[k][lo:hi].to(dev)) for k in he.FRAME_KEYS})
    out = eng.run()
    offs, H_all = shard.gather_results(eng.inp["off"], out["H"])
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, H_all=H_all.cpu().numpy(), offs_all=offs.cpu().numpy(),
             **{k: out[k].cpu().numpy() for k in ("cls", "loc", "cls_lp", "loc_lp", "idx", "idx_lp", "center", "sim_lp", "H")})
    shard.barrier()
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
