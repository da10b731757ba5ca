"""One process per GPU: shard independent pairs / sequences, with the only two exchanges the path has.

The reference's only multi-GPU inference idea is manual sharding of videos by index range per
CUDA_VISIBLE_DEVICES (tools/test.py:90-103, commented out).  Frames of one sequence are serially
dependent through H_total (hdn_tracker_proj_e2e.py:154,262-266), so the unit of sharding is the
sequence / the synthetic pair; there is no halo and no data-path collective.  What is exchanged:
  * ONE broadcast of the template pack from rank 0 when all pairs share a template,
  * ONE all-gather of the per-pair results (corner offsets [B,8] and H [B,9]) per step.
Works on NCCL (GPU) and gloo (CPU tests).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def block_range(n_items, rank, world):
    """Contiguous block split (first n_items % world ranks get one extra) -> [lo, hi)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def round_robin(n_items, rank, world):
    """Sequence ids owned by this rank (config 4: one sequence per GPU, more wrap around)."""
    return list(range(rank, n_items, world))


def broadcast_template_pack(tensors, src=0):
    """Broadcast a list of same-dtype tensors as ONE flat message (the cached template-side correlation
    kernels 6x[256,h,w] + 6x[256,h',w'] + gray template, ~1.3 MB at 127/255).  In place; returns the list."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.broadcast(flat, src=src)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].reshape(t.shape))
        o += n
    return tensors


def gather_results(offsets, H):
    """All-gather per-pair results: offsets [b,8], H [b,3,3] -> ([B,8], [B,3,3]) in rank order (one message)."""
    b = offsets.shape[0]
    packed = torch.cat([offsets.reshape(b, 8), H.reshape(b, 9)], 1).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        full = packed
    else:
        world = dist.get_world_size()
        full = torch.empty((world * b, 17), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(full, packed)
    return full[:, :8], full[:, 8:].reshape(-1, 3, 3)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing rule: report the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
