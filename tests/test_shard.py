"""CPU, world_size 2 over gloo: the sharding plumbing of hdn_b200/shard.py (block split, template broadcast,
result gather, max-over-ranks) and shard-equivalence of the oracle chain: N ranks == 1 rank, element for element."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from hdn_b200 import shard
    from oracle import torch_port
    r, lr, w = shard.init(backend="gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(1)
    # global problem: 6 pairs, one template shared by all (config 3) -- rank 0 owns the template, others start with garbage
    g = torch.Generator().manual_seed(0)
    B = 6
    x = torch.randn((B, 8, 13, 13), generator=g)
    k = torch.randn((1, 8, 5, 5), generator=g)
    off = torch.rand((B, 8), generator=g) * 16 - 8
    src = torch.tensor([0.0, 0, 0, 127, 127, 127, 127, 0]).repeat(B, 1)
    k_local = k.clone() if rank == 0 else torch.full_like(k, float("nan"))
    gray_t = torch.arange(4.0) if rank == 0 else torch.zeros(4)
    shard.broadcast_template_pack([k_local, gray_t], src=0)
    assert torch.equal(k_local, k) and torch.equal(gray_t, torch.arange(4.0))
    lo, hi = shard.block_range(B, rank, world)
    corr = torch_port.xcorr_depthwise(x[lo:hi], k_local)
    H = torch_port.dlt_solve(src[lo:hi], off[lo:hi]).squeeze(1)
    offs_all, H_all = shard.gather_results(off[lo:hi], H)
    t = shard.max_over_ranks(1.0 + rank, torch.device("cpu"))
    shard.barrier()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), corr=corr.numpy(), lo=lo, hi=hi, offs_all=offs_all.numpy(), H_all=H_all.numpy(), t=t)
    dist.destroy_process_group()


def test_block_and_round_robin_partitions():
    from hdn_b200 import shard
    for n in (1, 7, 8, 64, 512):
        for world in (1, 2, 3, 8):
            spans = [shard.block_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            rr = sorted(i for r in range(world) for i in shard.round_robin(n, r, world))
            assert rr == list(range(n))


def test_two_rank_gloo_equals_single_rank(tmp_path):
    from oracle import torch_port
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    B = 6
    x = torch.randn((B, 8, 13, 13), generator=g)
    k = torch.randn((1, 8, 5, 5), generator=g)
    off = torch.rand((B, 8), generator=g) * 16 - 8
    src = torch.tensor([0.0, 0, 0, 127, 127, 127, 127, 0]).repeat(B, 1)
    torch.set_num_threads(1)
    full = torch_port.xcorr_depthwise(x, k).numpy()
    H_full = torch_port.dlt_solve(src, off).squeeze(1).numpy()
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    stitched = np.concatenate([p["corr"] for p in parts], 0)
    assert np.array_equal(stitched, full)                       # N ranks == 1 rank, bit for bit
    for p in parts:
        assert np.array_equal(p["offs_all"], off.numpy())       # every rank holds the gathered result in rank order
        assert np.allclose(p["H_all"], H_full, rtol=0, atol=0)
        assert float(p["t"]) == 2.0                             # max over ranks of (1 + rank)
    assert [(int(p["lo"]), int(p["hi"])) for p in parts] == [(0, 3), (3, 6)]


def test_tools_import_surface():
    """Names the reference's tools/test.py (:14-19) and tools/demo.py (:15-19) import must exist in the mirror."""
    from hdn_b200 import compat
    compat.activate()
    import importlib
    surface = {
        "hdn.core.config": ["cfg"],
        "hdn.tracker.tracker_builder": ["build_tracker"],
        "hdn.utils.bbox": ["get_axis_aligned_bbox", "get_min_max_bbox", "get_w_h_from_poly", "get_points_from_xyxy", "get_points_from_xywh",
                           "poly2mask", "xywh2xyxy"],
        "hdn.utils.model_load": ["load_pretrain"],
        "toolkit.datasets": ["DatasetFactory"],
        "hdn.models.model_builder_e2e_unconstrained_v2": ["ModelBuilder"],
        "hdn.core.xcorr": ["xcorr_depthwise", "xcorr_depthwise_circular"],
        "homo_estimator.Deep_homography.Oneline_DLTv1.utils": ["DLT_solve", "transform", "transformer"],
        "hdn.models.logpolar": ["STN_Polar", "getPolarImg"],
    }
    for mod, names in surface.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), "%s.%s" % (mod, n)
