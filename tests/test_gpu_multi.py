"""GPU (needs >= 2 devices; skipped otherwise): "N GPUs == 1 GPU" on the CUDA path, launched with torchrun over
NCCL.  The batch is block-split over 2 ranks, the template pack is broadcast from rank 0 (the path's one collective besides the
result gather), and the stitched per-rank results must equal the single-GPU run element-wise -- bit for bit, since every pair is
computed by the same kernels in the same order whatever rank owns it (the worker launches one pair at a time: kernel choice and
split-K cluster size depend on the launch size, so only equal-sized launches are comparable bit for bit)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_equal_one_rank_on_the_cuda_path(tmp_path):
    B = 6
    worker = os.path.join(ROOT, "tests", "multi_gpu_worker.py")
    outs = {}
    for world in (1, 2):
        d = tmp_path / ("w%d" % world)
        d.mkdir()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29620 + world), worker, str(d), str(B)]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        outs[world] = [np.load(str(d / ("rank%d.npz" % r))) for r in range(world)]
    one = outs[1][0]
    two = outs[2]
    assert [(int(p["lo"]), int(p["hi"])) for p in two] == [(0, 3), (3, 6)]
    for k in ("cls", "loc", "cls_lp", "loc_lp", "idx", "idx_lp", "center", "sim_lp", "H"):
        assert np.array_equal(np.concatenate([p[k] for p in two], 0), one[k]), k
    for p in two:  # every rank holds the gathered offsets / H of the whole batch, in rank order
        assert np.array_equal(p["H_all"], one["H_all"]) and np.array_equal(p["offs_all"], one["offs_all"])
