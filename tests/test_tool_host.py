"""CPU: the reference's benchmark runner, tools/test.py, executed UNCHANGED against the mirrored packages (hdn_b200.run_tool)
as far as a machine without a GPU can take it: argument parsing, every `from hdn... / toolkit... import` of the tool, the merge of
the reference's own full YAML, ModelBuilder() and load_pretrain() of a reference-format checkpoint (836 tensors) all run; the
first `.cuda()` (tools/test.py:69) is where this box stops.  Needs the reference checkout; skipped where it is absent (GPU box).
The GPU side of the same loop is tests/test_gpu_tool.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("HDN_REFERENCE_ROOT", "/root/reference")
TOOL = os.path.join(REF, "tools", "test.py")

DRIVER = r'''
import os, sys, torch
sys.path.insert(0, %(root)r)
from hdn_b200 import compat, run_tool, synthetic
compat.activate()
from hdn.core.config import cfg
cfg.merge_from_file(%(yaml)r)
from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
ckpt = os.path.join(%(tmp)r, "hdn_fixture.pth")
torch.save({"state_dict": {"module." + k: v for k, v in synthetic.fill_weights(ModelBuilder()).state_dict().items()}}, ckpt)  # DataParallel-style checkpoint

class ReachedCuda(Exception):
    pass
def stop(self, *a, **k):
    n = len(self.state_dict())
    raise ReachedCuda("model with %%d tensors reached .cuda()" %% n)
torch.nn.Module.cuda = stop
try:
    run_tool.main([%(tool)r, "--dataset", "POT210", "--config", %(yaml)r, "--snapshot", ckpt])
except ReachedCuda as e:
    import hdn, toolkit
    assert hdn.__file__.startswith(%(root)r) and toolkit.__file__.startswith(%(root)r), (hdn.__file__, toolkit.__file__)
    print("OK", e)
'''


@pytest.mark.skipif(not os.path.isfile(TOOL), reason="reference checkout not present")
def test_reference_test_tool_runs_unchanged_up_to_the_first_cuda_call(tmp_path):
    yaml = os.path.join(REF, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml")  # the reference's own full YAML
    code = DRIVER % {"root": ROOT, "yaml": yaml, "tmp": str(tmp_path), "tool": TOOL}
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=str(tmp_path),
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "OK model with" in res.stdout and "tensors reached .cuda()" in res.stdout, res.stdout[-2000:]
