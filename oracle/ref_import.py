"""Import the UNMODIFIED reference (zhanxinrui/HDN at /root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (and nothing in the
product) to run the reference's own Python on CPU so that golden vectors can be
committed under tests/golden/.  /root/reference does not exist on the GPU box,
so nothing that runs there may import this module.

What is patched, and why:
  * sys.path: /root/reference first, oracle/refshim (yacs/matplotlib/... stubs)
    second; the repo root is REMOVED so that `import hdn` resolves to the
    reference, not to hdn_b200/compat/hdn.
  * numpy: `np.float` alias (hdn/utils/transform.py uses the removed alias).
  * CPU mode: `Tensor.cuda`, `Module.cuda` -> identity and
    `torch.cuda.is_available` -> True, because the reference calls `.cuda()`
    unconditionally (logpolar.py:110-111, model_builder...py:100,146) and
    track_proj only binds batch_indices_tensor under `is_available()`
    (model_builder_e2e_unconstrained_v2.py:176-178).
  * model_zoo.load_url -> never called: HomoModelBuilder(pretrained=True) is
    forced to pretrained=False (no network).
"""
import os
import sys

REF_ROOT = os.environ.get("HDN_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
_installed = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "hdn"))


def install():
    """Idempotent.  After this, `import hdn...` / `import homo_estimator...` hit the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    import numpy as np
    import torch

    # 1. path surgery: the reference must win over our compat mirror.
    bad = {os.path.realpath(p) for p in (_REPO, os.path.join(_REPO, "hdn_b200", "compat"))}
    sys.path[:] = [p for p in sys.path if os.path.realpath(p or os.getcwd()) not in bad]
    for m in list(sys.modules):
        if m == "hdn" or m.startswith("hdn.") or m == "homo_estimator" or m.startswith("homo_estimator."):
            del sys.modules[m]
    sys.path.insert(0, os.path.join(_HERE, "refshim"))
    sys.path.insert(0, REF_ROOT)

    # 2. numpy alias removed in 1.24
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int

    # 3. CPU mode
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    torch.cuda.current_device = lambda: 0
    for storage in (getattr(torch, "UntypedStorage", None), getattr(torch.storage, "TypedStorage", None)):
        if storage is not None:  # load_pretrain maps checkpoints with `storage.cuda(device)` (hdn/utils/model_load.py:50-53)
            storage.cuda = lambda self, *a, **k: self

    # 4. no downloads
    import torch.utils.model_zoo as model_zoo

    def _no_net(*a, **k):
        raise RuntimeError("offline: model_zoo.load_url blocked")

    model_zoo.load_url = _no_net
    _installed = True


def force_homo_backbone_offline():
    """HomoModelBuilder(pretrained=True) (model_builder...py:67) would hit the network."""
    install()
    import homo_estimator.Deep_homography.Oneline_DLTv1.backbone as hb
    import homo_estimator.Deep_homography.Oneline_DLTv1.models.homo_model_builder as hmb

    orig = hb.get_backbone

    def offline(name, pretrained=False, **kw):
        return orig(name, pretrained=False, **kw)

    hb.get_backbone = offline
    hmb.get_backbone = offline
