"""Run the mirrored packages (hdn_b200/compat) on the CPU by backing their operator calls with oracle/torch_port.py.

TEST / BASELINE INFRASTRUCTURE ONLY.  The product has no CPU path (hdn_b200.ops rejects CPU tensors).  bench.py's CPU arm needs
"the reference's own CPU forward" at model and tracker level on the GPU box, where /root/reference does not exist; what the
reference executes there is: torch convolutions / BatchNorm (oneDNN), grouped F.conv2d for the correlations, F.grid_sample,
torch.inverse + gathers for DLT / warp, NumPy soft-max / arg-max, OpenCV on the host.  `install()` rebinds the operator names of
the mirror to those very calls (oracle/torch_port.py, pinned to the reference's outputs by tests/test_oracle_golden.py), so that
ModelBuilder / hdnTrackerHomo from hdn_b200/compat run end to end on CPU tensors with the reference's arithmetic.
Only bench.py (`--impl reference`, `cpu_baseline`) and tests may call this.
"""
import numpy as np
import torch

from . import torch_port as tp


def _packed(cls, loc, window=None, win_influence=0.0):
    """CPU stand-in of ops.score_argmax_packed: same byte layout (idx i64 | pscore f64 | score f32 | loc[:, idx] f32)."""
    win = window.detach().cpu().numpy() if isinstance(window, torch.Tensor) else window
    idx, ps, sc, g = tp.score_argmax(cls.detach().cpu(), loc.detach().cpu(), win, win_influence)
    raw = idx.astype("<i8").tobytes() + ps.astype("<f8").tobytes() + sc.astype("<f4").tobytes() + np.ascontiguousarray(g, "<f4").tobytes()
    raw += b"\0" * ((-len(raw)) % 8)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8)


def _dlt_warp(src_p, off_set, I1, M=None, M_inv=None):
    Hm = tp.dlt_solve(src_p, off_set).squeeze(1)
    Mt = torch.tensor(M, dtype=torch.float32).reshape(3, 3) if M is not None else None
    Mi = torch.tensor(M_inv, dtype=torch.float32).reshape(3, 3) if M_inv is not None else None
    return Hm, tp.homo_warp(I1, Hm, Mt, Mi)


def _multi(xs, kernels, circular=False, outs=None):
    f = tp.xcorr_depthwise_circular if circular else tp.xcorr_depthwise
    return [f(x, k) for x, k in zip(xs, kernels)]


def install():
    """Idempotent.  After this the mirror's models / trackers run on CPU tensors (cfg.CUDA must be False)."""
    from hdn_b200 import compat, ops
    compat.activate()
    import hdn.core.xcorr as cx
    import hdn.models.head.ban as ban
    import hdn.models.head.ban_lp as ban_lp
    patches = {"xcorr_depthwise": lambda x, k, out=None: tp.xcorr_depthwise(x, k),
               "xcorr_depthwise_circular": lambda x, k, out=None: tp.xcorr_depthwise_circular(x, k), "xcorr_depthwise_multi": _multi,
               "logpolar_sample": lambda x, polar=None, rot_delta=0.0, out_size=None, out=None: tp.logpolar(x, polar, rot_delta, int(out_size)),
               "DLT_solve": tp.dlt_solve, "dlt_warp": _dlt_warp, "score_argmax_packed": _packed,
               "score_argmax_host": lambda cls, loc, window=None, w=0.0: ops.unpack_scores(_packed(cls, loc, window, w).numpy(), cls.shape[0], loc.shape[1])}
    for name, fn in patches.items():
        setattr(ops, name, fn)
    for mod in (cx, ban, ban_lp):
        for name in ("xcorr_depthwise", "xcorr_depthwise_circular", "xcorr_depthwise_multi"):
            if hasattr(mod, name):
                setattr(mod, name, patches[name])
    ban.DepthwiseXCorr.correlate = staticmethod(patches["xcorr_depthwise"])
    ban_lp.DepthwiseXCorrCirc.correlate = staticmethod(patches["xcorr_depthwise_circular"])
    return True


def build_cpu_model(instance=255, exemplar=127, variant=None):
    """The mirrored ModelBuilder on the CPU with the seeded weight fixture."""
    import os
    from hdn_b200 import synthetic
    install()
    from hdn.core.config import cfg
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg.merge_from_file(os.path.join(root, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml"))
    cfg.TRACK.INSTANCE_SIZE, cfg.TRACK.EXEMPLAR_SIZE = instance, exemplar
    cfg.CUDA = False
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    return synthetic.fill_weights(ModelBuilder(), variant=variant).eval(), cfg
