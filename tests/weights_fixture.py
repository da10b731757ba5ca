"""Deterministic weight fixture shared by oracle/gen_golden_model.py (which loads it into the REFERENCE model)
and tests/ (which load it into hdn_b200's mirror).  No checkpoint ships with the reference and an 84 M-parameter
file cannot be committed, so every tensor is regenerated from a seed derived from its state-dict name.

Values are chosen so activations stay O(1)-O(100) through ~50 layers in eval mode (kaiming fan-in convolutions,
BatchNorm with statistics near identity, the BN that closes each residual block damped), and the last layer of each
head is scaled (SCALES, calibrated once against the reference) so logits / offsets have a useful dynamic range.
"""
import math
import zlib

import torch
import torch.nn as nn

# final-layer scale factors, calibrated by `python oracle/gen_golden_model.py --calibrate` (printed there)
SCALES = {"head_cls": 1.218e-06, "head_loc": 1.297e-06, "head_lp_cls": 7.473e-08, "head_lp_loc": 2.397e-08, "fc": 1.351}


def _gen(name):
    return torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)


def _randn(name, shape):
    return torch.randn(tuple(shape), generator=_gen(name))


def fill(model, scales=None):
    """In-place deterministic initialisation of every parameter and buffer of `model` (eval-mode fixture)."""
    scales = dict(SCALES, **(scales or {}))
    with torch.no_grad():
        for mname, m in model.named_modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.in_channels // m.groups * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(_randn(mname + ".weight", m.weight.shape) * math.sqrt(2.0 / fan_in))
                if m.bias is not None:
                    m.bias.copy_(_randn(mname + ".bias", m.bias.shape) * 0.1)
            elif isinstance(m, nn.BatchNorm2d):
                closing = mname.endswith(".bn3") or (mname.endswith(".bn2") and "hm_net" in mname)
                gamma = 0.3 if closing else 1.0
                m.weight.copy_(gamma * (1.0 + 0.1 * _randn(mname + ".weight", m.weight.shape)))
                m.bias.copy_(0.05 * _randn(mname + ".bias", m.bias.shape))
                m.running_mean.copy_(0.05 * _randn(mname + ".running_mean", m.running_mean.shape))
                m.running_var.copy_(1.0 + 0.1 * torch.rand(tuple(m.running_var.shape), generator=_gen(mname + ".running_var")))
                m.num_batches_tracked.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.copy_(_randn(mname + ".weight", m.weight.shape) * math.sqrt(1.0 / m.in_features))
                m.bias.copy_(_randn(mname + ".bias", m.bias.shape))
        for pname, p in model.named_parameters():
            if pname.endswith(("cls_weight", "loc_weight", "loc_scale")):
                p.copy_(1.0 + 0.2 * _randn(pname, p.shape))
        sd = model.state_dict()
        for key, t in sd.items():
            for prefix, tag in (("head.", "head"), ("head_lp.", "head_lp")):
                if key.startswith(prefix) and ".head.3." in key:
                    branch = "cls" if ".cls." in key else "loc"
                    t.mul_(scales["%s_%s" % (tag, branch)])
            if key.startswith("hm_net.fc."):
                t.mul_(scales["fc"])
    model.eval()
    return model
