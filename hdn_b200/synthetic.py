"""Seeded synthetic data: inputs, planar-object sequences and a deterministic weight fixture.

Shared by the golden generator (which loads this file BY PATH into a process where `hdn` is the reference), the
tests, bench.py and hdn_b200.runner.  No checkpoint ships with the reference and there is no network, so weights
are regenerated from a seed derived from each state-dict name (`fill_weights`).  Pure NumPy / OpenCV / torch.
"""
import math
import zlib

import cv2
import numpy as np
import torch
import torch.nn as nn


def crop_tensor(seed, shape):
    """A smooth-ish random BGR crop in [0,255] as float32 NCHW (what get_subwindow would hand the model)."""
    rng = np.random.default_rng(seed)
    b, c, h, w = shape
    coarse = rng.random((b, c, h // 8 + 2, w // 8 + 2)).astype(np.float32)
    out = np.empty(shape, np.float32)
    for i in range(b):
        for j in range(c):
            up = cv2.resize(coarse[i, j], (w, h), interpolation=cv2.INTER_CUBIC)
            out[i, j] = np.clip(up * 200.0 + 27.0 + rng.standard_normal((h, w)).astype(np.float32) * 6.0, 0, 255)
    return out


def texture(seed, h, w):
    """A planar texture with structure at several scales (uint8 BGR)."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 3), np.float32)
    for cells in (4, 9, 19, 41):
        layer = rng.random((cells, cells, 3)).astype(np.float32)
        img += cv2.resize(layer, (w, h), interpolation=cv2.INTER_NEAREST if cells > 15 else cv2.INTER_CUBIC) / 4.0
    return np.clip(img * 255.0, 0, 255).astype(np.uint8)


def sequence(seed, n_frames, size=(360, 480), obj=(120, 160), events=None):
    """A planar object on a background, moved by a smooth random homography walk.
    -> frames [n] uint8 BGR, gt polygons [n,8] (x1,y1,..,x4,y4 TL,TR,BR,BL of the object).
    events: {frame index: 'occlude' | 'invert' | 'flat'} -- that frame shows an unrelated texture / the photometric negative /
    a constant colour instead (the frames that trip the tracker's pscore / lp-score / homo_score gates)."""
    rng = np.random.default_rng(seed)
    H_img, W_img = size
    oh, ow = obj
    bg = texture(seed + 1, H_img, W_img)
    fg = texture(seed + 2, oh, ow)
    x0, y0 = (W_img - ow) / 2.0, (H_img - oh) / 2.0
    base = np.array([[x0, y0], [x0 + ow, y0], [x0 + ow, y0 + oh], [x0, y0 + oh]], np.float32)
    frames, polys = [], []
    cur = base.copy()
    vel = np.zeros((4, 2), np.float32)
    src = np.array([[0, 0], [ow, 0], [ow, oh], [0, oh]], np.float32)
    for t in range(n_frames):
        if t > 0:
            vel = 0.7 * vel + rng.normal(0, 0.9, (4, 2)).astype(np.float32) + rng.normal(0, 1.2, (1, 2)).astype(np.float32)
            cur = cur + vel
        Hm = cv2.getPerspectiveTransform(src, cur)
        frame = bg.copy()
        warped = cv2.warpPerspective(fg, Hm, (W_img, H_img))
        mask = cv2.warpPerspective(np.full((oh, ow), 255, np.uint8), Hm, (W_img, H_img))
        frame[mask > 127] = warped[mask > 127]
        kind = (events or {}).get(t)
        if kind == "occlude":
            frame = texture(seed + 1000 + t, H_img, W_img)
        elif kind == "invert":
            frame = 255 - frame
        elif kind == "flat":
            frame = np.full_like(frame, 128)
        elif kind is not None:
            raise ValueError("unknown event %r" % (kind,))
        frames.append(frame)
        polys.append(cur.reshape(-1).copy())
    return frames, np.asarray(polys, np.float32)


# ------------------------------------------------------------------------------ deterministic weights
# Every tensor is regenerated from a seed derived from its state-dict name.  Values keep activations O(1)-O(100)
# through ~50 layers in eval mode (kaiming fan-in convolutions, BatchNorm near identity, the BN that closes each
# residual block damped); the last layer of each head is scaled (SCALES, calibrated once against the reference with
# `python oracle/gen_golden_model.py --calibrate`) so logits / offsets have a useful dynamic range.

# final-layer scale factors, calibrated by `python oracle/gen_golden_model.py --calibrate` (printed there)
SCALES = {"head_cls": 1.218e-06, "head_loc": 1.297e-06, "head_lp_cls": 7.473e-08, "head_lp_loc": 2.397e-08, "fc": 1.351}


def _gen(name):
    return torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)


def _randn(name, shape):
    return torch.randn(tuple(shape), generator=_gen(name))


# Second calibration ("gates" variant): the classification heads respond MONOTONICALLY to the correlation strength (final 1x1
# weights +1 on the foreground row, -1 on the background row, plus a bias: logit ~ the norm of the correlation vector), so the log-polar peak sits at the
# auto-correlation maximum with a confident score, and a frame whose content does not match the template drops below the
# tracker's gates (pscore < 0.05, lp score < 0.25) while a photometric mismatch pushes homo_score past 2.5.  Values are
# calibrated against the reference by `python oracle/gen_golden_model.py --calibrate-gates` (printed there).
GATES = {"head_cls": 1e-8, "head_cls_bias": 13.0, "head_loc": 1.297e-06, "head_lp_cls": 8.6e-9, "head_lp_cls_bias": 45.5, "head_lp_loc": 3e-9,
         "head_lp_loc_bias0": -1.0, "fc": 1.351, "share_gain": 19.93}


def fill_weights(model, scales=None, variant=None):
    """In-place deterministic initialisation of every parameter and buffer of `model` (eval-mode fixture).
    variant='gates': the second calibration described above (scales default to GATES)."""
    scales = dict(GATES if variant == "gates" else SCALES, **(scales or {}))
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
                    if variant == "gates" and branch == "cls":
                        if key.endswith("weight"):  # row 1 = foreground: +1, row 0 = background: -1
                            t.fill_(1.0)
                            t[0].fill_(-1.0)
                        else:  # bias: the background logit carries the offset
                            t.zero_()
                            t[0] = scales["%s_cls_bias" % tag]
                    if variant == "gates" and tag == "head_lp" and branch == "loc" and key.endswith("bias"):
                        # constant log-radius offset: cancels the one-cell bias of the untrained log-polar peak (scale stays near 1)
                        level = int(key.split(".box")[1][0]) - 2
                        t.zero_()
                        t[0] = scales["head_lp_loc_bias0"] / float(sd["head_lp.loc_scale"][level])
                        continue
                    t.mul_(scales["%s_%s" % (tag, branch)] if not (variant == "gates" and branch == "cls" and key.endswith("bias")) else 1.0)
            if key.startswith("hm_net.fc."):
                t.mul_(scales["fc"])
            if variant == "gates" and key in ("hm_net.ShareFeature.ShareFeature.7.weight", "hm_net.ShareFeature.ShareFeature.7.bias"):
                t.mul_(scales["share_gain"])  # last BatchNorm of the 1->4->8->1 feature extractor: sets the scale of homo_score
    model.eval()
    return model
