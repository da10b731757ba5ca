"""CPU port of the reference's hot path on the SAME library calls the reference makes.

TEST INFRASTRUCTURE ONLY (see oracle/hdn_oracle.c header for the rule): used as
the second checker in tests/ and as the timed CPU baseline in bench.py
(`cpu_baseline.kind == "port"`, and `--impl reference`).  The reference is pure
Python on top of torch (grouped F.conv2d, F.pad, F.grid_sample, torch.inverse,
torch.gather); /root/reference cannot travel to the GPU box, so this file
restates each operator on those same torch entry points, which is what the
reference's CPU forward executes.  Pinned against goldens generated from the
real reference (tests/test_oracle_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# K1  -- hdn/core/xcorr.py:37-46
def xcorr_depthwise(x, k):
    B, C, Hk, Wk = k.shape
    Bx = x.shape[0]
    if B == 1 and Bx > 1:  # template shared by the batch
        k = k.expand(Bx, C, Hk, Wk)
        B = Bx
    y = F.conv2d(x.reshape(1, B * C, x.shape[2], x.shape[3]), k.reshape(B * C, 1, Hk, Wk), groups=B * C)
    return y.reshape(B, C, y.shape[2], y.shape[3])


# K2  -- hdn/core/xcorr.py:48-61 (rows wrap, then columns replicate with the unpadded W//2)
def xcorr_depthwise_circular(x, k):
    ph, pw = x.shape[2] // 2, x.shape[3] // 2
    xp = F.pad(x, (0, 0, ph, ph), mode="circular")
    xp = F.pad(xp, (pw, pw, 0, 0), mode="replicate")
    return xcorr_depthwise(xp, k)


# K3  -- hdn/models/logpolar.py:58-74 (grid), :103-118 (normalise), :124 (sample)
def logpolar_grid(S, delta_rot, polar, H, W):
    j = torch.linspace(0, S - 1, S)
    i = torch.linspace(0, S - 1, S)
    mag = math.log(S / 2) / S
    rho = torch.exp(mag * j) - 1.0
    theta = i * 2.0 * math.pi / S + delta_rot
    th, rh = torch.meshgrid([theta, rho], indexing="ij")
    gx = (rh * torch.cos(th)).unsqueeze(0) + polar[:, 0].reshape(-1, 1, 1)
    gy = (rh * torch.sin(th)).unsqueeze(0) + polar[:, 1].reshape(-1, 1, 1)
    return torch.stack((gx / (H // 2), gy / (W // 2)), dim=3)


def logpolar(img, polar, delta_rot, S):
    B, _, H, W = img.shape
    if polar is None:
        polar = torch.zeros(B, 2)
    grid = logpolar_grid(S, delta_rot, polar, H, W)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=False)


# K5  -- Oneline_DLTv1/utils.py:7-67 (8-vector case: one quad, points re-ordered [0,1,3,2])
def dlt_solve(src, off):
    B = src.shape[0]
    order = torch.tensor([0, 1, 3, 2])
    s = src.reshape(B, 4, 2)[:, order]
    d = s + off.reshape(B, 4, 2)[:, order]
    one = torch.ones(B, 4, 1)
    xy1 = torch.cat((s, one), 2)
    z = torch.zeros_like(xy1)
    M1 = torch.cat((torch.cat((xy1, z), 2), torch.cat((z, xy1), 2)), 2).reshape(B, 8, 6)
    M2 = torch.matmul(d.reshape(-1, 2, 1), s.reshape(-1, 1, 2)).reshape(B, 8, 2)
    A = torch.cat((M1, -M2), 2)
    h8 = torch.matmul(torch.inverse(A), d.reshape(B, 8, 1)).reshape(B, 8)
    return torch.cat((h8, torch.ones(B, 1)), 1).reshape(B, 1, 3, 3)


# K4  -- Oneline_DLTv1/utils.py:257-274 -> :70-254
def homo_warp(img, Hm, M=None, Minv=None):
    B, Ch, H, W = img.shape
    if M is None:
        M = torch.tensor([[W / 2.0, 0.0, W / 2.0], [0.0, H / 2.0, H / 2.0], [0.0, 0.0, 1.0]])
        Minv = torch.inverse(M)
    theta = torch.matmul(torch.matmul(Minv.expand(B, 3, 3), Hm.reshape(B, 3, 3)), M.expand(B, 3, 3))
    xt = torch.linspace(-1.0, 1.0, W).reshape(1, W).expand(H, W).reshape(1, -1)
    yt = torch.linspace(-1.0, 1.0, H).reshape(H, 1).expand(H, W).reshape(1, -1)
    grid = torch.cat((xt, yt, torch.ones_like(xt)), 0)
    T = torch.matmul(theta, grid.unsqueeze(0).expand(B, 3, H * W))
    ts = T[:, 2].reshape(-1)
    ts = ts + 1e-6 * (1.0 - torch.ge(ts.abs(), 1e-7).float())
    x = (T[:, 0].reshape(-1) / ts + 1.0) * W / 2.0
    y = (T[:, 1].reshape(-1) / ts + 1.0) * H / 2.0
    x0 = torch.floor(x).int()
    y0 = torch.floor(y).int()
    x1 = (x0 + 1).clamp(0, W - 1)
    y1 = (y0 + 1).clamp(0, H - 1)
    x0 = x0.clamp(0, W - 1)
    y0 = y0.clamp(0, H - 1)
    base = (torch.arange(B) * (H * W)).repeat_interleave(H * W)
    flat = img.permute(0, 2, 3, 1).reshape(-1, Ch)

    def take(yy, xx):
        idx = (base + yy.long() * W + xx.long()).unsqueeze(1).expand(-1, Ch)
        return torch.gather(flat, 0, idx)

    Ia, Ib, Ic, Id = take(y0, x0), take(y1, x0), take(y0, x1), take(y1, x1)
    x0f, x1f, y0f, y1f = x0.float(), x1.float(), y0.float(), y1.float()
    wa = ((x1f - x) * (y1f - y)).unsqueeze(1)
    wb = ((x1f - x) * (y - y0f)).unsqueeze(1)
    wc = ((x - x0f) * (y1f - y)).unsqueeze(1)
    wd = ((x - x0f) * (y - y0f)).unsqueeze(1)
    out = wa * Ia + wb * Ib + wc * Ic + wd * Id
    return out.reshape(B, H, W, Ch).permute(0, 3, 1, 2)


# K6  -- hdn_tracker.py:82-89, proj_e2e:172-174 (per batch item, NumPy after the softmax like the reference)
def score_argmax(cls, loc, window=None, win_influence=0.0):
    B, _, N, _ = cls.shape
    L = loc.shape[1]
    idx = np.empty(B, np.int64)
    ps = np.empty(B, np.float64)
    sc = np.empty(B, np.float32)
    g = np.empty((B, L), np.float32)
    for b in range(B):
        s = cls[b].reshape(2, -1).permute(1, 0).softmax(1)[:, 1].numpy()
        p = s * (1 - win_influence) + window * win_influence if window is not None else s
        i = int(np.argmax(p))
        idx[b], ps[b], sc[b] = i, p[i], s[i]
        g[b] = loc[b].reshape(L, -1)[:, i].numpy()
    return idx, ps, sc, g


def m1_chain(feats, threads=None):
    """One pass of the corr+warp+DLT chain on CPU tensors (same work list as hdn_b200.engine.M1Engine.run).

    feats: dict with xs/ks (6 pairs), xl/kl (6 lp pairs), img, polar, gray, src, off.
    """
    if threads:
        torch.set_num_threads(threads)
    out = {}
    out["corr"] = [xcorr_depthwise(x, k) for x, k in zip(feats["xs"], feats["ks"])]
    out["corr_lp"] = [xcorr_depthwise_circular(x, k) for x, k in zip(feats["xl"], feats["kl"])]
    out["x_lp"] = logpolar(feats["img"], feats.get("polar"), 0.0, feats["S"])
    Hm = dlt_solve(feats["src"], feats["off"]).squeeze(1)
    out["H"] = Hm
    out["warp"] = homo_warp(feats["gray"], Hm, feats["M"], feats["Minv"])
    return out


# ---- the BAN heads from neck features: what hdn_b200.head_engine.HeadEngine fuses ------------------------------
def _conv_bn_relu(x, w, scale, shift):
    """nn.Conv2d(bias=False) -> eval-mode BatchNorm2d (folded to scale / shift) -> ReLU   (ban.py:56-61)"""
    return F.relu(F.conv2d(x, w) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))


def ban_head(w, z_fs, x_fs, circular, kernels=None):
    """MultiBAN.forward / MultiCircBAN.forward (hdn/models/head/ban.py:73-78, 102-127; ban_lp.py:33-40, 65-92) on the torch calls
    the reference makes.  w: HeadWeights.raw (per branch cls2, loc2, cls3, loc3, cls4, loc4).  kernels: pre-computed template side
    (None = recompute `conv_kernel(z_f)` like the reference does on every frame, ban.py:74).  -> (cls, loc)"""
    corr = xcorr_depthwise_circular if circular else xcorr_depthwise
    outs = []
    for i in range(6):
        k = kernels[i] if kernels is not None else _conv_bn_relu(z_fs[i // 2], w["kernel_w"][i], w["kernel_scale"][i], w["kernel_shift"][i])
        s = _conv_bn_relu(x_fs[i // 2], w["search_w"][i], w["search_scale"][i], w["search_shift"][i])
        h = _conv_bn_relu(corr(s, k), w["hidden_w"][i], w["hidden_scale"][i], w["hidden_shift"][i])
        outs.append(F.conv2d(h, w["w2"][i].view(w["w2"][i].shape[0], -1, 1, 1), w["b2"][i]))
    cls = sum(outs[2 * l] * w["cls_w"][l] for l in range(3))
    loc = sum(outs[2 * l + 1] * w["loc_scale"][l] * w["loc_w"][l] for l in range(3))
    return cls, loc


def template_kernels(w, z_fs):
    return [_conv_bn_relu(z_fs[i // 2], w["kernel_w"][i], w["kernel_scale"][i], w["kernel_shift"][i]) for i in range(6)]


def fused_chain(feats, w_sim, w_lp, window, win_influence, kernels=None):
    """One pass of the chain HeadEngine runs, on CPU tensors: both BAN heads from neck features, K6 epilogues, K3, K5 + K4.
    kernels: (k_sim, k_lp) hoisted template kernels or None (reference behaviour: recomputed per call)."""
    out = {}
    out["cls"], out["loc"] = ban_head(w_sim, feats["zf"], feats["xf"], False, kernels[0] if kernels else None)
    out["cls_lp"], out["loc_lp"] = ban_head(w_lp, feats["zf_lp"], feats["xf_lp"], True, kernels[1] if kernels else None)
    out["idx"], out["pscore"], out["score"], out["center"] = score_argmax(out["cls"], out["loc"], window, win_influence)
    out["idx_lp"], out["pscore_lp"], out["score_lp"], out["sim_lp"] = score_argmax(out["cls_lp"], out["loc_lp"], None, 0.0)
    out["x_lp"] = logpolar(feats["img"], None, 0.0, feats["S"])
    Hm = dlt_solve(feats["src"], feats["off"]).squeeze(1)
    out["H"] = Hm
    out["warp"] = homo_warp(feats["gray"], Hm, feats["M"], feats["Minv"])
    return out
