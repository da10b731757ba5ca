#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (CPU, this container).

TEST INFRASTRUCTURE ONLY.  Run from anywhere:  python oracle/gen_golden.py
Writes tests/golden/ops_*.npz.  Each file stores the exact inputs and the
reference's outputs for one hot-path operator, so the oracle restatements
(oracle/hdn_oracle.c, oracle/torch_port.py) and the CUDA kernels can be pinned
without /root/reference being present (it is absent on the GPU box).

Reference call sites exercised (all unmodified, imported from /root/reference):
  K1  hdn/core/xcorr.py:37-46            xcorr_depthwise
  K2  hdn/core/xcorr.py:48-61            xcorr_depthwise_circular
  K3  hdn/models/logpolar.py:50-134      STN_Polar.forward
  K4  Oneline_DLTv1/utils.py:257-274     transform (-> transformer :70-254)
  K5  Oneline_DLTv1/utils.py:7-67        DLT_solve
  K6  hdn/tracker/hdn_tracker.py:82-89   _convert_score
      hdn/tracker/base_tracker.py:54-59  _convert_c
      hdn/tracker/hdn_tracker.py:51-67   _convert_logpolar_simi
      hdn/tracker/hdn_tracker_proj_e2e.py:168-185,199-212 window/argmax/gates
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

ref_import.install()
import torch  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)  # deterministic accumulation order in oneDNN


def rng(seed):
    return np.random.default_rng(seed)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %8.1f KB  %s" % (name, os.path.getsize(path) / 1024.0, {k: v.shape for k, v in arrays.items()}))


# ----------------------------------------------------------------------------- K1 / K2
def gen_xcorr():
    from hdn.core.xcorr import xcorr_depthwise, xcorr_depthwise_circular

    cases = {  # name: (B, C, Hx, Wx, Hk, Wk, circular)
        "k1_native_29x5": (2, 8, 29, 29, 5, 5, 0),       # 127/255 crops: 29^2 (*) 5^2 -> 25^2
        "k1_256_61x29": (1, 4, 61, 61, 29, 29, 0),       # 256/512 crops
        "k1_win15_39x15": (1, 4, 39, 39, 15, 15, 0),     # config 5: 15x15 window
        "k1_ragged_11x7_k3x4": (2, 3, 11, 7, 3, 4, 0),   # non-square, C not multiple of 4
        "k1_full_7x7": (1, 5, 7, 7, 7, 7, 0),            # kernel == input -> 1x1 output
        "k2_native_13": (2, 8, 13, 13, 13, 13, 1),       # lp branch: 13^2 circ (*) 13^2 -> 13^2
        "k2_256_29": (1, 4, 29, 29, 29, 29, 1),          # lp branch at INSTANCE_SIZE=512
        "k2_ragged_9x6_k5x3": (2, 3, 9, 6, 5, 3, 1),     # kernel smaller than input, non-square
    }
    for i, (name, (B, C, Hx, Wx, Hk, Wk, circ)) in enumerate(cases.items()):
        g = rng(100 + i)
        x = g.standard_normal((B, C, Hx, Wx)).astype(np.float32)
        k = (g.standard_normal((B, C, Hk, Wk)) * 0.1).astype(np.float32)
        fn = xcorr_depthwise_circular if circ else xcorr_depthwise
        out = fn(t(x), t(k)).numpy()
        save("ops_" + name, x=x, k=k, out=out, circular=np.int32(circ))


# ----------------------------------------------------------------------------- K3
def gen_logpolar():
    from hdn.core.config import cfg
    from hdn.models.logpolar import STN_Polar

    cases = {  # name: (INSTANCE_SIZE, B, Ch, H, W, polar, delta)
        "k3_native_255": (255, 2, 3, 255, 255, [[0.0, 0.0], [3.25, -7.5]], [0, 0]),
        "k3_rot_255": (255, 1, 1, 255, 255, [[0.0, 0.0]], [0, 0.3]),
        "k3_512": (512, 1, 1, 512, 512, [[0.0, 0.0]], [0, 0]),
        "k3_border_64": (64, 1, 2, 40, 40, [[15.0, -12.0]], [0, -1.1]),  # samples leave the image -> border clamp
    }
    for i, (name, (inst, B, Ch, H, W, polar, delta)) in enumerate(cases.items()):
        g = rng(300 + i)
        img = (g.random((B, Ch, H, W)) * 255.0).astype(np.float32)
        pol = np.asarray(polar, np.float32)
        stn = STN_Polar(inst)
        out, grid = stn(t(img), t(pol), delta)
        arrays = dict(polar=pol, delta=np.asarray(delta, np.float64), inst=np.int32(inst), out=out.numpy(), seed=np.int64(300 + i),
                      shape=np.asarray([B, Ch, H, W], np.int64))
        if img.size <= 70000:
            arrays["img"] = img
        save("ops_" + name, **arrays)


# ----------------------------------------------------------------------------- K5 / K4
H4P = np.asarray([0, 0, 0, 127, 127, 127, 127, 0], np.float32)  # merge_tmp_search, get_img_info.py:93-98


def gen_dlt_and_warp():
    from homo_estimator.Deep_homography.Oneline_DLTv1.utils import DLT_solve, transform

    g = rng(500)
    B = 16
    src = np.tile(H4P, (B, 1))
    off = g.uniform(-8, 8, (B, 8)).astype(np.float32)
    off[0] = 0.0                       # identity
    off[1] = g.uniform(-40, 40, 8)     # large displacement
    src2 = src.copy()
    src2[2] = [10, 20, 12, 90, 100, 110, 95, 15]  # a non-canonical source quad
    H = DLT_solve(t(src2), t(off)).numpy()  # [B,1,3,3]
    save("ops_k5_dlt", src=src2, off=off, H=H)

    # K4: the exact call made by ModelBuilder.track_proj (model_builder...py:196-210)
    Bw = 4
    img = g.standard_normal((Bw, 1, 127, 127)).astype(np.float32)
    Hm = H[:Bw, 0].copy()
    Hm[2] = DLT_solve(t(src[:1]), t((g.uniform(-30, 30, (1, 8))).astype(np.float32))).numpy()[0, 0]
    M = np.asarray([[63.5, 0, 63.5], [0, 63.5, 63.5], [0, 0, 1]], np.float32)
    M_t = t(M)
    M_inv = torch.inverse(M_t)
    M_tile = M_t.unsqueeze(0).expand(Bw, 3, 3)
    M_tile_inv = M_inv.unsqueeze(0).expand(Bw, 3, 3)
    patch_idx = t(np.tile(np.arange(127 * 127, dtype=np.float32), (Bw, 1)))
    y_t = torch.arange(0, Bw * 127 * 127, 127 * 127)
    batch_idx = y_t.unsqueeze(1).expand(Bw, 127 * 127).reshape(-1)
    out = transform(127, 127, M_tile_inv, t(Hm), M_tile, t(img), patch_idx, batch_idx).numpy()
    save("ops_k4_warp", img=img, H=Hm, out=out)

    # smaller multi-channel / non-square variant through the same function (M built for that size)
    Bw, Ch, Hh, Ww = 2, 2, 24, 40
    img = g.standard_normal((Bw, Ch, Hh, Ww)).astype(np.float32)
    Hs = np.stack([np.eye(3, dtype=np.float32),
                   np.asarray([[1.05, 0.08, -1.5], [-0.06, 0.97, 2.25], [2e-4, -3e-4, 1]], np.float32)])
    M = np.asarray([[Ww / 2.0, 0, Ww / 2.0], [0, Hh / 2.0, Hh / 2.0], [0, 0, 1]], np.float32)
    M_t = t(M)
    M_inv = torch.inverse(M_t)
    patch_idx = t(np.tile(np.arange(Hh * Ww, dtype=np.float32), (Bw, 1)))
    y_t = torch.arange(0, Bw * Hh * Ww, Hh * Ww)
    batch_idx = y_t.unsqueeze(1).expand(Bw, Hh * Ww).reshape(-1)
    out = transform(Hh, Ww, M_inv.unsqueeze(0).expand(Bw, 3, 3), t(Hs), M_t.unsqueeze(0).expand(Bw, 3, 3), t(img), patch_idx,
                    batch_idx).numpy()
    save("ops_k4_warp_small", img=img, H=Hs, M=M, M_inv=M_inv.numpy(), out=out)


# ----------------------------------------------------------------------------- K6
def gen_score():
    from hdn.core.config import cfg
    from hdn.tracker.hdn_tracker import hdnTracker
    from hdn.tracker.base_tracker import SiameseTracker

    class _Self:  # the three methods only read these attributes
        cls_out_channels = 2

    g = rng(600)
    win_infl = 0.1632532824922313  # experiments/tracker_homo_config/proj_e2e_GOT_unconstrained_v2.yaml:52
    N = 25
    hanning = np.hanning(N)
    window = np.outer(hanning, hanning).flatten()
    points = hdnTracker.generate_points(_Self(), 8, N)
    Bn = 6
    cls = (g.standard_normal((Bn, 2, N, N)) * 2.0).astype(np.float32)
    loc = g.standard_normal((Bn, 2, N, N)).astype(np.float32)
    cls[1, :, :, :] = 0.0            # all-tie: argmax must return first index (np.argmax rule)
    cls[2, 0] = 8.0                  # everything background -> pscore below the 0.05 gate away from centre
    cls[2, 1] = -8.0
    idxs, pbest, sbest, centers = [], [], [], []
    for b in range(Bn):
        score = hdnTracker._convert_score(_Self(), t(cls[b:b + 1].copy()))
        pred_c = SiameseTracker._convert_c(_Self(), t(loc[b:b + 1].copy()), points)
        pscore = score * (1 - win_infl) + window * win_infl          # proj_e2e:172-173
        best = int(np.argmax(pscore))                                 # :174
        idxs.append(best)
        pbest.append(float(pscore[best]))
        sbest.append(float(score[best]))
        centers.append(pred_c[:, best].copy())
    save("ops_k6_score", cls=cls, loc=loc, window=window, win_infl=np.float64(win_infl), idx=np.asarray(idxs, np.int64),
         pscore=np.asarray(pbest, np.float64), score=np.asarray(sbest, np.float32), center=np.asarray(centers, np.float32),
         points=points)

    # lp branch: no window, 4-channel loc, scale/rot decode (proj_e2e:199-212, hdn_tracker.py:51-67)
    N = 13
    points_lp = hdnTracker.generate_points_lp(_Self(), cfg.POINT.STRIDE_LP, cfg.POINT.STRIDE_LP, N)
    cls = (g.standard_normal((Bn, 2, N, N)) * 2.0).astype(np.float32)
    loc = (g.standard_normal((Bn, 4, N, N)) * 0.5).astype(np.float32)
    idxs, sbest, sims = [], [], []
    for b in range(Bn):
        score = hdnTracker._convert_score(_Self(), t(cls[b:b + 1].copy()))
        best = int(np.argmax(score))
        pred = hdnTracker._convert_logpolar_simi(_Self(), t(loc[b:b + 1].copy()), points_lp, best, 0)
        idxs.append(best)
        sbest.append(float(score[best]))
        sims.append(pred[:, best].copy())
    save("ops_k6_score_lp", cls=cls, loc=loc, idx=np.asarray(idxs, np.int64), score=np.asarray(sbest, np.float32),
         sim=np.asarray(sims, np.float32), points=points_lp, stride_lp=np.int32(cfg.POINT.STRIDE_LP),
         exemplar=np.int32(cfg.TRAIN.EXEMPLAR_SIZE))


if __name__ == "__main__":
    which = sys.argv[1:] or ["xcorr", "logpolar", "dlt", "score"]
    if "xcorr" in which:
        gen_xcorr()
    if "logpolar" in which:
        gen_logpolar()
    if "dlt" in which:
        gen_dlt_and_warp()
    if "score" in which:
        gen_score()
