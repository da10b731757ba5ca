"""Batched corr+warp+DLT chain ("M1" in SURVEY.md 8(d)): the per-pair hot path for B independent pairs.

Per pair:   6 x K1  template<->search depth-wise correlation (3 levels x {cls, loc})    ban.py:102-127
            6 x K2  circular correlation of the log-polar branch                        ban_lp.py:65-92
            K3      log-polar resampling of the search crop                             logpolar.py:120-134
            K5+K4   4-point DLT of the predicted corner offsets, then the projective warp
                                                                                        model_builder...py:195-210
            2 x K6  softmax/window/arg-max epilogues on the head outputs                proj_e2e.py:168-212
All on given fp32 features (the dense backbone/neck/head convolutions are not part of this chain).

The chain is 6 kernel launches per step (the 6 same-shape correlations of a branch share ONE launch),
replayed from a CUDA graph.  `run_host` is the end-to-end form: pinned host buffers in, pinned host
buffers out, H2D / compute / D2H pipelined over batch chunks on three streams.
"""
import math

import numpy as np
import torch

from . import _lib, ops

WORKLOADS = {
    # name: template/search crop sizes -> feature-map shapes ([probed] there)
    "256/512": dict(sim_x=61, sim_k=29, lp_x=29, lp_k=29, img=512, S=256, score=33, score_lp=29),
    "127/255": dict(sim_x=29, sim_k=5, lp_x=13, lp_k=13, img=255, S=127, score=25, score_lp=13),
    "win15": dict(sim_x=39, sim_k=15, lp_x=0, lp_k=0, img=0, S=0, score=0, score_lp=0),  # config 5: K1 only
}
C = 256
NPROB = 6
H4P = [0.0, 0.0, 0.0, 127.0, 127.0, 127.0, 127.0, 0.0]  # get_img_info.py:93-98
WIN_INFL = 0.1632532824922313  # experiments/tracker_homo_config/proj_e2e_GOT_unconstrained_v2.yaml:52


def xcorr_bytes(Hx, Wx, Hk, Wk, circular, shared_k_over=1):
    """Algorithmic bytes of one correlation per pair: 4*C*(Hx*Wx + h*w + Ho*Wo)."""
    Ho, Wo = ops.xcorr_out_hw(Hx, Wx, Hk, Wk, circular)
    return 4 * C * (Hx * Wx + Hk * Wk / shared_k_over + Ho * Wo)


def xcorr_flops(Hx, Wx, Hk, Wk, circular):
    Ho, Wo = ops.xcorr_out_hw(Hx, Wx, Hk, Wk, circular)
    return 2 * C * Ho * Wo * Hk * Wk


def algorithmic_bytes_per_pair(workload, B=1, shared_template=False):
    """Compulsory fp32 traffic per pair: every tensor of the chain touched once."""
    w = WORKLOADS[workload]
    share = B if shared_template else 1
    out = {"k1": NPROB * xcorr_bytes(w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"], False, share)}
    if w["lp_x"]:
        out["k2"] = NPROB * xcorr_bytes(w["lp_x"], w["lp_x"], w["lp_k"], w["lp_k"], True, share)
        out["k3"] = 12 * (w["img"] ** 2 + w["S"] ** 2)
        out["k4"] = 4 * 2 * 127 * 127 + 36
        out["k5"] = 100
    out["total"] = sum(out.values())
    return out


def flops_per_pair(workload):
    w = WORKLOADS[workload]
    f = NPROB * xcorr_flops(w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"], False)
    if w["lp_x"]:
        f += NPROB * xcorr_flops(w["lp_x"], w["lp_x"], w["lp_k"], w["lp_k"], True)
    return f


def make_inputs(workload, B, seed=1, device="cpu", shared_template=False, pin=False):
    """Synthetic inputs of SURVEY 8(d) config 2/3/4/5 shapes (features ~ N(0,1)*0.1, images U[0,255), offsets U(-8,8))."""
    w = WORKLOADS[workload]
    g = torch.Generator(device="cpu").manual_seed(seed)
    Bk = 1 if shared_template else B

    def feat(*shape):
        t = torch.randn(shape, generator=g) * 0.1
        return t

    d = {"xs": [feat(B, C, w["sim_x"], w["sim_x"]) for _ in range(NPROB)], "ks": [feat(Bk, C, w["sim_k"], w["sim_k"]) for _ in range(NPROB)]}
    if w["lp_x"]:
        d["xl"] = [feat(B, C, w["lp_x"], w["lp_x"]) for _ in range(NPROB)]
        d["kl"] = [feat(Bk, C, w["lp_k"], w["lp_k"]) for _ in range(NPROB)]
        d["img"] = torch.rand((B, 3, w["img"], w["img"]), generator=g) * 255.0
        d["gray"] = torch.randn((B, 1, 127, 127), generator=g)
        d["off"] = torch.rand((B, 8), generator=g) * 16.0 - 8.0
        d["src"] = torch.tensor(H4P).repeat(B, 1)
        d["cls"] = torch.randn((B, 2, w["score"], w["score"]), generator=g) * 2.0
        d["loc"] = torch.randn((B, 2, w["score"], w["score"]), generator=g)
        d["cls_lp"] = torch.randn((B, 2, w["score_lp"], w["score_lp"]), generator=g) * 2.0
        d["loc_lp"] = torch.randn((B, 4, w["score_lp"], w["score_lp"]), generator=g) * 0.5

    def place(t):
        if device != "cpu":
            return t.to(device)
        return t.pin_memory() if pin else t

    return {k: ([place(t) for t in v] if isinstance(v, list) else place(v)) for k, v in d.items()}


class M1Engine:
    """Device-resident chain for a fixed (workload, B).  All buffers are allocated once."""

    def __init__(self, workload="256/512", B=64, device="cuda", shared_template=False, use_graph=True):
        _lib.lib()  # fail loudly if the CUDA library is missing
        self.w = WORKLOADS[workload]
        self.workload, self.B, self.device, self.shared = workload, B, torch.device(device), shared_template
        self.full = bool(self.w["lp_x"])
        w, dev = self.w, self.device
        e = lambda *s, dt=torch.float32: torch.empty(s, device=dev, dtype=dt)  # noqa: E731
        so, _ = ops.xcorr_out_hw(w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"], False)
        self.out = {"corr": [e(B, C, so, so) for _ in range(NPROB)]}
        if self.full:
            lo, _ = ops.xcorr_out_hw(w["lp_x"], w["lp_x"], w["lp_k"], w["lp_k"], True)
            self.out.update(corr_lp=[e(B, C, lo, lo) for _ in range(NPROB)], x_lp=e(B, 3, w["S"], w["S"]), H=e(B, 3, 3), warp=e(B, 1, 127, 127),
                            idx=e(B, dt=torch.int64), pscore=e(B, dt=torch.float64), score=e(B), center=e(B, 2),
                            idx_lp=e(B, dt=torch.int64), pscore_lp=e(B, dt=torch.float64), score_lp=e(B), sim_lp=e(B, 4))
            n = w["score"]
            self.window = torch.from_numpy(np.outer(np.hanning(n), np.hanning(n)).flatten()).to(dev)
        self.inp = None
        self.graph = None
        self.use_graph = use_graph
        self.launches_per_step = 6 if self.full else 1
        self.h2d_streams = 1  # run_host: number of upload streams (2 = alternate the copies between two copy engines)

    # ---- device-resident path ---------------------------------------------------------------------
    def bind(self, inputs):
        """inputs: dict of DEVICE tensors shaped like make_inputs(...).  A template shared by the whole batch (ks / kl of batch size 1:
        config 3, the tracker) has its row spectra taken here, once per template, for the transform-domain kernels."""
        self.inp = dict(inputs)
        self.graph = None
        w = self.w
        if self.inp["ks"][0].shape[0] == 1 and self.B > 1:
            spec = ops.xcorr_template_spectra(self.inp["ks"], w["sim_x"], w["sim_x"], False)
            if spec is not None:
                self.inp["ks_spec"] = spec
            if self.full:
                spec = ops.xcorr_template_spectra(self.inp["kl"], w["lp_x"], w["lp_x"], True)
                if spec is not None:
                    self.inp["kl_spec"] = spec

    def _launch(self, inp, out, B):
        L = _lib.lib()
        st = ops._stream()
        w = self.w
        vp = ops._vp
        arr = vp * NPROB
        kB = inp["ks"][0].shape[0]
        kbs = 0 if (kB == 1 and B > 1) else C * w["sim_k"] ** 2
        if "ks_spec" in inp:
            _lib.check(L.hdn_xcorr_dw_multi_spec_f32(NPROB, arr(*[t.data_ptr() for t in inp["xs"]]), arr(*[t.data_ptr() for t in inp["ks_spec"]]),
                                                     arr(*[t.data_ptr() for t in out["corr"]]), B, C, w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"],
                                                     0, st), "K1 (cached template spectra)")
        else:
            _lib.check(L.hdn_xcorr_dw_multi_f32(NPROB, arr(*[t.data_ptr() for t in inp["xs"]]), arr(*[t.data_ptr() for t in inp["ks"]]),
                                                arr(*[t.data_ptr() for t in out["corr"]]), B, C, w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"], 0,
                                                kbs, st), "K1")
        if not self.full:
            return
        kbs = 0 if (kB == 1 and B > 1) else C * w["lp_k"] ** 2
        if "kl_spec" in inp:
            _lib.check(L.hdn_xcorr_dw_multi_spec_f32(NPROB, arr(*[t.data_ptr() for t in inp["xl"]]), arr(*[t.data_ptr() for t in inp["kl_spec"]]),
                                                     arr(*[t.data_ptr() for t in out["corr_lp"]]), B, C, w["lp_x"], w["lp_x"], w["lp_k"], w["lp_k"],
                                                     1, st), "K2 (cached template spectra)")
        else:
            _lib.check(L.hdn_xcorr_dw_multi_f32(NPROB, arr(*[t.data_ptr() for t in inp["xl"]]), arr(*[t.data_ptr() for t in inp["kl"]]),
                                                arr(*[t.data_ptr() for t in out["corr_lp"]]), B, C, w["lp_x"], w["lp_x"], w["lp_k"], w["lp_k"], 1,
                                                kbs, st), "K2")
        p = lambda t: vp(t.data_ptr())  # noqa: E731
        _lib.check(L.hdn_logpolar_f32(p(inp["img"]), None, 0.0, p(out["x_lp"]), B, 3, w["img"], w["img"], w["S"], st), "K3")
        _lib.check(L.hdn_dlt_warp_f32(p(inp["src"]), p(inp["off"]), p(inp["gray"]), None, None, p(out["H"]), p(out["warp"]), B, 1, 127, 127, st),
                   "K5+K4")
        _lib.check(L.hdn_score_argmax_f32(p(inp["cls"]), p(inp["loc"]), p(self.window), WIN_INFL, p(out["idx"]), p(out["pscore"]),
                                          p(out["score"]), p(out["center"]), B, 2, w["score"], st), "K6")
        _lib.check(L.hdn_score_argmax_f32(p(inp["cls_lp"]), p(inp["loc_lp"]), None, 0.0, p(out["idx_lp"]), p(out["pscore_lp"]),
                                          p(out["score_lp"]), p(out["sim_lp"]), B, 4, w["score_lp"], st), "K6lp")

    def run(self):
        """One step over the bound device inputs (graph replay after the first call)."""
        if not self.use_graph:
            self._launch(self.inp, self.out, self.B)
            return self.out
        if self.graph is None:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._launch(self.inp, self.out, self.B)  # warm (sets smem attributes) before capture
            torch.cuda.current_stream(self.device).wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._launch(self.inp, self.out, self.B)
        self.graph.replay()
        return self.out

    # ---- end-to-end path: pinned host in -> pinned host out ---------------------------------------
    def _views(self, d, lo, hi):
        r = {}
        for k, v in d.items():
            if isinstance(v, list):
                r[k] = [t if t.shape[0] == 1 and self.B > 1 and k in ("ks", "kl") else t[lo:hi] for t in v]
            else:
                r[k] = v[lo:hi]
        return r

    def alloc_host_io(self, host_inputs):
        """Device staging for host inputs + pinned host outputs; returns (h2d_bytes, d2h_bytes) per step."""
        self.dev_in = {k: ([torch.empty_like(t, device=self.device) for t in v] if isinstance(v, list) else torch.empty_like(v, device=self.device))
                       for k, v in host_inputs.items()}
        self.host_out = {k: ([torch.empty_like(t, device="cpu").pin_memory() for t in v] if isinstance(v, list)
                             else torch.empty_like(v, device="cpu").pin_memory()) for k, v in self.out.items()}
        self.s_h2d, self.s_cmp, self.s_d2h, self.s_h2d2 = (torch.cuda.Stream(device=self.device) for _ in range(4))
        nbytes = lambda d: sum(sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) else v.numel() * v.element_size()  # noqa: E731
                               for v in d.values())
        return nbytes(host_inputs), nbytes(self.host_out)

    def run_host(self, host_inputs, chunks=8):
        """Pinned host buffers in, pinned host buffers out.  The batch is cut into `chunks` slices; slice i+1 uploads
        while slice i computes and slice i-1 downloads (three streams, events between them)."""
        B = self.B
        cur = torch.cuda.current_stream(self.device)
        up = (self.s_h2d, self.s_h2d2) if self.h2d_streams == 2 else (self.s_h2d,)
        for s in up + (self.s_cmp, self.s_d2h):
            s.wait_stream(cur)
        step = math.ceil(B / chunks)
        shared_done = False
        for lo in range(0, B, step):
            hi = min(B, lo + step)
            hin, din = self._views(host_inputs, lo, hi), self._views(self.dev_in, lo, hi)
            # uploads alternate between the copy streams: the per-copy set-up of one overlaps the transfer of the other
            n_copy, ev_ups = 0, []
            for k, v in hin.items():
                pairs = zip(din[k], v) if isinstance(v, list) else [(din[k], v)]
                for dt, ht in pairs:
                    if self.shared and k in ("ks", "kl") and shared_done:
                        continue
                    with torch.cuda.stream(up[n_copy % len(up)]):
                        dt.copy_(ht, non_blocking=True)
                    n_copy += 1
            shared_done = True
            for s in up:
                ev = torch.cuda.Event()
                ev.record(s)
                ev_ups.append(ev)
            dout = self._views(self.out, lo, hi)
            with torch.cuda.stream(self.s_cmp):
                for ev in ev_ups:
                    self.s_cmp.wait_event(ev)
                self._launch(din, dout, hi - lo)
                ev_c = torch.cuda.Event()
                ev_c.record(self.s_cmp)
            hout = self._views(self.host_out, lo, hi)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev_c)
                for k, v in dout.items():
                    pairs = zip(hout[k], v) if isinstance(v, list) else [(hout[k], v)]
                    for ht, dt in pairs:
                        ht.copy_(dt, non_blocking=True)
        for s in up + (self.s_cmp, self.s_d2h):
            cur.wait_stream(s)
        return self.host_out
