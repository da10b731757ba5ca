"""Fused BAN heads + warp / DLT chain from NECK features: the plug-in boundary one stage further up.

`M1Engine` (engine.py) takes the *outputs of the head's 3x3 convolutions* from the caller -- tensors that are born on the
device in the real pipeline, 54 MB per 256/512 pair, which made its end-to-end form PCIe-bound at 1.2 k frames/s.  This
engine takes what `ModelBuilder.track_new` / `track_new_lp` have after the neck (model_builder_e2e_unconstrained_v2.py:
134-137, 147-151): 18.3 MB per pair in, ~40 KB out.  Per batch of pairs:

    MultiBAN.forward      (hdn/models/head/ban.py:102-127)        MultiCircBAN.forward (ban_lp.py:65-92)
      conv_search 3x3+BN+ReLU x6   1 launch  tcgen05 3xTF32 implicit GEMM (hdn_conv_gemm_multi_f32)
      xcorr_depthwise x6           1 launch  transform-domain / direct correlation (hdn_xcorr_dw_multi_f32)      K1 / K2
      head 1x1+BN+ReLU, 1x1 x6     1 launch  tcgen05, hidden map stays on chip (hdn_head_project_multi_f32)
      level-weighted sum + K6      1 launch  (hdn_head_score_f32): combined maps + arg-max / gathered offsets          K6
    STN_Polar                      1 launch  (hdn_logpolar_f32)                                                        K3
    DLT_solve + transform          1 launch  (hdn_dlt_warp_f32)                                                     K5+K4

The template side (`conv_kernel(z_f)`, which the reference recomputes every frame, ban.py:74) is done once in
`set_template`.  Batches are processed in CHUNKS of a few pairs so that the intermediates of a chunk (6 x [256,61,61] search
features, 6 x [256,33,33] correlation maps per pair) are produced and consumed while still in the 126 MB L2 and the same
buffers are overwritten by the next chunk: they never need to reach HBM.  `run_host` pipelines H2D / compute / D2H over the chunks.
"""
import math

import numpy as np
import torch

from . import _lib, ops
from .engine import C, H4P, NPROB, WIN_INFL

# neck-feature sizes per workload [probed, SURVEY 8(a) a4-a7]: search / template maps of the similarity and log-polar branches
NECK = {
    "256/512": dict(xf=63, zf=31, xf_lp=31, zf_lp=31, img=512, S=256),
    "127/255": dict(xf=31, zf=7, xf_lp=15, zf_lp=15, img=255, S=127),
}
LEVELS = 3


class HeadWeights:
    """Weights of one Multi(Circ)BAN in the layout the kernels take: per branch [cls2, loc2, cls3, loc3, cls4, loc4] the packed
    conv_search / conv_kernel / head[0] weights with their folded BatchNorms, head[3] weight + bias, and the level weights."""

    FIELDS = ("search", "kernel", "hidden")

    def __init__(self, loc_channels):
        self.L = loc_channels
        self.raw = {}

    @classmethod
    def synthetic(cls, seed, loc_channels, gain, device="cpu"):
        """Random-init weights of the architecture (no checkpoint ships with the reference): kaiming convolutions, near-identity
        BatchNorms; the last 1x1 is scaled by `gain` so the logits keep a useful dynamic range (no saturated soft-max)."""
        g = torch.Generator().manual_seed(seed)
        rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
        self = cls(loc_channels)
        r = self.raw
        for name, k in (("search", 3), ("kernel", 3), ("hidden", 1)):
            r[name + "_w"] = [rn(C, C, k, k) * math.sqrt(2.0 / (C * k * k)) for _ in range(NPROB)]
            r[name + "_scale"] = [1.0 + 0.1 * rn(C) for _ in range(NPROB)]
            r[name + "_shift"] = [0.05 * rn(C) for _ in range(NPROB)]
        r["w2"] = [rn(2 if i % 2 == 0 else loc_channels, C) * (gain / math.sqrt(C)) for i in range(NPROB)]
        r["b2"] = [rn(2 if i % 2 == 0 else loc_channels) * 0.1 for i in range(NPROB)]
        r["cls_w"] = torch.softmax(1.0 + 0.2 * rn(LEVELS), 0).tolist()
        r["loc_w"] = torch.softmax(1.0 + 0.2 * rn(LEVELS), 0).tolist()
        r["loc_scale"] = (1.0 + 0.2 * rn(LEVELS)).tolist()
        return self.to(device)

    @classmethod
    def from_module(cls, head):
        """From a (mirrored or reference) MultiBAN / MultiCircBAN module in eval mode."""
        import torch.nn.functional as F
        branches = []
        for i in range(LEVELS):
            box = getattr(head, "box%d" % (i + 2))
            branches += [box.cls, box.loc]
        self = cls(branches[1].head[3].out_channels)
        r = self.raw

        def fold(bn):
            scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float()
            return scale, (bn.bias - bn.running_mean * scale).detach().float()

        for name, get in (("search", lambda b: b.conv_search), ("kernel", lambda b: b.conv_kernel), ("hidden", lambda b: b.head)):
            r[name + "_w"] = [get(b)[0].weight.detach().float() for b in branches]
            folded = [fold(get(b)[1]) for b in branches]
            r[name + "_scale"] = [f[0] for f in folded]
            r[name + "_shift"] = [f[1] for f in folded]
        r["w2"] = [b.head[3].weight.detach().float().reshape(b.head[3].out_channels, -1) for b in branches]
        r["b2"] = [b.head[3].bias.detach().float() for b in branches]
        if getattr(head, "weighted", False):
            r["cls_w"] = F.softmax(head.cls_weight.detach(), 0).tolist()
            r["loc_w"] = F.softmax(head.loc_weight.detach(), 0).tolist()
        else:
            r["cls_w"] = r["loc_w"] = [1.0 / LEVELS] * LEVELS
        r["loc_scale"] = head.loc_scale.detach().tolist()
        return self

    def to(self, device):
        for k, v in self.raw.items():
            if isinstance(v, list) and v and isinstance(v[0], torch.Tensor):
                self.raw[k] = [t.to(device).contiguous() for t in v]
        return self

    def pack(self):
        """Tensor-core records of the three convolutions (device weights required)."""
        self.packed = {name: [ops.pack_conv_weight(w) for w in self.raw[name + "_w"]] for name in self.FIELDS}
        return self


def make_inputs(workload, B, seed=1, shared_template=False, pin=False, u8_crop=False):
    """Synthetic NECK features (BatchNorm outputs: zero-mean, unit scale), crops U[0,255), offsets U(-8,8) -- SURVEY 8(d) config 2/3."""
    n = NECK[workload]
    g = torch.Generator(device="cpu").manual_seed(seed)
    Bk = 1 if shared_template else B
    rn = lambda *s: torch.randn(s, generator=g)  # noqa: E731
    d = {
        "xf": [rn(B, C, n["xf"], n["xf"]) for _ in range(LEVELS)], "xf_lp": [rn(B, C, n["xf_lp"], n["xf_lp"]) for _ in range(LEVELS)],
        "zf": [rn(Bk, C, n["zf"], n["zf"]) for _ in range(LEVELS)], "zf_lp": [rn(Bk, C, n["zf_lp"], n["zf_lp"]) for _ in range(LEVELS)],
        # the search crop: what get_subwindow hands over -- integer pixel values, as fp32 (the reference's upload) or as the uint8 they are
        "img": torch.randint(0, 256, (B, 3, n["img"], n["img"]), generator=g, dtype=torch.uint8) if u8_crop
        else torch.randint(0, 256, (B, 3, n["img"], n["img"]), generator=g, dtype=torch.uint8).float(),
        "gray": torch.randn((B, 1, 127, 127), generator=g),
        "off": torch.rand((B, 8), generator=g) * 16.0 - 8.0,
        "src": torch.tensor(H4P).repeat(B, 1),
    }
    place = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    return {k: ([place(t) for t in v] if isinstance(v, list) else place(v)) for k, v in d.items()}


FRAME_KEYS = ("xf", "xf_lp", "img", "gray", "off", "src")  # what crosses the boundary every frame (the template does not)
GAINS = {"256/512": (0.0123, 0.0071), "127/255": (0.36, 0.038)}  # last-layer gains of the synthetic heads: logit std ~ 2 (probed on the CPU port)


class HeadEngine:
    def __init__(self, workload="256/512", B=64, device="cuda", chunk=4, weights=None, seed=11):
        _lib.lib()  # fail loudly if the CUDA library is missing
        self.workload, self.B, self.device, self.chunk = workload, B, torch.device(device), max(1, min(chunk, B))
        self.n = NECK[workload]
        if weights is None:
            gs, gl = GAINS[workload]
            weights = (HeadWeights.synthetic(seed, 2, gs), HeadWeights.synthetic(seed + 1, 4, gl))
        self.w_sim, self.w_lp = (w.to(self.device).pack() for w in weights)
        n, dev, c = self.n, self.device, self.chunk
        e = lambda *s, dt=torch.float32: torch.empty(s, device=dev, dtype=dt)  # noqa: E731
        self.k_sim_hw = n["zf"] - 2
        self.k_lp_hw = n["zf_lp"] - 2
        self.s_sim_hw, self.s_lp_hw = n["xf"] - 2, n["xf_lp"] - 2
        self.N = ops.xcorr_out_hw(self.s_sim_hw, self.s_sim_hw, self.k_sim_hw, self.k_sim_hw, False)[0]
        self.N_lp = ops.xcorr_out_hw(self.s_lp_hw, self.s_lp_hw, self.k_lp_hw, self.k_lp_hw, True)[0]
        # chunk-sized intermediates, overwritten chunk after chunk (L2-resident between producer and consumer)
        self.S = [e(c, C, self.s_sim_hw, self.s_sim_hw) for _ in range(NPROB)]
        self.F = [e(c, C, self.N, self.N) for _ in range(NPROB)]
        self.P = [e(C // 128, c, 2, self.N * self.N) for _ in range(NPROB)]
        self.S_lp = [e(c, C, self.s_lp_hw, self.s_lp_hw) for _ in range(NPROB)]
        self.F_lp = [e(c, C, self.N_lp, self.N_lp) for _ in range(NPROB)]
        self.P_lp = [e(C // 128, c, 2 if i % 2 == 0 else 4, self.N_lp * self.N_lp) for i in range(NPROB)]
        # per-batch results
        self.out = {"cls": e(B, 2, self.N, self.N), "loc": e(B, 2, self.N, self.N), "cls_lp": e(B, 2, self.N_lp, self.N_lp),
                    "loc_lp": e(B, 4, self.N_lp, self.N_lp), "idx": e(B, dt=torch.int64), "pscore": e(B, dt=torch.float64), "score": e(B),
                    "center": e(B, 2), "idx_lp": e(B, dt=torch.int64), "pscore_lp": e(B, dt=torch.float64), "score_lp": e(B), "sim_lp": e(B, 4),
                    "H": e(B, 3, 3), "x_lp": e(B, 3, n["S"], n["S"]), "warp": e(B, 1, 127, 127)}
        self.window = torch.from_numpy(np.outer(np.hanning(self.N), np.hanning(self.N)).flatten()).to(dev)
        self.k_sim = self.k_lp = self.k_sim_spec = self.k_lp_spec = None
        self.inp = None
        self.launches_per_chunk = 11

    # ---- template side: once per template (the reference redoes it every frame, ban.py:74) --------------------------
    def set_template(self, zf, zf_lp):
        """zf / zf_lp: 3 neck maps each, [Bk,256,h,h] on the device; Bk = B (one template per pair) or 1 (shared by the batch)."""
        def kernels(w, feats):
            xs = [feats[i // 2] for i in range(NPROB)]
            return ops.conv_gemm_multi(xs, w.packed["kernel"], w.raw["kernel_scale"], w.raw["kernel_shift"], ksize=3, relu=True, valid=True)
        self.k_sim = kernels(self.w_sim, zf)
        self.k_lp = kernels(self.w_lp, zf_lp)
        # a template shared by the batch: its row spectra for the transform-domain correlation, once (None: no such kernel for the shape)
        shared = self.k_sim[0].shape[0] == 1
        self.k_sim_spec = ops.xcorr_template_spectra(self.k_sim, self.s_sim_hw, self.s_sim_hw, False) if shared else None
        self.k_lp_spec = ops.xcorr_template_spectra(self.k_lp, self.s_lp_hw, self.s_lp_hw, True) if shared else None
        return self

    def bind(self, inputs):
        """inputs: dict of DEVICE tensors shaped like make_inputs(...) (frame keys only are read)."""
        self.inp = inputs

    # ---- one chunk of pairs [lo, hi) ---------------------------------------------------------------------------------
    def _launch_chunk(self, inp, lo, hi, src_lo=None):
        """inp tensors are indexed [src_lo : src_lo + hi - lo] (src_lo = lo unless they are chunk-local staging buffers)."""
        L = _lib.lib()
        st = ops._stream()
        vp, pa = ops._vp, ops._ptr_array
        b = hi - lo
        s0 = lo if src_lo is None else src_lo
        o = self.out
        p = lambda t: vp(t.data_ptr())  # noqa: E731
        fl3 = ctypes_floats

        def head(w, xf, S, F, P, ks, circular, s_hw, k_hw, N, Lloc, maps, scores, window, winf, spec=None):
            xs = [xf[i // 2][s0:s0 + b] for i in range(NPROB)]
            _lib.check(L.hdn_conv_gemm_multi_f32(NPROB, pa(xs), pa(w.packed["search"]), pa(w.raw["search_scale"]), pa(w.raw["search_shift"]), pa(S), b,
                                                 C, C, s_hw + 2, s_hw + 2, 3, 1, 1, 1, st), "conv_search")
            kB = ks[0].shape[0]
            kk = [k if kB == 1 else k[lo:hi] for k in ks]
            kbs = 0 if (kB == 1 and b > 1) else C * k_hw * k_hw
            if spec is not None:
                _lib.check(L.hdn_xcorr_dw_multi_spec_f32(NPROB, pa(S), pa(spec), pa(F), b, C, s_hw, s_hw, k_hw, k_hw, int(circular), st), "xcorr (spectra)")
            else:
                _lib.check(L.hdn_xcorr_dw_multi_f32(NPROB, pa(S), pa(kk), pa(F), b, C, s_hw, s_hw, k_hw, k_hw, int(circular), kbs, st), "xcorr")
            if Lloc == 2:
                _lib.check(L.hdn_head_project_multi_f32(NPROB, pa(F), pa(w.packed["hidden"]), pa(w.raw["hidden_scale"]), pa(w.raw["hidden_shift"]),
                                                        pa(w.raw["w2"]), pa(P), b, C, N, N, 2, st), "head project")
            else:
                for sl, Lc in ((slice(0, None, 2), 2), (slice(1, None, 2), Lloc)):
                    _lib.check(L.hdn_head_project_multi_f32(LEVELS, pa(F[sl]), pa(w.packed["hidden"][sl]), pa(w.raw["hidden_scale"][sl]),
                                                            pa(w.raw["hidden_shift"][sl]), pa(w.raw["w2"][sl]), pa(P[sl]), b, C, N, N, Lc, st),
                               "head project")
            cls_map, loc_map = maps
            idx, ps, sc, g = scores
            # (the chunk-sized buffers S / F / P are used densely from their start with batch b: a short last chunk simply uses less)
            _lib.check(L.hdn_head_score_f32(LEVELS, C // 128, pa(P[0::2]), pa(P[1::2]), pa(w.raw["b2"][0::2]), pa(w.raw["b2"][1::2]),
                                            fl3(w.raw["cls_w"]), fl3(w.raw["loc_scale"]), fl3(w.raw["loc_w"]), p(cls_map[lo:hi]), p(loc_map[lo:hi]),
                                            p(window) if window is not None else None, winf, p(idx[lo:hi]), p(ps[lo:hi]), p(sc[lo:hi]), p(g[lo:hi]),
                                            b, Lloc, N, st), "head score")

        head(self.w_sim, inp["xf"], self.S, self.F, self.P, self.k_sim, False, self.s_sim_hw, self.k_sim_hw, self.N, 2, (o["cls"], o["loc"]),
             (o["idx"], o["pscore"], o["score"], o["center"]), self.window, WIN_INFL, self.k_sim_spec)
        head(self.w_lp, inp["xf_lp"], self.S_lp, self.F_lp, self.P_lp, self.k_lp, True, self.s_lp_hw, self.k_lp_hw, self.N_lp, 4,
             (o["cls_lp"], o["loc_lp"]), (o["idx_lp"], o["pscore_lp"], o["score_lp"], o["sim_lp"]), None, 0.0, self.k_lp_spec)
        n = self.n
        k3 = L.hdn_logpolar_u8 if inp["img"].dtype == torch.uint8 else L.hdn_logpolar_f32
        _lib.check(k3(p(inp["img"][s0:s0 + b]), None, 0.0, p(o["x_lp"][lo:hi]), b, 3, n["img"], n["img"], n["S"], st), "K3")
        _lib.check(L.hdn_dlt_warp_f32(p(inp["src"][s0:s0 + b]), p(inp["off"][s0:s0 + b]), p(inp["gray"][s0:s0 + b]), None, None, p(o["H"][lo:hi]),
                                      p(o["warp"][lo:hi]), b, 1, 127, 127, st), "K5+K4")

    def run(self):
        """One step over the bound device inputs, chunk by chunk."""
        if self.k_sim is None:
            raise RuntimeError("HeadEngine.set_template() first")
        for lo in range(0, self.B, self.chunk):
            self._launch_chunk(self.inp, lo, min(self.B, lo + self.chunk))
        return self.out

    # ---- end to end: pinned host in -> pinned host out -------------------------------------------------------------------
    HOST_OUT = ("cls", "loc", "cls_lp", "loc_lp", "idx", "pscore", "score", "center", "idx_lp", "pscore_lp", "score_lp", "sim_lp", "H")

    def alloc_host_io(self, host_inputs, slots=3):
        """Chunk-sized device staging ring for the per-frame inputs + pinned host buffers for the results.
        -> (h2d_bytes, d2h_bytes) per step.  x_lp and the warped patch stay on the device: their consumers (the log-polar
        backbone pass, ShareFeature) run there in the real pipeline."""
        c = self.chunk
        self.slots = slots
        self.stage = [{k: ([torch.empty((c,) + tuple(t.shape[1:]), device=self.device, dtype=t.dtype) for t in host_inputs[k]]
                           if isinstance(host_inputs[k], list)
                           else torch.empty((c,) + tuple(host_inputs[k].shape[1:]), device=self.device, dtype=host_inputs[k].dtype)) for k in FRAME_KEYS}
                      for _ in range(slots)]
        # two result sets: the host reads step i's results while step i + 1 is in flight
        self.host_out_sets = [{k: torch.empty_like(self.out[k], device="cpu").pin_memory() for k in self.HOST_OUT} for _ in range(2)]
        self.host_out = self.host_out_sets[0]
        self.s_h2d, self.s_cmp, self.s_d2h = (torch.cuda.Stream(device=self.device) for _ in range(3))
        self.ev_free = [None] * slots
        self.ev_read = None   # results of the previous step have left the device (its result buffers may be overwritten)
        self.steps_in_flight = 0
        nb = lambda t: t.numel() * t.element_size()  # noqa: E731
        h2d = sum(sum(nb(t) for t in host_inputs[k]) if isinstance(host_inputs[k], list) else nb(host_inputs[k]) for k in FRAME_KEYS)
        return h2d, sum(nb(t) for t in self.host_out.values())

    def run_host(self, host_inputs, wait=True):
        """Per-frame inputs in pinned host memory -> results in pinned host memory.  Chunk i+1 uploads while chunk i computes;
        the (small) results of the whole batch are downloaded once at the end.
        wait=False: return as soon as the step is queued (the caller's stream is not joined), so the uploads of the next step
        overlap this step's last chunks -- a stream of batches keeps the link busy; call finish() before reading the LAST
        returned result set.  Result sets alternate between two pinned buffers."""
        B, c = self.B, self.chunk
        cur = torch.cuda.current_stream(self.device)
        if self.steps_in_flight == 0:
            for s in (self.s_h2d, self.s_cmp, self.s_d2h):
                s.wait_stream(cur)
        host_out = self.host_out_sets[self.steps_in_flight % 2] if not wait else self.host_out_sets[0]
        for i, lo in enumerate(range(0, B, c)):
            hi = min(B, lo + c)
            slot = i % self.slots
            stage = self.stage[slot]
            with torch.cuda.stream(self.s_h2d):
                if self.ev_free[slot] is not None:
                    self.s_h2d.wait_event(self.ev_free[slot])  # the chunk that used this slot has been consumed
                for k in FRAME_KEYS:
                    v = host_inputs[k]
                    pairs = zip(stage[k], v) if isinstance(v, list) else [(stage[k], v)]
                    for dt, ht in pairs:
                        dt[:hi - lo].copy_(ht[lo:hi], non_blocking=True)
                ev_up = torch.cuda.Event()
                ev_up.record(self.s_h2d)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_up)
                if i == 0 and self.ev_read is not None:
                    self.s_cmp.wait_event(self.ev_read)  # the previous step's results have been read out of self.out
                self._launch_chunk(stage, lo, hi, src_lo=0)
                ev = torch.cuda.Event()
                ev.record(self.s_cmp)
                self.ev_free[slot] = ev
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_stream(self.s_cmp)
            for k in self.HOST_OUT:
                host_out[k].copy_(self.out[k], non_blocking=True)
            self.ev_read = torch.cuda.Event()
            self.ev_read.record(self.s_d2h)
        self.host_out = host_out
        if wait:
            self.finish()
        else:
            self.steps_in_flight += 1
        return host_out

    def finish(self):
        """Join the pipeline streams into the caller's stream (results are in host memory once that stream is synchronised)."""
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            cur.wait_stream(s)
        self.steps_in_flight = 0


def ctypes_floats(values):
    import ctypes
    return (ctypes.c_float * len(values))(*[float(v) for v in values])


def algorithmic_bytes_per_pair(workload):
    """Compulsory fp32 traffic of the fused chain per pair: the per-frame inputs once, the results once (weights and the cached
    template kernels are shared by the batch)."""
    n = NECK[workload]
    N = n["xf"] - 2 - (n["zf"] - 2) + 1
    N_lp = n["xf_lp"] - 2
    inp = 4 * (LEVELS * C * (n["xf"] ** 2 + n["xf_lp"] ** 2) + 3 * n["img"] ** 2 + 127 * 127 + 16)  # (uint8 crop: 3 * img^2 bytes less 3/4)
    out = 4 * (4 * N * N + 6 * N_lp * N_lp + 3 * n["S"] ** 2 + 127 * 127 + 9) + 2 * 28 + 24
    return {"in": inp, "out": out, "total": inp + out}


def flops_per_pair(workload):
    """Dense work of the fused chain per pair (multiply-add = 2): conv_search, the correlations as direct sums, the 1x1 tails."""
    n = NECK[workload]
    s, k = n["xf"] - 2, n["zf"] - 2
    N = s - k + 1
    sl, kl = n["xf_lp"] - 2, n["zf_lp"] - 2
    conv = NPROB * 2 * C * C * 9 * (s * s + sl * sl)
    corr = NPROB * 2 * C * (N * N * k * k + sl * sl * kl * kl)
    tail = NPROB * 2 * C * C * (N * N + sl * sl) + LEVELS * 2 * C * (4 * N * N + 6 * sl * sl)
    return {"conv_search": conv, "xcorr_direct_equivalent": corr, "head_tail": tail, "total": conv + corr + tail}
