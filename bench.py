#!/usr/bin/env python
"""bench.py -- frames/sec of the corr+warp+DLT forward (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload 256/512|127/255|win15|stream] [--batch B]

N > 1 is launched by torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.  Default workload = BASELINE configs[1]
(batch 64 synthetic 256/512 pairs per GPU); a step is one pass of the hot path over one batch.

  value      M1 chain, device-resident: 6xK1 + 6xK2 + K3 + K5/K4 + 2xK6 on head features already in HBM (hdn_b200.engine.M1Engine),
             all ranks, max-over-ranks CUDA-event time.  Same scope as round 1.
  e2e        the path through the plug-in boundary with HOST buffers (hdn_b200.head_engine.HeadEngine.run_host): per-frame NECK
             features + crops in pinned host memory -> fused BAN heads (conv_search, K1/K2, 1x1 tail, level sum + K6), K3, K5/K4 ->
             results in pinned host memory; H2D / D2H inside the timed region.  (Round 1 took post-conv_search head features from
             the host -- 54 MB / pair, PCIe-bound; the boundary now sits where the tensors are born on the device: 18 MB / pair.)
             `e2e_m1` keeps the round-1 boundary for continuity.
  fused      the e2e workload with inputs resident in HBM (kernel-only rate of the fused chain).
  roofline   the dominant HBM-side kernel (the 6-problem K1 launch): algorithmic bytes / its average CUDA-event duration inside
             the timed region vs the measured HBM copy peak; `traffic` = dram bytes of that launch measured by ncu in this run.
  cpu_baseline  oracle/torch_port.py (the reference's own torch calls) on this box's host cores: the fused chain (e2e's scope) and
             the M1 chain (value's scope), all threads and one thread, median of the repeats.

--impl reference times that CPU port alone (rank 0 only): `value` = M1 chain (like-for-like with the GPU arm's `value`),
`e2e.value` = fused chain from neck features (like-for-like with the GPU arm's `e2e`).
--workload win15 (config 5) and --workload stream (config 4) print the same kind of line for those configurations.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec (corr+warp+DLT forward)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="256/512", choices=["256/512", "127/255", "win15", "stream"])
    ap.add_argument("--batch", type=int, default=None, help="pairs per GPU (default 64; 256 for win15)")
    ap.add_argument("--shared-template", type=int, default=-1, help="1 = one template for the whole batch (default at N>1: config 3)")
    ap.add_argument("--chunk", type=int, default=8, help="pairs per chunk of the fused chain (intermediates of a chunk stay in L2)")
    ap.add_argument("--chunks", type=int, default=12, help="pipeline depth of the round-1 boundary's end-to-end path (e2e_m1)")
    ap.add_argument("--f32-crop", action="store_true", help="e2e: upload the search crop widened to fp32 (as the reference does) instead of the uint8 it is")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-m1", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-m2", action="store_true", help="skip the full-forward (M2) block")
    ap.add_argument("--m2-batch", type=int, default=None, help="pairs per M2 step (default: --batch)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu dram-traffic measurement of the dominant kernel")
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="time budget of the cpu_baseline leg")
    ap.add_argument("--xcorr-algo", default="auto", choices=["auto", "direct", "fft", "fft_phased", "fft_pipe", "fft_ws"], help="auto = FFT correlation where it beats the direct sum")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    # --workload stream (config 4)
    ap.add_argument("--sequences", type=int, default=8)
    ap.add_argument("--frames", type=int, default=501)
    ap.add_argument("--frame-size", default="720x1280")
    ap.add_argument("--lockstep", type=int, default=0, help="stream: sequences advanced together per GPU (0 = one at a time)")
    ap.add_argument("--host-preproc", action="store_true", help="stream: keep the frame pre-processing on the host (OpenCV), as in round 1")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("bf16_tflops_sustained", 0) or 0)
    return 6650.0, "fallback (B200_PROFILING.md)", 1400.0


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        good = [(t, r) for t, r in self.rows if len(r) >= 7]
        rows = [r for t, r in good if t0 <= t <= t1 + 0.1]
        if not rows and good:  # timed region shorter than the sampling period: take the sample nearest to it
            rows = [min(good, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))[1]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        f = lambda s: float(s) if s.replace(".", "", 1).isdigit() else float("nan")  # noqa: E731
        return {"sm_mhz": statistics.median(f(r[0]) for r in rows), "sm_max_mhz": f(rows[0][1]), "power_w_max": max(f(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------ CPU arm
def _median_rate(run, pairs, steps, warmup):
    """-> (pairs / median step seconds, median ms, spread = (max - min) / median of the step times)"""
    for _ in range(warmup):
        run()
    ts = []
    for _ in range(steps):
        t = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t)
    med = statistics.median(ts)
    return pairs / med, 1e3 * med, (max(ts) - min(ts)) / med if med > 0 else 0.0


def cpu_m1_runner(workload, n, shared=False):
    """The M1 chain (value's scope) on CPU tensors: oracle/torch_port.m1_chain on n pairs."""
    import torch
    from hdn_b200 import engine
    from oracle import c_oracle, torch_port
    w = engine.WORKLOADS[workload]
    f = engine.make_inputs(workload, n, seed=1, shared_template=shared)
    if w["lp_x"]:
        M, Mi = c_oracle.default_M(127, 127)
        f.update(S=w["S"], M=torch.from_numpy(M), Minv=torch.from_numpy(Mi), polar=None)
        return lambda: torch_port.m1_chain(f)
    return lambda: [torch_port.xcorr_depthwise(x, k) for x, k in zip(f["xs"], f["ks"])]


def cpu_fused_runner(workload, n, shared=False, hoist=False):
    """The fused chain (e2e's scope) on CPU tensors: both BAN heads from neck features + K6 + K3 + K5/K4 (oracle/torch_port.fused_chain).
    hoist=False: the template-side `conv_kernel(z_f)` is recomputed every call, as the reference does (ban.py:74)."""
    import numpy as np
    import torch
    from hdn_b200 import head_engine as he
    from hdn_b200.engine import WIN_INFL
    from oracle import c_oracle, torch_port
    f = he.make_inputs(workload, n, seed=1, shared_template=shared)
    gs, gl = he.GAINS[workload]
    w_sim, w_lp = he.HeadWeights.synthetic(11, 2, gs).raw, he.HeadWeights.synthetic(12, 4, gl).raw
    M, Mi = c_oracle.default_M(127, 127)
    f.update(S=he.NECK[workload]["S"], M=torch.from_numpy(M), Minv=torch.from_numpy(Mi))
    N = he.NECK[workload]["xf"] - 2 - (he.NECK[workload]["zf"] - 2) + 1
    win = np.outer(np.hanning(N), np.hanning(N)).flatten()
    kernels = (torch_port.template_kernels(w_sim, f["zf"]), torch_port.template_kernels(w_lp, f["zf_lp"])) if hoist else None
    return lambda: torch_port.fused_chain(f, w_sim, w_lp, win, WIN_INFL, kernels)


def cpu_rates(workload, budget_s, steps=5, warmup=1, pairs=64, shared=False, single_thread=True):
    """CPU port timed on the host cores within ~budget_s seconds.  -> dict for the JSON line."""
    import torch
    cores = host_threads()
    torch.set_num_threads(cores)
    out = {"cores": cores, "kind": "port", "unit": UNIT, "torch": torch.__version__}
    fused = workload in ("256/512", "127/255")
    legs = [("m1", cpu_m1_runner)] + ([("fused", cpu_fused_runner)] if fused else [])
    share = budget_s / (len(legs) + (0.5 if single_thread else 0))
    for name, make in legs:
        probe = make(workload, 1, shared)
        probe()
        t = time.perf_counter()
        probe()
        t_pair = max(time.perf_counter() - t, 1e-4)
        # batched calls amortise: assume a pair costs ~60 % of the single-pair call when sizing the sample
        n = max(1, min(pairs, int(share / (0.6 * t_pair) / (steps + warmup))))
        rate, ms, spread = _median_rate(make(workload, n, shared), n, steps, warmup)
        out[name] = {"value": rate, "ms_per_step": ms, "pairs_per_step": n, "steps": steps, "spread": spread, "same_config": n == pairs}
    if single_thread:
        torch.set_num_threads(1)
        name, make = legs[-1]
        n1 = 1
        rate, ms, spread = _median_rate(make(workload, n1, shared), n1, 3, 1)
        out["single_thread"] = {"scope": name, "value": rate, "ms_per_step": ms, "pairs_per_step": n1, "steps": 3}
        torch.set_num_threads(cores)
    return out


def workload_name(workload, scope="m1"):
    if workload == "win15":
        return "15x15 window K1 only (config 5)"
    if workload == "stream":
        return "POT-shaped stream (config 4): hdnTrackerHomo.init/track_new per sequence, native 127/255 crops"
    if scope == "fused":
        return "%s crops: fused BAN heads from neck features (6x conv_search + 6xK1 + 6x 1x1 tail, same for the lp branch with K2) + K3 + K5/K4 + 2xK6 per pair" % workload
    return "%s crops: 6xK1 + 6xK2 + K3 + K5/K4 + 2xK6 per pair" % workload


def cpu_m2_rate(workload, repeats=3):
    """Model-level CPU forward of ONE pair: the mirrored ModelBuilder on CPU tensors with its operators backed
    by oracle/torch_port.py (oracle/cpu_shim.py) -- track_new + track_new_lp + track_proj of a 256/512 (or 127/255) crop."""
    import torch
    from oracle import cpu_shim
    from hdn_b200 import synthetic
    ex, inst = (256, 512) if workload == "256/512" else (127, 255)
    model, cfg = cpu_shim.build_cpu_model(inst, ex)
    z = torch.from_numpy(synthetic.crop_tensor(1, (1, 6, ex, ex)))
    x = torch.from_numpy(synthetic.crop_tensor(2, (1, 3, inst, inst)))
    pair = torch.randn(1, 2, 127, 127)
    h4p = torch.tensor([[0.0, 0.0, 0.0, 127.0, 127.0, 127.0, 127.0, 0.0]])
    with torch.no_grad():
        model.template(z)

        def frame():
            model._stage1_packed(x, cfg.TRACK.WINDOW_INFLUENCE)
            model._stage2_packed(x)
            model._stage3_packed(pair, h4p)
        rate, ms, spread = _median_rate(frame, 1, repeats, 1)
    return {"value": rate, "unit": UNIT, "ms_per_step": ms, "pairs_per_step": 1, "steps": repeats, "spread": spread,
            "scope": "full forward of one %s pair on CPU (mirrored ModelBuilder, operators = the reference's torch calls)" % workload}


def run_reference_stream(a):
    """CPU arm of config 4: hdnTrackerHomo.init / track_new of the mirrored tracker on CPU tensors (oracle/cpu_shim.py: the
    reference's own torch / NumPy / OpenCV calls) over a bounded sample of one synthetic sequence of the same frame size."""
    import numpy as np
    import torch
    from oracle import cpu_shim
    from hdn_b200 import synthetic
    cores = host_threads()
    torch.set_num_threads(cores)
    H, W = (int(v) for v in a.frame_size.lower().split("x"))
    n = max(3, min(a.steps, 12))
    model, cfg = cpu_shim.build_cpu_model()
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    tracker = build_tracker(model)
    frames, polys = synthetic.sequence(100, n + 2, size=(H, W), obj=(H // 3, W // 3))
    gt = polys[0]
    cx, cy, w, h = get_min_max_bbox(np.array(gt))
    with torch.no_grad():
        tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
        tracker.track_new(1, frames[1], None, None, None)  # warm-up frame
        ts = []
        for i in range(2, n + 2):
            t = time.perf_counter()
            tracker.track_new(i, frames[i], None, None, None)
            ts.append(time.perf_counter() - t)
    med = statistics.median(ts)
    sample = "%d frames of one %dx%d synthetic sequence, mirrored tracker on CPU (oracle/cpu_shim.py), torch %s, %d threads, median frame" % (
        n, W, H, torch.__version__, cores)
    line = {"impl": "reference", "metric": "frames/sec (hdnTrackerHomo.init/track_new, POT-shaped stream)", "value": 1.0 / med, "unit": UNIT,
            "n_gpus": a.gpus, "steps": n, "warmup": 1, "ms_per_step": 1e3 * med, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name("stream"), "frame_size": [H, W], "sample_frames": n, "parallelism": "host cores of rank 0 (CPU arm)"},
            "cpu_baseline": {"value": 1.0 / med, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "spread": (max(ts) - min(ts)) / med},
            "e2e": {"value": 1.0 / med, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference(a, rank):
    if rank != 0:
        return
    if a.workload == "stream":
        return run_reference_stream(a)
    B = a.batch or (256 if a.workload == "win15" else 64)
    # K steps, each a bounded sample; the whole run within a few minutes
    r = cpu_rates(a.workload, budget_s=150.0, steps=max(a.steps, 1), warmup=max(a.warmup, 1), pairs=B, single_thread=False)
    m1, fused = r["m1"], r.get("fused")
    e2e_leg = fused or m1
    m2 = None
    if fused and not a.no_m2:
        try:
            m2 = cpu_m2_rate(a.workload)
        except Exception as e:
            m2 = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    sample = "%d step(s) x %d pair(s) of the M1 chain%s, torch %s CPU fp32 (oracle/torch_port.py), %d threads, median step" % (
        m1["steps"], m1["pairs_per_step"], (" / x %d pair(s) of the fused chain from neck features" % fused["pairs_per_step"]) if fused else "",
        r["torch"], r["cores"])
    line = {"impl": "reference", "metric": METRIC, "value": m1["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": m1["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.workload), "e2e_workload": workload_name(a.workload, "fused") if fused else workload_name(a.workload),
                       "value_scope": "M1 chain on head features (like-for-like with the GPU arm's `value`)",
                       "e2e_scope": "fused chain from neck features (like-for-like with the GPU arm's `e2e`); the template-side conv_kernel is "
                                    "recomputed every call as the reference does (ban.py:74)" if fused else "same as value",
                       "sample_pairs_per_step": m1["pairs_per_step"], "e2e_sample_pairs_per_step": e2e_leg["pairs_per_step"], "pairs_per_gpu": B,
                       "same_config": bool(m1["same_config"] and e2e_leg["same_config"]), "channels": 256, "template": "per pair",
                       "parallelism": "host cores of rank 0 (CPU arm)"},
            "cpu_baseline": {"value": m1["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample, "spread": m1["spread"],
                             "fused": fused, "m2": m2},
            "e2e": {"value": e2e_leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "ms_per_step": e2e_leg["ms_per_step"],
                    "spread": e2e_leg["spread"]},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ dram traffic of the dominant kernel (ncu)
def traffic_child(a):
    """Launched under ncu by measure_traffic(): the K1 launch of the bench configuration, twice (inputs generated on the device)."""
    import torch
    from hdn_b200 import engine, ops
    B = a.batch or (256 if a.workload == "win15" else 64)
    w = engine.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    ops.set_xcorr_algo(a.xcorr_algo)
    xs = [torch.randn(B, engine.C, w["sim_x"], w["sim_x"], device=dev) for _ in range(engine.NPROB)]
    ks = [torch.randn(B, engine.C, w["sim_k"], w["sim_k"], device=dev) * 0.1 for _ in range(engine.NPROB)]
    for _ in range(2):
        ops.xcorr_depthwise_multi(xs, ks)
    torch.cuda.synchronize()


def measure_traffic(a):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE K1 launch at the bench configuration, from an ncu pass over a child
    process of this very command (B200_PROFILING.md).  -> (bytes | None, how)"""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base", "-k", "regex:xcorr", "-s", "1", "-c", "1", "--csv",
           sys.executable, os.path.abspath(__file__), "--traffic-child", "--workload", a.workload, "--xcorr-algo", a.xcorr_algo] + (["--batch", str(a.batch)] if a.batch else [])
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    except (OSError, subprocess.TimeoutExpired) as e:
        return None, "ncu failed: %s" % type(e).__name__
    total, seen = 0.0, 0
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith('"')]
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        name, val, unit = row.get("Metric Name", ""), row.get("Metric Value", ""), row.get("Metric Unit", "")
        if name.startswith("dram__bytes_") and val:
            v = float(val.replace(",", ""))
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            total += v
            seen += 1
    if seen < 2:
        return None, "ncu gave no dram metrics (rc %d): %s" % (res.returncode, (res.stderr or res.stdout).strip().splitlines()[-1][:160] if (res.stderr or res.stdout).strip() else "")
    return total, "ncu dram__bytes_read.sum + dram__bytes_write.sum, one K1 launch, measured in this run"


# ------------------------------------------------------------------------------------------ M2: the full forward
M2_GFLOP_PER_PAIR = {"256/512": 510.0, "127/255": 124.0}  # SURVEY 8(d) [probed with FlopCounterMode on the reference]: backbones + necks + heads + homography net


def m2_block(workload, B, steps=3, warmup=2):
    """SURVEY 8(d) "M2": the three network stages of a frame (ModelBuilder.track_new / track_new_lp / track_proj with their K6
    epilogues) for a batch of B crops through the mirrored model -- ResNet-50 x2, necks, fused BAN heads, ResNet-34 homography net,
    K3, K5/K4 -- with random-init weights of the architecture.  -> frames/s and dense TFLOP/s (fp32 accuracy: 3xTF32 on tcgen05 for
    the stride-1 layers, cuDNN fp32 for the stem / strided layers) against the measured bf16 peak."""
    import torch
    from hdn_b200 import compat, synthetic
    compat.activate()
    from hdn.core.config import cfg
    cfg.merge_from_file(os.path.join(ROOT, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml"))
    ex, inst = (256, 512) if workload == "256/512" else (127, 255)
    cfg.TRACK.INSTANCE_SIZE, cfg.TRACK.EXEMPLAR_SIZE = inst, ex
    cfg.CUDA = True
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    model = synthetic.fill_weights(ModelBuilder()).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.rand((B, 6, ex, ex), device="cuda", generator=g) * 255.0
    x = torch.rand((B, 3, inst, inst), device="cuda", generator=g) * 255.0
    pair = torch.randn((B, 2, 127, 127), device="cuda", generator=g)
    h4p = torch.tensor([[0.0, 0.0, 0.0, 127.0, 127.0, 127.0, 127.0, 0.0]], device="cuda").repeat(B, 1)
    with torch.no_grad():
        model.template(z)  # once per sequence, outside the per-frame step

        def frame():
            a_ = model._stage1_packed(x, cfg.TRACK.WINDOW_INFLUENCE)
            b_ = model._stage2_packed(x)
            c_ = model._stage3_packed(pair, h4p)
            return a_, b_, c_

        for _ in range(warmup):
            frame()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = frame()
        e1.record()
        torch.cuda.synchronize()
        # where the step goes: one extra pass with an event after every stage / sub-module
        marks = [("start", torch.cuda.Event(enable_timing=True))]
        marks[0][1].record()

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

        feats = model.backbone(x); mark("backbone(search crop)")
        xf = model.neck(feats); mark("neck")
        model.head.fused(xf, model._k_sim, model._window(model.head.out_size(xf, model._k_sim), x.device), cfg.TRACK.WINDOW_INFLUENCE, want_maps=False)
        mark("fused BAN head + K6")
        model._stage2_packed(x); mark("stage 2: K3 + backbone(log-polar crop) + neck + fused head_lp + K6")
        model._stage3_packed(pair, h4p); mark("stage 3: ShareFeature + ResNet-34 + K5/K4 + scores")
        torch.cuda.synchronize()
        stages = {marks[i + 1][0]: marks[i][1].elapsed_time(marks[i + 1][1]) for i in range(len(marks) - 1)}
    ms = e0.elapsed_time(e1) / steps
    _, _, bf16 = peaks()
    tf = M2_GFLOP_PER_PAIR[workload] * B / ms  # GFLOP per ms = TFLOP/s
    del model
    torch.cuda.empty_cache()
    return {"value": B / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "pairs_per_step": B, "steps": steps, "gflop_per_pair": M2_GFLOP_PER_PAIR[workload],
            "tflops": tf, "stage_ms": stages, "bf16_peak_tflops_sustained": bf16, "frac_of_bf16_peak": tf / bf16 if bf16 else None,
            "precision": "fp32-accurate: 3xTF32 on tcgen05 (3 tensor-core MACs per MAC) for every 1x1 / 3x3 convolution of both ResNets, the necks and "
                         "the heads; fp32 FMA direct sums for the 7x7 stems and PreShareFeature; no cuDNN convolution",
            "check": float(out[0].float().sum().item() * 0 + out[2][0, 8].item())}


# ------------------------------------------------------------------------------------------ GPU arm: chain workloads
def main():
    a = parse()
    if a.traffic_child:
        return traffic_child(a)
    from hdn_b200 import shard
    rank, local_rank, world = shard.env_world()
    if a.impl == "reference":
        return run_reference(a, rank)

    import torch
    from hdn_b200 import _lib, engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    if a.workload == "stream":
        from hdn_b200 import stream_bench
        return stream_bench.run(a)
    shard.init()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = a.batch or (256 if a.workload == "win15" else 64)
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        sampler.start()  # before the (slow) input generation so nvidia-smi is already streaming when timing starts
    shared = (world > 1) if a.shared_template < 0 else bool(a.shared_template)
    full = a.workload != "win15"

    host_in = engine.make_inputs(a.workload, B, seed=1 + rank, shared_template=shared, pin=not a.no_e2e_m1 and not a.no_e2e)
    eng = engine.M1Engine(a.workload, B, dev, shared_template=shared, use_graph=False)
    dev_in = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in host_in.items()}
    if shared and world > 1:  # config 3: ONE NCCL broadcast of rank 0's template pack
        shard.broadcast_template_pack(dev_in["ks"] + dev_in.get("kl", []), src=0)
    eng.bind(dev_in)
    L = _lib.lib()
    from hdn_b200 import ops
    ops.set_xcorr_algo(a.xcorr_algo)
    k1_fft = bool(L.hdn_xcorr_uses_fft(engine.C, eng.w["sim_x"], eng.w["sim_x"], eng.w["sim_k"], eng.w["sim_k"], 0,
                                       0 if shared and B > 1 else engine.C * eng.w["sim_k"] ** 2))

    # split M1Engine._launch into its individual launches so that each gets its own CUDA-event pair inside the timed region
    import ctypes
    vp = ctypes.c_void_p
    w = eng.w
    arr = vp * engine.NPROB
    p = lambda t: vp(t.data_ptr())  # noqa: E731
    st = lambda: vp(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    kbs1 = 0 if shared and B > 1 else engine.C * w["sim_k"] ** 2
    inp, out = eng.inp, eng.out
    # (a template shared by the batch -- config 3 -- was given its row spectra once in eng.bind, outside the step like the broadcast)
    if "ks_spec" in inp:
        launchers = [("k1", lambda: _lib.check(L.hdn_xcorr_dw_multi_spec_f32(engine.NPROB, arr(*[t.data_ptr() for t in inp["xs"]]),
                                                                               arr(*[t.data_ptr() for t in inp["ks_spec"]]), arr(*[t.data_ptr() for t in out["corr"]]),
                                                                               B, engine.C, w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"], 0, st()), "K1"))]
    else:
        launchers = [("k1", lambda: _lib.check(L.hdn_xcorr_dw_multi_f32(engine.NPROB, arr(*[t.data_ptr() for t in inp["xs"]]), arr(*[t.data_ptr() for t in inp["ks"]]),
                                                                          arr(*[t.data_ptr() for t in out["corr"]]), B, engine.C, w["sim_x"], w["sim_x"],
                                                                          w["sim_k"], w["sim_k"], 0, kbs1, st()), "K1"))]
    if full:
        kbs2 = 0 if shared and B > 1 else engine.C * w["lp_k"] ** 2
        if "kl_spec" in inp:
            k2 = lambda: _lib.check(L.hdn_xcorr_dw_multi_spec_f32(engine.NPROB, arr(*[t.data_ptr() for t in inp["xl"]]), arr(*[t.data_ptr() for t in inp["kl_spec"]]),  # noqa: E731
                                                                  arr(*[t.data_ptr() for t in out["corr_lp"]]), B, engine.C, w["lp_x"], w["lp_x"], w["lp_k"], w["lp_k"],
                                                                  1, st()), "K2")
        else:
            k2 = lambda: _lib.check(L.hdn_xcorr_dw_multi_f32(engine.NPROB, arr(*[t.data_ptr() for t in inp["xl"]]), arr(*[t.data_ptr() for t in inp["kl"]]),  # noqa: E731
                                                             arr(*[t.data_ptr() for t in out["corr_lp"]]), B, engine.C, w["lp_x"], w["lp_x"], w["lp_k"],
                                                             w["lp_k"], 1, kbs2, st()), "K2")
        launchers += [
            ("k2", k2),
            ("k3", lambda: _lib.check(L.hdn_logpolar_f32(p(inp["img"]), None, 0.0, p(out["x_lp"]), B, 3, w["img"], w["img"], w["S"], st()), "K3")),
            ("k5k4", lambda: _lib.check(L.hdn_dlt_warp_f32(p(inp["src"]), p(inp["off"]), p(inp["gray"]), None, None, p(out["H"]), p(out["warp"]), B, 1, 127,
                                                           127, st()), "K5+K4")),
            ("k6", lambda: _lib.check(L.hdn_score_argmax_f32(p(inp["cls"]), p(inp["loc"]), p(eng.window), engine.WIN_INFL, p(out["idx"]), p(out["pscore"]),
                                                             p(out["score"]), p(out["center"]), B, 2, w["score"], st()), "K6")),
            ("k6lp", lambda: _lib.check(L.hdn_score_argmax_f32(p(inp["cls_lp"]), p(inp["loc_lp"]), None, 0.0, p(out["idx_lp"]), p(out["pscore_lp"]),
                                                               p(out["score_lp"]), p(out["sim_lp"]), B, 4, w["score_lp"], st()), "K6lp")),
        ]

    def step(events):
        cur = torch.cuda.current_stream()
        marks = [torch.cuda.Event(enable_timing=True)]
        marks[0].record(cur)
        for _, fn in launchers:
            fn()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(cur)
            marks.append(ev)
        events.append(marks)
        if world > 1 and full:
            shard.gather_results(eng.inp["off"], eng.out["H"])

    def timed(fn, steps, after=None):
        """barrier + synchronize on both sides, CUDA events, max over ranks -> ms for `steps` calls of fn (+ `after`: joins side streams)"""
        torch.cuda.synchronize()
        shard.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        torch.cuda.synchronize()
        shard.barrier()
        torch.cuda.synchronize()
        return shard.max_over_ranks(e0.elapsed_time(e1), dev)

    W = max(a.warmup, 3)
    # ---- device-resident timed region (M1 chain) -----------------------------------------------------
    for _ in range(W):
        step([])
    launches0 = _lib.launch_count()
    events = []
    t0 = sampler.mark()
    ms_total = timed(lambda: step(events), a.steps)
    t1 = sampler.mark()
    launches = _lib.launch_count() - launches0
    ms_per_step = ms_total / a.steps
    value = world * B * a.steps / (ms_total * 1e-3)
    clocks = sampler.summary(t0, t1) if rank == 0 else None
    kern_ms = {}
    for marks in events:
        for i, (n, _) in enumerate(launchers):
            kern_ms.setdefault(n, []).append(marks[i].elapsed_time(marks[i + 1]))
    kern_avg = {n: sum(v) / len(v) for n, v in kern_ms.items()}

    # ---- fused chain from neck features: device-resident, then end to end through pinned host buffers -------------
    e2e = fused = e2e_m1 = None
    if full and not a.no_e2e:
        from hdn_b200 import head_engine as he
        hhost = he.make_inputs(a.workload, B, seed=101 + rank, shared_template=shared, pin=True, u8_crop=not a.f32_crop)
        heng = he.HeadEngine(a.workload, B, dev, chunk=a.chunk)
        up = lambda v: [t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)  # noqa: E731
        zf, zf_lp = up(hhost["zf"]), up(hhost["zf_lp"])
        if shared and world > 1:
            shard.broadcast_template_pack(zf + zf_lp, src=0)
        heng.set_template(zf, zf_lp)  # once per template: outside the per-frame step, as in the tracker
        heng.bind({k: up(hhost[k]) for k in he.FRAME_KEYS})

        def fused_step():
            heng.run()
            if world > 1:
                shard.gather_results(heng.inp["off"], heng.out["H"])

        for _ in range(W):
            fused_step()
        n0 = _lib.launch_count()
        ms_f = timed(fused_step, a.steps)
        fl = he.flops_per_pair(a.workload)
        fused = {"value": world * B * a.steps / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f / a.steps, "chunk_pairs": heng.chunk,
                 "gpu_launches": int(_lib.launch_count() - n0), "conv_tflops_fp32_equivalent": fl["conv_search"] * B * a.steps / (ms_f * 1e-3) / 1e12,
                 "gflop_per_pair": fl["total"] / 1e9}
        h2d, d2h = heng.alloc_host_io(hhost)

        def e2e_step():
            # a stream of batches: step i + 1 starts uploading while step i's last chunks compute (results alternate between two
            # pinned result sets); the timed region ends with finish() + synchronize, i.e. with every result in host memory
            heng.run_host(hhost, wait=False)
            if world > 1:
                torch.cuda.current_stream().wait_stream(heng.s_cmp)
                shard.gather_results(heng.inp["off"], heng.out["H"])

        for _ in range(W):
            e2e_step()
        heng.finish()
        n0 = _lib.launch_count()
        ms_e = timed(e2e_step, a.steps, after=heng.finish)
        e2e = {"value": world * B * a.steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e / a.steps, "chunk_pairs": heng.chunk, "gpu_launches": int(_lib.launch_count() - n0),
               "h2d_gbs_per_gpu": h2d / (ms_e / a.steps * 1e-3) / 1e9, "crop_dtype": "f32" if a.f32_crop else "u8",
               "scope": "per-frame neck features + crops in pinned host memory -> fused BAN heads, K3, K5/K4, K6 -> results in pinned host memory "
                        "(HeadEngine.run_host); template kernels cached on the device by set_template, as in the tracker",
               "check": float(heng.host_out["cls"][0, 0, 0, 0])}  # a value read back on the host
        del heng, hhost
    if not a.no_e2e and (not full or not a.no_e2e_m1):
        # round-1 boundary (post-conv_search head features from the host): kept for continuity; the only e2e of the K1-only workload
        h2d, d2h = eng.alloc_host_io(host_in)
        for _ in range(W):
            eng.run_host(host_in, a.chunks)

        def m1_host_step():
            eng.run_host(host_in, a.chunks)
            if world > 1 and full:
                shard.gather_results(eng.dev_in["off"], eng.out["H"])

        ms_e = timed(m1_host_step, a.steps)
        e2e_m1 = {"value": world * B * a.steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "ms_per_step": ms_e / a.steps, "chunks": a.chunks, "scope": "round-1 boundary: head features in / correlation maps out (M1Engine.run_host)",
                  "check": float(eng.host_out["corr"][0][0, 0, 0, 0])}
        if e2e is None:
            e2e, e2e_m1 = e2e_m1, None
    if rank == 0:
        sampler.stop()
    if rank != 0:
        return

    # ---- roofline of the dominant HBM-side kernel --------------------------------------------------------
    peak, peak_src, _ = peaks()
    ab = engine.algorithmic_bytes_per_pair(a.workload, B, shared)
    k1_bytes = ab["k1"] * B
    k1_flops = engine.NPROB * engine.xcorr_flops(w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"], False) * B
    k1_s = kern_avg["k1"] * 1e-3
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    achieved = k1_bytes / k1_s / 1e9
    traffic, traffic_src = (None, "skipped (--no-traffic)") if (a.no_traffic or world > 1 and shared) else measure_traffic(a)
    roofline = {"kernel": "%s (K1, 6 problems/launch)" % ("xcorr_fft_kernel: transform-domain correlation (row FFTs + per-frequency column correlation)" if k1_fft else "xcorr_staged_kernel: direct sum"),
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": k1_bytes, "launch_ms": kern_avg["k1"],
                # flops of the DIRECT sum (2*C*Ho*Wo*h*w); the FFT kernel delivers the same result with ~5x fewer operations, so for it
                # this is a direct-equivalent rate (it may exceed the FMA peak) and fp32_frac says how far past the direct kernel's bound it is
                "fp32_tflops": k1_flops / k1_s / 1e12, "fp32_peak_tflops": fp32_peak, "fp32_frac": k1_flops / k1_s / 1e12 / fp32_peak,
                "fp32_is_direct_equivalent": k1_fft, "flop_per_byte": k1_flops / k1_bytes,
                "chain_gbs": ab["total"] * B / (ms_per_step * 1e-3) / 1e9, "chain_frac": ab["total"] * B / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernel_ms": kern_avg,
                "kernel_frac": {n: (ab[key] * B / (kern_avg[n] * 1e-3) / 1e9 / peak) for n, key in (("k1", "k1"), ("k2", "k2"), ("k3", "k3")) if n in kern_avg and key in ab}}

    m2 = None
    if full and world == 1 and not a.no_m2:
        try:
            m2 = m2_block(a.workload, a.m2_batch or B)
        except Exception as e:  # the M2 block is informative; the line stands without it
            m2 = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    cpu = None
    if not a.no_cpu and world == 1:
        r = cpu_rates(a.workload, a.cpu_seconds, pairs=B)
        lead = r.get("fused") or r["m1"]
        cpu = {"value": lead["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": "%s chain, %d step(s) x %d pair(s), torch %s CPU fp32 (oracle/torch_port.py), %d threads, median step" % (
                   "fused (e2e scope: BAN heads from neck features + K3 + K5/K4 + K6)" if "fused" in r else "M1", lead["steps"], lead["pairs_per_step"],
                   r["torch"], r["cores"]),
               "spread": lead["spread"], "same_config": lead["same_config"], "m1_chain": r["m1"], "single_thread": r.get("single_thread")}
        if full and not a.no_m2:
            try:
                cpu["m2"] = cpu_m2_rate(a.workload, repeats=2)
            except Exception as e:
                cpu["m2"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.workload),
                       "e2e_workload": workload_name(a.workload, "fused") if (full and fused) else workload_name(a.workload),
                       "value_scope": "M1 chain on head features resident in HBM (round-1 scope)",
                       "pairs_per_gpu": B, "global_batch": B * world, "channels": engine.C,
                       "template": ("shared, NCCL broadcast from rank 0" + ("; its row spectra cached once per template" if "ks_spec" in inp else "")) if shared else "per pair",
                       "parallelism": "pairs sharded over %d GPU(s), no data-path collective" % world,
                       "l2": "inputs per step (%.2f GB) exceed the 126 MB L2; no flush needed" % (ab["total"] * B / 1e9)},
            "clocks": clocks, "e2e": e2e, "fused": fused, "e2e_m1": e2e_m1, "m2": m2, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
