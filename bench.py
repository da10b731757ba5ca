#!/usr/bin/env python
"""bench.py -- frames/sec of the corr+warp+DLT forward (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 256/512|127/255|win15] [--batch B]

N > 1 is launched by torchrun (one rank per GPU).  A step is one pass of the chain over one batch of
B synthetic pairs per GPU (hdn_b200.engine).  Rank 0 prints ONE JSON line.

  value      device-resident throughput (inputs in HBM), all ranks, max-over-ranks time
  e2e        the same chain through M1Engine.run_host: pinned host buffers in and out, copies inside the timed region
  roofline   dominant kernel (the 6-problem K1 launch): algorithmic bytes / its average CUDA-event duration
             vs the measured HBM copy peak; fp32 FMA figures beside it (the 256/512 shape is FMA-bound)
  cpu_baseline  oracle/torch_port.py (the reference's own torch calls) on this box's host cores, bounded sample

--impl reference times that CPU port alone (rank 0 only).
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
    ap.add_argument("--workload", default="256/512", choices=["256/512", "127/255", "win15"])
    ap.add_argument("--batch", type=int, default=None, help="pairs per GPU (default 64; 256 for win15)")
    ap.add_argument("--shared-template", type=int, default=-1, help="1 = one template for the whole batch (default at N>1: config 3)")
    ap.add_argument("--chunks", type=int, default=12, help="pipeline depth of the end-to-end path (12 measured best: 1,220 vs 1,201 frames/s at 8)")
    ap.add_argument("--h2d-streams", type=int, default=1, choices=[1, 2], help="upload streams of the end-to-end path")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--xcorr-algo", default="auto", choices=["auto", "direct", "fft"], help="auto = FFT correlation where it beats the direct sum")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(workload, pairs, fft=False):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (per pair x pairs), if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        e = d.get(workload + ("#fft" if fft else "")) or (None if fft else d.get(workload))
        if e:
            return e["bytes_per_pair"] * pairs
    return None


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
def cpu_chain_rate(workload, seconds, threads=None, steps=None, warmup=1):
    """Time oracle/torch_port.m1_chain on host cores over a bounded sample.  -> dict(value, cores, sample, ms_per_step, pairs)"""
    import torch
    from hdn_b200 import engine
    from oracle import c_oracle, torch_port

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = engine.WORKLOADS[workload]

    def feats(n):
        f = engine.make_inputs(workload, n, seed=1)
        if w["lp_x"]:
            M, Mi = c_oracle.default_M(127, 127)
            f.update(S=w["S"], M=torch.from_numpy(M), Minv=torch.from_numpy(Mi), polar=None)
        return f

    def run(f):
        if w["lp_x"]:
            return torch_port.m1_chain(f)
        return [torch_port.xcorr_depthwise(x, k) for x, k in zip(f["xs"], f["ks"])]

    f1 = feats(1)
    run(f1)
    t = time.perf_counter()
    run(f1)
    t_pair = time.perf_counter() - t
    if steps is None:  # cpu_baseline leg: one sample sized to ~`seconds`
        n, steps = max(1, min(16, int(seconds / max(t_pair, 1e-4) / 3))), 3
    else:              # --impl reference: K steps, each a bounded sample; whole run within a few minutes
        n = max(1, min(8, int(150.0 / max(t_pair, 1e-4) / max(steps + warmup, 1))))
    f = feats(n)
    for _ in range(warmup):
        run(f)
    t = time.perf_counter()
    for _ in range(steps):
        run(f)
    dt = time.perf_counter() - t
    return {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d step(s) x %d pair(s) of the %s chain, torch %s CPU fp32 (oracle/torch_port.py), %d threads" % (steps, n, workload, torch.__version__, cores),
            "ms_per_step": 1e3 * dt / steps, "pairs": n}


def workload_name(workload):
    return "%s crops: 6xK1 + 6xK2 + K3 + K5/K4 + 2xK6 per pair" % workload if workload != "win15" else "15x15 window K1 only (config 5)"


def run_reference(a, rank):
    if rank != 0:
        return
    r = cpu_chain_rate(a.workload, a.cpu_seconds, steps=a.steps, warmup=max(a.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.workload), "sample_pairs_per_step": r["pairs"], "channels": 256, "template": "per pair",
                       "parallelism": "host cores of rank 0 (CPU arm)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    a = parse()
    from hdn_b200 import shard
    rank, local_rank, world = shard.env_world()
    if a.impl == "reference":
        return run_reference(a, rank)

    import torch
    from hdn_b200 import _lib, engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    shard.init()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = a.batch or (256 if a.workload == "win15" else 64)
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        sampler.start()  # before the (slow) input generation so nvidia-smi is already streaming when timing starts
    shared = (world > 1) if a.shared_template < 0 else bool(a.shared_template)
    full = a.workload != "win15"

    host_in = engine.make_inputs(a.workload, B, seed=1 + rank, shared_template=shared, pin=True)
    eng = engine.M1Engine(a.workload, B, dev, shared_template=shared, use_graph=False)
    dev_in = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in host_in.items()}
    if shared and world > 1:  # config 3: ONE NCCL broadcast of rank 0's template pack
        shard.broadcast_template_pack(dev_in["ks"] + dev_in.get("kl", []), src=0)
    eng.bind(dev_in)
    L = _lib.lib()
    from hdn_b200 import ops
    ops.set_xcorr_algo(a.xcorr_algo)
    k1_fft = bool(L.hdn_xcorr_uses_fft(engine.C, eng.w["sim_x"], eng.w["sim_x"], eng.w["sim_k"], eng.w["sim_k"], 0, engine.C * eng.w["sim_k"] ** 2))

    def step(events=None):
        """One pass; with `events`, bracket every kernel launch with CUDA events on the launching stream."""
        if events is None:
            eng._launch(eng.inp, eng.out, B)
        else:
            names = ["k1", "k2", "k3", "k5k4", "k6", "k6lp"] if full else ["k1"]
            # re-issue the launches one by one so each gets its own event pair
            cur = torch.cuda.current_stream()
            marks = [torch.cuda.Event(enable_timing=True)]
            marks[0].record(cur)
            for i, _ in enumerate(eng_launchers):
                eng_launchers[i]()
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(cur)
                marks.append(ev)
            events.append((names, marks))
        if world > 1 and full:
            shard.gather_results(eng.inp["off"], eng.out["H"])

    # split M1Engine._launch into its individual launches for per-kernel timing
    import ctypes
    vp = ctypes.c_void_p
    w = eng.w
    arr = vp * engine.NPROB
    p = lambda t: vp(t.data_ptr())  # noqa: E731
    st = lambda: vp(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    kbs1 = 0 if shared and B > 1 else engine.C * w["sim_k"] ** 2
    inp, out = eng.inp, eng.out
    eng_launchers = [lambda: _lib.check(L.hdn_xcorr_dw_multi_f32(engine.NPROB, arr(*[t.data_ptr() for t in inp["xs"]]), arr(*[t.data_ptr() for t in inp["ks"]]),
                                                                   arr(*[t.data_ptr() for t in out["corr"]]), B, engine.C, w["sim_x"], w["sim_x"],
                                                                   w["sim_k"], w["sim_k"], 0, kbs1, st()), "K1")]
    if full:
        kbs2 = 0 if shared and B > 1 else engine.C * w["lp_k"] ** 2
        eng_launchers += [
            lambda: _lib.check(L.hdn_xcorr_dw_multi_f32(engine.NPROB, arr(*[t.data_ptr() for t in inp["xl"]]), arr(*[t.data_ptr() for t in inp["kl"]]),
                                                        arr(*[t.data_ptr() for t in out["corr_lp"]]), B, engine.C, w["lp_x"], w["lp_x"], w["lp_k"],
                                                        w["lp_k"], 1, kbs2, st()), "K2"),
            lambda: _lib.check(L.hdn_logpolar_f32(p(inp["img"]), None, 0.0, p(out["x_lp"]), B, 3, w["img"], w["img"], w["S"], st()), "K3"),
            lambda: _lib.check(L.hdn_dlt_warp_f32(p(inp["src"]), p(inp["off"]), p(inp["gray"]), None, None, p(out["H"]), p(out["warp"]), B, 1, 127,
                                                  127, st()), "K5+K4"),
            lambda: _lib.check(L.hdn_score_argmax_f32(p(inp["cls"]), p(inp["loc"]), p(eng.window), engine.WIN_INFL, p(out["idx"]), p(out["pscore"]),
                                                      p(out["score"]), p(out["center"]), B, 2, w["score"], st()), "K6"),
            lambda: _lib.check(L.hdn_score_argmax_f32(p(inp["cls_lp"]), p(inp["loc_lp"]), None, 0.0, p(out["idx_lp"]), p(out["pscore_lp"]),
                                                      p(out["score_lp"]), p(out["sim_lp"]), B, 4, w["score_lp"], st()), "K6lp"),
        ]

    # ---- device-resident timed region --------------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        step([])
    torch.cuda.synchronize()
    shard.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    events = []
    t0 = sampler.mark()
    ev0.record()
    for _ in range(a.steps):
        step(events)
    ev1.record()
    torch.cuda.synchronize()
    shard.barrier()
    torch.cuda.synchronize()
    t1 = sampler.mark()
    launches = _lib.launch_count() - launches0
    ms_total = shard.max_over_ranks(ev0.elapsed_time(ev1), dev)
    ms_per_step = ms_total / a.steps
    value = world * B * a.steps / (ms_total * 1e-3)
    clocks = sampler.summary(t0, t1) if rank == 0 else None

    kern_ms = {}
    for names, marks in events:
        for i, n in enumerate(names):
            kern_ms.setdefault(n, []).append(marks[i].elapsed_time(marks[i + 1]))
    kern_avg = {n: sum(v) / len(v) for n, v in kern_ms.items()}

    # ---- end-to-end timed region (pinned host in -> pinned host out) --------------------------------
    e2e = None
    if not a.no_e2e:
        h2d, d2h = eng.alloc_host_io(host_in)
        eng.h2d_streams = a.h2d_streams
        for _ in range(max(a.warmup, 3)):
            eng.run_host(host_in, a.chunks)
        torch.cuda.synchronize()
        shard.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            eng.run_host(host_in, a.chunks)
            if world > 1 and full:
                shard.gather_results(eng.dev_in["off"], eng.out["H"])
        e1.record()
        torch.cuda.synchronize()
        shard.barrier()
        ms_e2e = shard.max_over_ranks(e0.elapsed_time(e1), dev)
        e2e = {"value": world * B * a.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e2e / a.steps, "chunks": a.chunks,
               "check": float(eng.host_out["corr"][0][0, 0, 0, 0])}  # a value read back on the host
    if rank == 0:
        sampler.stop()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peak, peak_src = peaks()
    ab = engine.algorithmic_bytes_per_pair(a.workload, B, shared)
    k1_bytes = ab["k1"] * B
    k1_flops = engine.NPROB * engine.xcorr_flops(w["sim_x"], w["sim_x"], w["sim_k"], w["sim_k"], False) * B
    k1_s = kern_avg["k1"] * 1e-3
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    achieved = k1_bytes / k1_s / 1e9
    roofline = {"kernel": "%s (K1, 6 problems/launch)" % ("xcorr_fft_kernel: transform-domain correlation (row FFTs + per-frequency column correlation)" if k1_fft else "xcorr_staged_kernel: direct sum"),
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic_from_profiles(a.workload, B, k1_fft),
                "algorithmic_bytes_per_launch": k1_bytes, "launch_ms": kern_avg["k1"],
                # flops of the DIRECT sum (2*C*Ho*Wo*h*w); the FFT kernel delivers the same result with ~5x fewer operations, so for it
                # this is a direct-equivalent rate (it may exceed the FMA peak) and fp32_frac says how far past the direct kernel's bound it is
                "fp32_tflops": k1_flops / k1_s / 1e12, "fp32_peak_tflops": fp32_peak, "fp32_frac": k1_flops / k1_s / 1e12 / fp32_peak,
                "fp32_is_direct_equivalent": k1_fft, "flop_per_byte": k1_flops / k1_bytes,
                "chain_gbs": ab["total"] * B / (ms_per_step * 1e-3) / 1e9, "chain_frac": ab["total"] * B / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernel_ms": kern_avg}

    cpu = None
    if not a.no_cpu and world == 1:
        r = cpu_chain_rate(a.workload, a.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.workload),
                       "pairs_per_gpu": B, "global_batch": B * world, "channels": engine.C,
                       "template": "shared, NCCL broadcast from rank 0" if shared else "per pair",
                       "parallelism": "pairs sharded over %d GPU(s), no data-path collective" % world,
                       "l2": "inputs per step (%.2f GB) exceed the 126 MB L2; no flush needed" % (ab["total"] * B / 1e9)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
