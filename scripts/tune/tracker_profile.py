"""Where does a tracker frame spend its time?  Wraps the cv2 / model entry points with timers (dev tool)."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cv2
import numpy as np
import torch
from hdn_b200 import runner, synthetic

T = collections.defaultdict(float)
N = collections.defaultdict(int)


def timed(name, fn, sync=False):
    def w(*a, **k):
        if sync:
            torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn(*a, **k)
        if sync:
            torch.cuda.synchronize()
        T[name] += time.perf_counter() - t
        N[name] += 1
        return r
    return w


for f in ("warpPerspective", "warpAffine", "resize", "logPolar", "perspectiveTransform"):
    setattr(cv2, f, timed("cv2." + f, getattr(cv2, f)))
graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tracker, model = runner.build(runner.parse().config if False else os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml"), "", graphs)
for f in ("track_new_scored", "track_new_lp_scored", "track_proj_packed"):
    setattr(model, f, timed("model." + f, getattr(model, f), sync=True))
import hdn.tracker.base_tracker as bt
import hdn.tracker.hdn_tracker_proj_e2e as pe
bt.crop_window = timed("crop_window(total)", bt.crop_window)
pe.crop_window = bt.crop_window
bt.to_model_tensor = timed("to_model_tensor(H2D)", bt.to_model_tensor)
pe.get_search_info = timed("get_search_info", pe.get_search_info)
H, W = 720, 1280
frames, polys = synthetic.sequence(5, 40, size=(H, W), obj=(H // 3, W // 3))
runner.track_sequence(tracker, frames[:4], polys[:4])
T.clear(); N.clear()
print("cv2 threads:", cv2.getNumThreads(), "cpus:", os.cpu_count())
_, dt = runner.track_sequence(tracker, frames, polys)
nf = len(frames) - 1
print("total %.2f ms/frame (graphs=%d)" % (1e3 * dt / nf, graphs))
for k in sorted(T, key=lambda k: -T[k]):
    print("  %-28s %7.2f ms/frame  (%d calls/frame)" % (k, 1e3 * T[k] / nf, round(N[k] / nf)))
