"""A POT-format benchmark on disk, built from the synthetic homography-walk sequences.

Layout and JSON schema are the reference's (toolkit/benchmarks/POT/generate_json_for_POT.py:20-88, read by
toolkit/datasets/pot.py:70-89):
    <root>/POT210.json   {"V01_1": {"video_dir", "init_rect" [8], "img_names" [<rel path>...], "gt_rect" [[8]...],
                                     "flag" ["0"...], "homography" [[9]...]}, ...}
    <root>/V01/V01_1/img/0001.jpg ...
so `tools/test.py --dataset POT210` (whose dataset root is <tools>/../testing_dataset/POT, tools/test.py:58-62) and the
mirrored `toolkit.datasets.DatasetFactory` iterate it unchanged.  Video names follow POT's V<object>_<motion> scheme.
"""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np

from . import synthetic


def video_name(i):
    return "V%02d_%d" % (i // 7 + 1, i % 7 + 1)


def write_dataset(root, n_sequences=8, n_frames=501, size=(720, 1280), seed0=100, name="POT210", ext="jpg", events=None, workers=8,
                  only=None, merge=True):
    """Write the benchmark under `root` (idempotent: an existing JSON with the same parameters is kept).  -> path of the JSON.
    only: sequence indices this caller renders (ranks render their own sequences in parallel; each leaves a `<name>.part<i>` file);
    merge: assemble the JSON from the parts (all of them must exist: call after a barrier)."""
    os.makedirs(root, exist_ok=True)
    path = os.path.join(root, name + ".json")
    stamp = {"n_sequences": n_sequences, "n_frames": n_frames, "size": list(size), "seed0": seed0, "ext": ext}
    stamp_path = os.path.join(root, name + ".fixture")
    if os.path.exists(path) and os.path.exists(stamp_path) and json.load(open(stamp_path)) == stamp:
        return path
    H, W = size
    meta = {}
    pool = ThreadPoolExecutor(max_workers=workers)
    for s in (range(n_sequences) if only is None else only):
        vname = video_name(s)
        frames, polys = synthetic.sequence(seed0 + s, n_frames, size=size, obj=(H // 3, W // 3), events=events)
        rel_dir = os.path.join(vname.split("_")[0], vname, "img")
        os.makedirs(os.path.join(root, rel_dir), exist_ok=True)
        names = [os.path.join(rel_dir, "%04d.%s" % (t + 1, ext)) for t in range(n_frames)]
        list(pool.map(lambda nt: cv2.imwrite(os.path.join(root, nt[0]), nt[1]), zip(names, frames)))
        init = polys[0].reshape(4, 2).astype(np.float32)
        homo = [cv2.getPerspectiveTransform(init, p.reshape(4, 2).astype(np.float32)).reshape(-1).tolist() for p in polys]
        entry = {"video_dir": vname, "init_rect": [float(v) for v in polys[0]], "img_names": names,
                 "gt_rect": [[float(v) for v in p] for p in polys], "flag": ["0"] * n_frames, "homography": homo}
        with open(os.path.join(root, "%s.part%d" % (name, s)), "w") as fh:
            json.dump(entry, fh)
        del frames
    pool.shutdown()
    if not merge:
        return None
    for s in range(n_sequences):
        with open(os.path.join(root, "%s.part%d" % (name, s))) as fh:
            meta[video_name(s)] = json.load(fh)
    with open(path, "w") as fh:
        json.dump(meta, fh)
    with open(stamp_path, "w") as fh:
        json.dump(stamp, fh)
    for s in range(n_sequences):
        os.remove(os.path.join(root, "%s.part%d" % (name, s)))
    return path


def is_current(root, n_sequences, n_frames, size, seed0=100, name="POT210", ext="jpg"):
    stamp_path = os.path.join(root, name + ".fixture")
    want = {"n_sequences": n_sequences, "n_frames": n_frames, "size": list(size), "seed0": seed0, "ext": ext}
    return os.path.exists(os.path.join(root, name + ".json")) and os.path.exists(stamp_path) and json.load(open(stamp_path)) == want
