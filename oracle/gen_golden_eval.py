#!/usr/bin/env python
"""Golden vectors of the benchmark statistics: outputs of the REFERENCE's toolkit/utils/statistics.py functions (imported from
/root/reference, unmodified) on seeded trajectories -> tests/golden/eval_stats.npz.  TEST INFRASTRUCTURE ONLY.

The module imports a Cython extension (`region`, not built here) and shapely (absent) at import time; neither is used by the four
functions captured here, so both are stubbed for the import."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HDN_REFERENCE_ROOT", "/root/reference")


def reference_statistics():
    sys.path.insert(0, os.path.join(HERE, "refshim"))  # shapely stub
    pkg = types.ModuleType("toolkit"); pkg.__path__ = [os.path.join(REF, "toolkit")]
    utils = types.ModuleType("toolkit.utils"); utils.__path__ = [os.path.join(REF, "toolkit", "utils")]
    sys.modules.update({"toolkit": pkg, "toolkit.utils": utils, "toolkit.utils.region": types.ModuleType("toolkit.utils.region")})
    import importlib
    return importlib.import_module("toolkit.utils.statistics")


def main():
    st = reference_statistics()
    rng = np.random.default_rng(2024)
    n = 60
    gt_poly = np.cumsum(rng.normal(0, 3, (n, 8)), axis=0) + np.array([100, 100, 300, 110, 310, 260, 90, 250.0])
    noise = rng.normal(0, 1, (n, 8)) * np.linspace(0.2, 25, n)[:, None]  # error grows along the sequence
    res_poly = gt_poly + noise
    res_poly[45:] = 0  # a lost track scores as zeros
    box = lambda p: np.concatenate([p.reshape(-1, 4, 2).min(1), p.reshape(-1, 4, 2).max(1) - p.reshape(-1, 4, 2).min(1)], 1)  # noqa: E731
    gt_bb, res_bb = box(gt_poly), box(res_poly)
    gt_c, res_c = gt_bb[:, :2] + gt_bb[:, 2:] / 2, res_bb[:, :2] + res_bb[:, 2:] / 2
    gt_c[7] = [-3.0, 50.0]  # a frame without a positive ground-truth centre
    thr = np.arange(0, 51, 1)
    out = dict(gt_poly=gt_poly, res_poly=res_poly, gt_bb=gt_bb, res_bb=res_bb, gt_c=gt_c, res_c=res_c, thresholds=thr,
               overlap_ratio=st.overlap_ratio(gt_bb, res_bb), success_overlap=st.success_overlap(gt_bb, res_bb, n),
               success_error=st.success_error(gt_c, res_c, thr, n), success_4pts_error=st.success_4pts_error(gt_poly[1:], res_poly[1:], thr, n - 1))
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "eval_stats.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
