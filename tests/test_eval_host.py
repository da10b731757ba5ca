"""CPU-only: the mirrored benchmark statistics (hdn_b200/compat/toolkit/utils/statistics.py, evaluation/homo_benchmark.py)
against outputs of the reference's own toolkit/utils/statistics.py (tests/golden/eval_stats.npz, oracle/gen_golden_eval.py),
plus the pieces the reference takes from shapely (polygon centroid / validity) against closed forms."""
import os

import numpy as np

from conftest import load_golden
from hdn_b200 import compat

compat.activate()


def test_statistics_match_the_reference():
    from toolkit.utils import statistics as st
    g = load_golden("eval_stats")
    n = len(g["gt_poly"])
    assert np.array_equal(st.overlap_ratio(g["gt_bb"], g["res_bb"]), g["overlap_ratio"])
    assert np.array_equal(st.success_overlap(g["gt_bb"], g["res_bb"], n), g["success_overlap"])
    assert np.array_equal(st.success_error(g["gt_c"], g["res_c"], g["thresholds"], n), g["success_error"])
    assert np.array_equal(st.success_4pts_error(g["gt_poly"][1:], g["res_poly"][1:], g["thresholds"], n - 1), g["success_4pts_error"])


def test_polygon_centroid_and_validity():
    from toolkit.utils import statistics as st
    sq = np.array([[0, 0], [4, 0], [4, 2], [0, 2.0]])
    assert np.allclose(st.polygon_centroid(sq), [2, 1])
    assert np.allclose(st.polygon_centroid(sq[::-1]), [2, 1])  # orientation does not matter
    trap = np.array([[0, 0], [6, 0], [4, 3], [2, 3.0]])     # isosceles trapezoid: centroid height h/3 * (a + 2b)/(a + b)
    assert np.allclose(st.polygon_centroid(trap), [3, 3 / 3 * (6 + 2 * 2) / (6 + 2)])
    assert st.polygon_is_simple(sq.reshape(-1)) and st.polygon_is_simple(trap.reshape(-1))
    bow = np.array([[0, 0], [4, 2], [4, 0], [0, 2.0]])      # self-intersecting "bow tie"
    assert not st.polygon_is_simple(bow.reshape(-1))
    assert not st.polygon_is_simple(np.zeros(8))              # a lost track written as zeros
    gt = np.stack([sq.reshape(-1), sq.reshape(-1)])
    res = np.stack([sq.reshape(-1) + 3.0, bow.reshape(-1)])    # shifted by (3, 3); invalid -> centroid (0, 0)
    curve = st.success_centroid_error(gt, res, np.arange(0, 51), 2)
    d = [np.hypot(3, 3), np.hypot(2, 1)]
    assert np.array_equal(curve, [(np.array(d) <= t).sum() / 2 for t in range(51)])


class _Video:
    def __init__(self, name, gt, preds):
        self.name, self.gt_traj, self.pred_trajs = name, gt, preds


class _Dataset(list):
    tracker_names = ["ours"]
    tracker_path = None


def test_homo_benchmark_curves_and_results_round_trip(tmp_path):
    from toolkit.datasets import Video
    from toolkit.evaluation import HomoBenchmark
    g = load_golden("eval_stats")
    gt, res = g["gt_poly"], g["res_poly"][:50]  # a result shorter than the ground truth is zero-padded (lost track)
    bench = HomoBenchmark(_Dataset([_Video("v0", gt.tolist(), {"ours": res.tolist()})]))
    prec = bench.eval_4pts_precision()["ours"]["v0"]
    padded = np.concatenate([res, np.zeros((10, 8))])
    e = np.sqrt(((gt[1:] - padded[1:]) ** 2).sum(1) / 4)
    assert np.array_equal(prec, [(e <= t).sum() / 59 for t in range(51)])
    assert prec[0] == 0 and prec[-1] < 1 and np.all(np.diff(prec) >= 0)
    succ = bench.eval_bbox_overlap_success("ours")["ours"]["v0"]
    assert succ.shape == (21,) and np.all(np.diff(succ) <= 0)
    cen = bench.eval_centroid_precision(["ours"])["ours"]["v0"]
    assert cen.shape == (51,) and np.all(np.diff(cen) >= 0)
    s = HomoBenchmark.summary({"v0": prec})
    assert set(s) == {"precision@5", "precision@10", "precision@20", "mean_precision"} and s["precision@5"] <= s["precision@20"]
    # results written the way tools/test.py:237-243 writes them are read back by Video.load_tracker
    os.makedirs(tmp_path / "ours")
    with open(tmp_path / "ours" / "v0.txt", "w") as fh:
        for x in res.tolist():
            fh.write(" ".join(str(i) for i in x) + "\n")
    v = Video.__new__(Video)
    v.name, v.gt_traj, v.pred_trajs, v.tracker_names = "v0", gt.tolist(), {}, []
    assert np.array_equal(np.asarray(v.load_tracker(str(tmp_path), "ours", store=False)), res)
    v.load_tracker(str(tmp_path), ["ours"])
    assert v.tracker_names == ["ours"] and np.array_equal(np.asarray(v.pred_trajs["ours"]), res)
