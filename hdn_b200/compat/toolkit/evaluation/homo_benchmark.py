"""HomoBenchmark -- precision / success curves of a planar (homography) tracker, the part of
toolkit/evaluation/homo_benchmark.py that scores the 8-number polygon results tools/test.py writes.

Same constructor and method names / return structure as the reference ({tracker: {video: curve}}):
    eval_4pts_precision        :185-225  alignment-error precision, thresholds 0..50 px, first frame dropped
    eval_centroid_precision    :227-264  centroid-distance precision, thresholds 0..50 px
    eval_bbox_overlap_success  :69-106   IoU success of the polygons' bounding boxes, thresholds 0..1 step 0.05
The dataset is any iterable of videos with `.name`, `.gt_traj` ([N, 8] corner lists), `.pred_trajs` ({tracker: [N, 8]}) and
optionally `.load_tracker(path, tracker)`, plus `.tracker_names` / `.tracker_path` (toolkit.datasets provides them).
"""
import numpy as np

from ..utils.statistics import PIXEL_THRESHOLDS, success_4pts_error, success_centroid_error, success_overlap


class HomoBenchmark:
    def __init__(self, dataset):
        self.dataset = dataset

    @staticmethod
    def convert_points_to_bbox(points):
        """[N, 8] corner lists -> [N, 4] (x, y, w, h) of the axis-aligned bounding box."""
        pts = np.asarray(points, float).reshape(-1, 4, 2)
        lo, hi = pts.min(axis=1), pts.max(axis=1)
        return np.concatenate([lo, hi - lo], axis=1)

    def _trackers(self, eval_trackers):
        if eval_trackers is None:
            eval_trackers = self.dataset.tracker_names
        return [eval_trackers] if isinstance(eval_trackers, str) else list(eval_trackers)

    def _trajectories(self, video, tracker):
        """Ground truth and result of one video, result zero-padded to the ground truth's length (lost tracks score as misses)."""
        gt = np.asarray(video.gt_traj, float)
        if tracker in getattr(video, "pred_trajs", {}):
            res = np.asarray(video.pred_trajs[tracker], float)
        else:
            res = np.asarray(video.load_tracker(self.dataset.tracker_path, tracker, False), float)
        if res.shape[0] < gt.shape[0]:
            res = np.concatenate([res, np.zeros((gt.shape[0] - res.shape[0], 8))], axis=0)
        if hasattr(video, "val_ids"):
            gt, res = gt[video.val_ids], res[video.val_ids]
        return gt, res

    def _evaluate(self, eval_trackers, score):
        return {t: {v.name: score(*self._trajectories(v, t)) for v in self.dataset} for t in self._trackers(eval_trackers)}

    def eval_4pts_precision(self, eval_trackers=None):
        return self._evaluate(eval_trackers, lambda gt, res: success_4pts_error(gt[1:], res[1:], PIXEL_THRESHOLDS, len(gt) - 1))

    def eval_centroid_precision(self, eval_trackers=None):
        return self._evaluate(eval_trackers, lambda gt, res: success_centroid_error(gt, res, PIXEL_THRESHOLDS, len(gt)))

    def eval_bbox_overlap_success(self, eval_trackers=None):
        box = self.convert_points_to_bbox
        return self._evaluate(eval_trackers, lambda gt, res: success_overlap(box(gt), box(res), len(gt)))

    @staticmethod
    def summary(curves, at=(5, 10, 20)):
        """{video: precision curve over 0..50 px} -> mean precision at the given pixel thresholds and the mean over the curve."""
        c = np.mean([np.asarray(v) for v in curves.values()], axis=0)
        out = {"precision@%d" % t: float(c[t]) for t in at}
        out["mean_precision"] = float(c.mean())
        return out
