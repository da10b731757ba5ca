"""Accuracy statistics of the homography benchmarks -- the NumPy subset of toolkit/utils/statistics.py the planar-tracking
evaluation needs, same names / arguments / return values, no shapely or Cython dependency:

    overlap_ratio           :59-76    IoU of axis-aligned [x, y, w, h] boxes, clipped to [0, 1]
    success_overlap         :99-106   fraction of frames with IoU > t,  t = 0, 0.05, ..., 1
    success_error           :196-204  fraction of frames with centre distance <= t (frames whose ground-truth centre is not
                                      strictly positive keep the reference's -1 sentinel distance, i.e. they pass every t >= 0)
    success_4pts_error      :206-218  alignment error e = sqrt(mean over the 4 corners of the squared distance); fraction with e <= t
    success_centroid_error  :163-194  distance between polygon centroids (an invalid = self-intersecting result polygon has
                                      its centroid at the origin, as in the reference)

The reference takes polygon centroids / validity from shapely; here they are the shoelace centroid and an explicit edge-crossing
test, which is what shapely computes for quadrilaterals.
"""
import numpy as np

OVERLAP_THRESHOLDS = np.arange(0, 1.05, 0.05)
PIXEL_THRESHOLDS = np.arange(0, 51, 1)


def overlap_ratio(rect1, rect2):
    rect1, rect2 = np.asarray(rect1, float), np.asarray(rect2, float)
    w = np.minimum(rect1[:, 0] + rect1[:, 2], rect2[:, 0] + rect2[:, 2]) - np.maximum(rect1[:, 0], rect2[:, 0])
    h = np.minimum(rect1[:, 1] + rect1[:, 3], rect2[:, 1] + rect2[:, 3]) - np.maximum(rect1[:, 1], rect2[:, 1])
    inter = np.maximum(0, w) * np.maximum(0, h)
    union = rect1[:, 2] * rect1[:, 3] + rect2[:, 2] * rect2[:, 3] - inter
    return np.clip(inter / union, 0, 1)


def _fraction(values, thresholds, n_frame, strict):
    values = np.asarray(values)[None, :]
    t = np.asarray(thresholds)[:, None]
    hit = values > t if strict else values <= t
    return hit.sum(axis=1) / float(n_frame)


def success_overlap(gt_bb, result_bb, n_frame):
    return _fraction(overlap_ratio(gt_bb, result_bb), OVERLAP_THRESHOLDS, n_frame, strict=True)


def success_error(gt_center, result_center, thresholds, n_frame):
    gt_center, result_center = np.asarray(gt_center, float), np.asarray(result_center, float)
    dist = np.full(len(gt_center), -1.0)  # the reference's sentinel for frames without a (positive) ground-truth centre
    valid = (gt_center > 0).sum(axis=1) == 2
    dist[valid] = np.sqrt(((gt_center[valid] - result_center[valid]) ** 2).sum(axis=1))
    return _fraction(dist, thresholds, n_frame, strict=False)


def alignment_error(gt_poly, poly):
    """[N, 8] corner lists -> [N]: root of the mean squared corner distance (the benchmark's e_AL)."""
    d = np.asarray(gt_poly, float) - np.asarray(poly, float)
    return np.sqrt((d ** 2).sum(axis=1) / 4)


def success_4pts_error(gt_poly, poly, thresholds, n_frame):
    return _fraction(alignment_error(gt_poly, poly), thresholds, n_frame, strict=False)


def polygon_centroid(pts):
    """Area centroid of a simple polygon [K, 2] (shoelace); the vertex mean if the area vanishes."""
    pts = np.asarray(pts, float)
    x, y = pts[:, 0], pts[:, 1]
    xn, yn = np.roll(x, -1), np.roll(y, -1)
    cross = x * yn - xn * y
    area2 = cross.sum()
    if abs(area2) < 1e-12:
        return pts.mean(axis=0)
    return np.array([((x + xn) * cross).sum(), ((y + yn) * cross).sum()]) / (3.0 * area2)


def _segments_cross(p1, p2, p3, p4):
    def orient(a, b, c):
        return np.sign((b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]))
    return orient(p1, p2, p3) * orient(p1, p2, p4) < 0 and orient(p3, p4, p1) * orient(p3, p4, p2) < 0


def polygon_is_simple(quad):
    """A quadrilateral is valid (simple) unless its two pairs of opposite edges cross or it has no area."""
    q = np.asarray(quad, float).reshape(4, 2)
    if _segments_cross(q[0], q[1], q[2], q[3]) or _segments_cross(q[1], q[2], q[3], q[0]):
        return False
    x, y = q[:, 0], q[:, 1]
    return abs((x * np.roll(y, -1) - np.roll(x, -1) * y).sum()) > 1e-12


def success_centroid_error(gt_poly, res_poly, thresholds, n_frame):
    gt_poly, res_poly = np.asarray(gt_poly, float), np.asarray(res_poly, float)
    cg = np.array([polygon_centroid(g.reshape(4, 2)) for g in gt_poly])
    cr = np.array([polygon_centroid(r.reshape(4, 2)) if polygon_is_simple(r) else np.zeros(2) for r in res_poly])
    return _fraction(np.sqrt(((cg - cr) ** 2).sum(axis=1)), thresholds, n_frame, strict=False)
