"""Box / polygon helpers used by the tracker and the tools -- mirror of the inference-side functions of
hdn/utils/bbox.py (lines cited per function).  Pure NumPy / OpenCV host code, float64 like the reference.
"""
import math
from collections import namedtuple

import cv2
import numpy as np

Corner = namedtuple("Corner", "x1 y1 x2 y2")
BBox = Corner
Center = namedtuple("Center", "x y w h")


def corner2center(corner):
    """(x1,y1,x2,y2) -> (cx,cy,w,h); bbox.py:22-38."""
    x1, y1, x2, y2 = corner[0], corner[1], corner[2], corner[3]
    vals = ((x1 + x2) * 0.5, (y1 + y2) * 0.5, x2 - x1, y2 - y1)
    return Center(*vals) if isinstance(corner, Corner) else vals


def center2corner(center):
    """(cx,cy,w,h) -> (x1,y1,x2,y2); bbox.py:94-110."""
    x, y, w, h = center[0], center[1], center[2], center[3]
    vals = (x - w * 0.5, y - h * 0.5, x + w * 0.5, y + h * 0.5)
    return Corner(*vals) if isinstance(center, Center) else vals


def cetner2poly(center):
    """(cx,cy,w,h) -> 8-vector TL,TR,BR,BL (sic: the reference's spelling); bbox.py:40-56."""
    x, y, w, h = center[0], center[1], center[2], center[3]
    l, t, r, b = x - w * 0.5, y - h * 0.5, x + w * 0.5, y + h * 0.5
    return np.array([l, t, r, t, r, b, l, b])


def getRotMatrix(cx, cy, rot):
    """3x3 rotation by `rot` about (cx,cy): T(c) R T(-c); bbox.py:58-75."""
    c, s = np.cos(rot), np.sin(rot)
    return np.array([[c, -s, cx - cx * c + cy * s], [s, c, cy - cy * c - cx * s], [0, 0, 1]]).astype(float)


def transformPoly(polygon, m):
    """Apply a 3x3 (affine) matrix to polygon vertices, no perspective divide; bbox.py:77-90."""
    pts = np.ones([polygon.size // 2, 3])
    pts[:, :2] = polygon.reshape(-1, 2)
    return (pts @ m.transpose(1, 0))[:, :2]


def get_axis_aligned_bbox(region):
    """8-vector or (x,y,w,h) -> (cx,cy,w,h) with the VOT area-preserving shrink for polygons; bbox.py:166-200."""
    if region.size == 8:
        xs, ys = region[0::2], region[1::2]
        cx, cy = np.mean(xs), np.mean(ys)
        x1, x2, y1, y2 = min(xs), max(xs), min(ys), max(ys)
        quad = np.linalg.norm(region[0:2] - region[2:4]) * np.linalg.norm(region[2:4] - region[4:6])
        s = np.sqrt(quad / ((x2 - x1) * (y2 - y1)))
        return cx, cy, s * (x2 - x1) + 1, s * (y2 - y1) + 1
    x, y, w, h = region[0], region[1], region[2], region[3]
    return x + w / 2, y + h / 2, w, h


def get_min_max_bbox(region):
    """8-vector or (x,y,w,h) -> (cx,cy,w,h) of the min-max box (centre = vertex mean); bbox.py:204-224."""
    if region.size == 8:
        xs, ys = region[0::2], region[1::2]
        return np.mean(xs), np.mean(ys), max(xs) - min(xs), max(ys) - min(ys)
    x, y, w, h = region[0], region[1], region[2], region[3]
    return x + w / 2, y + h / 2, w, h


def xywh2xyxy(region):
    return region[0], region[1], region[0] + region[2], region[1] + region[3]


def _edge_angle(dx, dy):
    if abs(dx) < 0.5:
        return math.pi / 2
    if abs(dy) < 0.5:
        return 0
    return math.atan(dy / dx)


def get_w_h_from_poly(region):
    """Polygon -> (cx, cy, w, h, theta) of its minimum-area rectangle, w = the longer side; bbox.py:231-269."""
    quad = cv2.boxPoints(cv2.minAreaRect(np.array(region).reshape(-1, 2).astype(np.int32)))
    cx, cy = (quad[0][0] + quad[2][0]) / 2, (quad[0][1] + quad[2][1]) / 2
    e1 = (quad[1][0] - quad[0][0], quad[1][1] - quad[0][1])
    e2 = (quad[2][0] - quad[1][0], quad[2][1] - quad[1][1])
    len1, len2 = np.linalg.norm(e1), np.linalg.norm(e2)
    if len1 > len2:
        theta, w, h = _edge_angle(*e1), len1, len2
    else:
        theta, w, h = _edge_angle(*e2), len2, len1
    if theta > math.pi / 2:
        theta -= math.pi
    elif theta < -math.pi / 2:
        theta = math.pi - theta
    return (cx, cy, w, h, theta)


def get_points_from_xyxy(region):
    """(x,y,w,h) -> TL,TR,BL,BR 8-tuple (the reference's order for this helper); bbox.py:272-285."""
    x, y, w, h = region[0], region[1], region[2], region[3]
    return (x, y, x + w, y, x, y + h, x + w, y + h)


def get_points_from_xywh(region):
    """(x,y,w,h) -> TL,TR,BR,BL 8-tuple; bbox.py:288-301."""
    x, y, w, h = region[0], region[1], region[2], region[3]
    return (x, y, x + w, y, x + w, y + h, x, y + h)


def poly2mask(img_size, polygons):
    """Rasterise polygons into a uint8 mask; bbox.py:306-311."""
    mask = np.zeros(img_size, dtype=np.uint8)
    polys = np.asarray(polygons, np.int32)
    cv2.fillPoly(mask, polys.reshape(polys.shape[0], -1, 2), color=1)
    return mask
