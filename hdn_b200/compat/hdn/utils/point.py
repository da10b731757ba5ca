"""Anchor-point grids -- mirror of hdn/utils/point.py:13-29 (`generate_points`, `generate_points_lp`) and the `Point`
class (:82-99) the tracker constructs.  Points are the search-crop coordinates of the score-map cells relative to the
crop centre: -(size//2)*stride + stride*i  (e.g. -96 .. 96 for stride 8, size 25)."""
import numpy as np


def _grid(stride_x, stride_y, size):
    xs = (-(size // 2) * stride_x + stride_x * np.arange(0, size)).astype(np.float32)
    ys = (-(size // 2) * stride_y + stride_y * np.arange(0, size)).astype(np.float32)
    gx, gy = np.meshgrid(xs, ys)
    pts = np.zeros((size * size, 2), dtype=np.float32)
    pts[:, 0], pts[:, 1] = gx.flatten(), gy.flatten()
    return pts


def generate_points(stride, size):
    return _grid(stride, stride, size)


def generate_points_lp(stride_w, stride_h, size):
    return _grid(stride_w, stride_h, size)


class Point:
    """[2, size, size] grid of absolute crop coordinates: image_center + stride * (i - size//2)."""

    def __init__(self, stride, size, image_center):
        self.stride, self.size, self.image_center = stride, size, image_center
        origin = image_center - size // 2 * stride
        axis = np.array([origin + stride * i for i in np.arange(0, size)])
        gx, gy = np.meshgrid(axis, axis)
        self.points = np.zeros((2, size, size), dtype=np.float32)
        self.points[0], self.points[1] = gx.astype(np.float32), gy.astype(np.float32)
