"""Host-side geometry used once per frame by the tracker -- mirror of the inference-side functions of
hdn/utils/transform.py: img_rot_around_center (:69-100), get_mask_window (:203-243),
rot_scale_around_center_shift_tran (:250-298), shift_tran (:300-303), homo_add_shift (:244-247).
The training-only combine_affine_* helpers (:413-523) are not mirrored.  float64 NumPy / OpenCV like the reference.
"""
import math

import cv2
import numpy as np


def _rotation_about(cx, cy, rot, tx=0.0, ty=0.0):
    """T(c) R(rot) T(-c) (+ translation) as a float64 3x3."""
    c, s = math.cos(rot), math.sin(rot)
    return np.array([[c, -s, cx - cx * c + cy * s + tx], [s, c, cy - cy * c - cx * s + ty], [0, 0, 1]]).astype(float)


def img_rot_around_center(img, cx, cy, w, h, rot):
    """Rotate the whole frame about (cx,cy): cv2.warpAffine, flags=2 (bicubic), replicate border."""
    return cv2.warpAffine(img, _rotation_about(cx, cy, rot)[:2], (w, h), flags=2, borderMode=cv2.BORDER_REPLICATE)


def get_mask_window(w, h, rot, sx, sy, out_size_w, out_size_h):
    """A floor(w) x floor(h) block of ones, rotated about its centre and placed with its centre at (sx,sy)."""
    w, h = math.floor(w), math.floor(h)
    ones = np.ones([h, w]).astype("float32")
    m = _rotation_about(w / 2, h / 2, rot, sx - w / 2, sy - h / 2)
    return cv2.warpAffine(ones, m[:2], (out_size_w, out_size_h))


def shift_tran(sx, sy):
    return np.array([[1, 0, sx], [0, 1, sy], [0, 0, 1]]).astype(float)


def homo_add_shift(H, shift):
    H[0][2] += shift[0]
    H[1][2] += shift[1]
    return H


def rot_scale_around_center_shift_tran(cx, cy, rot, scale, sx, sy):
    """Similarity as a 3x3: shift by (sx,sy), then scale about (cx,cy) (skipped when scale is 0 or exactly 1), then
    rotate about (cx,cy) (skipped when rot is exactly 0): R S T."""
    tran = shift_tran(sx, sy)
    if abs(scale) > 0 and scale != 1:
        tran = np.array([[scale, 0, cx * (1 - scale)], [0, scale, cy * (1 - scale)], [0, 0, 1]]).astype(float) @ tran
    if abs(rot) > 0:
        tran = _rotation_about(cx, cy, rot) @ tran
    return tran


def img_proj_trans(img, trans, w, h):
    return cv2.warpPerspective(img, trans, (w, h), borderMode=cv2.BORDER_REPLICATE)
