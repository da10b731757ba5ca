"""Device-side frame pre-processing for the tracker.

`FramePreproc` keeps one uint8 frame on the device and answers the tracker's pixel requests there, bit-compatible with the
OpenCV calls of the reference (hdn_b200/csrc/preproc.cu; restated and pinned against cv2 in oracle/cv_port.py):

    upload(img)                      one pinned H2D copy of the BGR frame (2.8 MB at 1280x720) instead of three crops
    warp_perspective(M)              cv2.warpPerspective(img, M, (w, h), borderMode=BORDER_REPLICATE)       proj_e2e:154
    crop(pos, model_sz, original_sz) SiameseTracker.get_subwindow -> float32 [1,3,S,S] on the device        base_tracker.py:61-136
    rotate(cx, cy, rot)              img_rot_around_center (cubic cv2.warpAffine)                            transform.py:69-100
    crop_gray(...)                   get_subwindow + get_search_info's gray normalisation -> [1,S,S]          get_img_info.py:42-70

The reference spends ~45 ms per 1280x720 frame in these host calls [SURVEY 3.2]; here the host only does the 3x3 algebra.
"""
import ctypes
import math

import cv2
import numpy as np
import torch

from . import _lib, ops

_MEAN = (ctypes.c_double * 3)(118.93, 113.97, 102.60)  # get_img_info.py:50-51
_STD = (ctypes.c_double * 3)(69.85, 68.81, 72.45)
_CUBIC = {}


def cubic_table(device):
    """The 32x32x16 int16 bicubic weight table (OpenCV's BicubicTab_i) on `device`, built once per device."""
    key = str(device)
    if key not in _CUBIC:
        host = np.zeros(32 * 32 * 16, np.int16)
        _lib.check(_lib.lib().hdn_cubic_table_host(host.ctypes.data_as(ctypes.c_void_p)), "hdn_cubic_table_host")
        _CUBIC[key] = torch.from_numpy(host).to(device)
    return _CUBIC[key]


def crop_geometry(pos, original_sz):
    """Window of get_subwindow in frame coordinates (base_tracker.py:76-92): top-left pixel and side in pixels, and the
    (xmin, ymin, xmax+1, ymax+1) box in padded-frame coordinates that get_subwindow_for_homo also returns."""
    half = (original_sz - 1) / 2
    x0 = np.floor(pos[0] - half + 0.5)
    y0 = np.floor(pos[1] - half + 0.5)
    n = int(x0 + original_sz - 1 + 1) - int(x0)
    return int(x0), int(y0), n


class FramePreproc:
    def __init__(self, device="cuda"):
        _lib.lib()
        self.device = torch.device(device)
        self.H = self.W = 0
        self.frame = self.warped = self.rotated = None
        self.pinned = None

    def _ensure(self, H, W):
        if (H, W) != (self.H, self.W):
            self.H, self.W = H, W
            mk = lambda: torch.empty((H, W, 3), dtype=torch.uint8, device=self.device)  # noqa: E731
            self.frame, self.warped, self.rotated = mk(), mk(), mk()
            self.pinned = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()

    def upload(self, img):
        """BGR uint8 [H, W, 3] host array -> device (through a pinned staging buffer)."""
        if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
            raise ValueError("FramePreproc.upload expects a BGR uint8 [H,W,3] frame")
        self._ensure(img.shape[0], img.shape[1])
        self.pinned.numpy()[...] = img
        self.frame.copy_(self.pinned, non_blocking=True)
        return self.frame

    def warp_perspective(self, M, src=None):
        """-> device frame = cv2.warpPerspective(src, M, (W, H), borderMode=BORDER_REPLICATE)."""
        src = self.frame if src is None else src
        minv = np.ascontiguousarray(cv2.invert(np.asarray(M, np.float64))[1], np.float64)  # what cv::warpPerspective does to M first
        with ops._on_device(src):
            st = _lib.lib().hdn_warp_perspective_u8(ops._ptr(src), ops._ptr(self.warped), self.H, self.W,
                                                    minv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ops._stream())
        _lib.check(st, "hdn_warp_perspective_u8")
        return self.warped

    def rotate(self, src, cx, cy, rot):
        """-> device frame = img_rot_around_center(src, cx, cy, W, H, rot) (transform.py:69-100)."""
        cc, ss = math.cos(rot), math.sin(rot)
        M = np.array([[cc, -ss, cx - cx * cc + cy * ss], [ss, cc, cy - cy * cc - cx * ss]], np.float64)
        with ops._on_device(src):
            st = _lib.lib().hdn_warp_affine_cubic_u8(ops._ptr(src), ops._ptr(self.rotated), self.H, self.W,
                                                     M.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ops._ptr(cubic_table(self.device)), ops._stream())
        _lib.check(st, "hdn_warp_affine_cubic_u8")
        return self.rotated

    def crop(self, src, pos, model_sz, original_sz, avg_chans, gray=False, out=None):
        """get_subwindow(src, pos, model_sz, original_sz, avg_chans) on the device -> float32 [1,3,S,S] (gray: [1,S,S], normalised)."""
        x0, y0, n = crop_geometry(pos, original_sz)
        S = int(model_sz)
        if out is None:
            out = torch.empty((1, S, S) if gray else (1, 3, S, S), dtype=torch.float32, device=self.device)
        fill = (ctypes.c_uint8 * 3)(*[int(v) for v in np.asarray(avg_chans).astype(np.uint8)])
        with ops._on_device(src):
            st = _lib.lib().hdn_crop_resize_u8(ops._ptr(src), self.H, self.W, x0, y0, n, fill, S, int(gray), _MEAN, _STD, ops._ptr(out), ops._stream())
        _lib.check(st, "hdn_crop_resize_u8")
        return out
