"""Tracker base classes -- mirror of hdn/tracker/base_tracker.py (BaseTracker :18-36, SiameseTracker :39-213).

`get_subwindow` / `get_subwindow_for_homo` (:61-136 / :138-213) cut a square window of side `original_sz` centred at
`pos` out of a BGR frame, pad what falls outside with the channel means, resize to `model_sz` with cv2.resize and
(islog) append the cv2.logPolar image as 3 more channels.  Host-side OpenCV like the reference (SURVEY 8(f)-1 lists
the device-side version as a later row); the upload goes through pinned memory.
"""
import cv2
import numpy as np
import torch

from hdn.core.config import cfg
from hdn.models.logpolar import getLinearPolarImg, getPolarImg


class BaseTracker(object):
    def init(self, img, bbox):
        raise NotImplementedError

    def track(self, img):
        raise NotImplementedError


def _fill(region, colour):
    """region[..., c] = colour[c]; per channel, because NumPy broadcasts a length-3 inner axis ~10x slower than a scalar fill."""
    if region.size:
        for ch in range(region.shape[2]):
            region[:, :, ch] = colour[ch]


def crop_window(im, pos, model_sz, original_sz, avg_chans, islog=False):
    """-> (float32 ndarray [1,C,model_sz,model_sz], (xmin, ymin, xmax+1, ymax+1) in padded-frame coordinates)."""
    if isinstance(pos, float):
        pos = [pos, pos]
    if im.ndim == 2:
        im = im.reshape(im.shape[0], im.shape[1], 1)
    rows, cols, chans = im.shape
    half = (original_sz - 1) / 2
    x0 = np.floor(pos[0] - half + 0.5)
    y0 = np.floor(pos[1] - half + 0.5)
    x1, y1 = x0 + original_sz - 1, y0 + original_sz - 1
    left, top = int(max(0.0, -x0)), int(max(0.0, -y0))
    right, bottom = int(max(0.0, x1 - cols + 1)), int(max(0.0, y1 - rows + 1))
    x0, x1, y0, y1 = x0 + left, x1 + left, y0 + top, y1 + top
    ya, yb, xa, xb = int(y0), int(y1 + 1), int(x0), int(x1 + 1)  # window in the coordinates of the padded frame
    if left or top or right or bottom:
        # The reference materialises the whole padded frame (base_tracker.py:99-112) and slices it; every padded pixel is the
        # channel mean (cast to uint8 on assignment) and every other pixel is the frame, so only the window is built here.
        patch = np.empty((yb - ya, xb - xa, chans), np.uint8)
        fill = np.asarray(avg_chans).astype(np.uint8)  # the same float -> uint8 cast the reference's slice assignment performs
        ia, ib = max(ya, top), min(yb, top + rows)
        ja, jb = max(xa, left), min(xb, left + cols)
        if ib > ia and jb > ja:
            patch[ia - ya:ib - ya, ja - xa:jb - xa, :] = im[ia - top:ib - top, ja - left:jb - left, :]
            _fill(patch[:ia - ya], fill)
            _fill(patch[ib - ya:], fill)
            _fill(patch[ia - ya:ib - ya, :ja - xa], fill)
            _fill(patch[ia - ya:ib - ya, jb - xa:], fill)
        else:
            _fill(patch, fill)
    else:
        patch = im[ya:yb, xa:xb, :]
    if not np.array_equal(model_sz, original_sz):
        patch = cv2.resize(patch, (model_sz, model_sz))
    if islog:
        log_img = getPolarImg(patch) if islog == 1 else getLinearPolarImg(patch)
        patch = np.concatenate((patch.reshape(patch.shape[0], patch.shape[1], -1), log_img.reshape(log_img.shape[0], log_img.shape[1], -1)), 2)
    if patch.ndim == 2:
        patch = patch.reshape(patch.shape[0], patch.shape[1], 1)
    return np.ascontiguousarray(patch.transpose(2, 0, 1)[np.newaxis].astype(np.float32)), (x0, y0, x1 + 1, y1 + 1)


def to_model_tensor(array):
    t = torch.from_numpy(array)
    if cfg.CUDA:
        t = t.pin_memory().cuda(non_blocking=True) if torch.cuda.is_available() else t.cuda()
    return t


class SiameseTracker(BaseTracker):
    def _convert_delta(self, delta):
        return delta.permute(1, 2, 3, 0).contiguous().view(4, -1).detach().cpu().numpy()

    def _convert_c(self, delta, point):
        """base_tracker.py:54-59: centre prediction = anchor point - 8 * predicted offset, per score-map cell."""
        delta = delta.permute(1, 2, 3, 0).contiguous().view(2, -1).detach().cpu().numpy()
        delta[0, :] = point[:, 0] - delta[0, :] * 8
        delta[1, :] = point[:, 1] - delta[1, :] * 8
        return delta

    def get_subwindow(self, im, pos, model_sz, original_sz, avg_chans, islog=False):
        return to_model_tensor(crop_window(im, pos, model_sz, original_sz, avg_chans, islog)[0])

    def get_subwindow_for_homo(self, im, pos, model_sz, original_sz, avg_chans, islog=False):
        patch, box = crop_window(im, pos, model_sz, original_sz, avg_chans, islog)
        return to_model_tensor(patch), box
