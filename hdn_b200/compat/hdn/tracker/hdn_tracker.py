"""Similarity tracker -- mirror of hdn/tracker/hdn_tracker.py (hdnTracker :18-301).

It provides the score / offset decoding the homography tracker inherits (`_convert_score` :82-89,
`_convert_logpolar_simi` :51-67, `generate_points*` :32-49) and the similarity-only `init` / `track_new`
(:110-301; cfg.TRACK.TYPE == 'hdnTracker').

Device/host split: the reference copies the whole cls / loc maps to the host every stage and does softmax, window
blend and arg-max in NumPy.  Here stage epilogues run on the device (K6, hdn_score_argmax_f32) and ONE packed copy
brings back (idx, pscore, score, loc[:, idx]); the decode below then applies the reference's NumPy expressions --
same dtypes, same order -- to that single column, so every value downstream is what the reference computes.
"""
import math

import cv2
import numpy as np

from hdn.core.config import cfg
from hdn.tracker.base_tracker import SiameseTracker
from hdn.utils.bbox import cetner2poly, getRotMatrix, transformPoly
from hdn.utils.point import Point, generate_points, generate_points_lp
from hdn.utils.transform import img_rot_around_center


def decode_center(points, idx, loc_col):
    """Column `idx` of base_tracker.py:54-59 `_convert_c`: point - 8 * loc, float32."""
    col = np.asarray(loc_col, np.float32).reshape(2, 1).copy()
    col[0, :] = points[idx:idx + 1, 0] - col[0, :] * 8
    col[1, :] = points[idx:idx + 1, 1] - col[1, :] * 8
    return col[:, 0]


def decode_logpolar(points_lp, idx, loc_col):
    """Column `idx` of hdn_tracker.py:51-67 `_convert_logpolar_simi` -> float32 [scale, scale, rot, .]."""
    d = np.asarray(loc_col, np.float32).reshape(4, 1).copy()
    pt = points_lp[idx:idx + 1]
    d[2, :] = pt[:, 1] - d[2, :] * cfg.POINT.STRIDE_LP
    d[3, :] = pt[:, 1] + d[3, :] * cfg.POINT.STRIDE_LP
    d[0, :] = pt[:, 0] - d[0, :] * cfg.POINT.STRIDE_LP
    d[1, :] = pt[:, 0] + d[1, :] * cfg.POINT.STRIDE_LP
    rotation = d[2, :] * (2 * np.pi / cfg.TRAIN.EXEMPLAR_SIZE)
    mag = np.log(cfg.TRAIN.EXEMPLAR_SIZE / 2) / cfg.TRAIN.EXEMPLAR_SIZE
    d[0, :] = np.exp(d[0, :] * mag)
    d[1, :] = d[0, :]
    d[2, :] = rotation
    return d[:, 0]


class hdnTracker(SiameseTracker):
    def __init__(self, model):
        super().__init__()
        self.score_size = (cfg.TRACK.INSTANCE_SIZE - cfg.TRACK.EXEMPLAR_SIZE) // cfg.POINT.STRIDE + 1 + cfg.TRACK.BASE_SIZE
        hanning = np.hanning(self.score_size)
        self.window = np.outer(hanning, hanning).flatten()
        self.cls_out_channels = cfg.BAN.KWARGS.cls_out_channels
        self.points = generate_points(cfg.POINT.STRIDE, self.score_size)
        self.p = Point(cfg.POINT.STRIDE, cfg.TRAIN.OUTPUT_SIZE, cfg.TRAIN.EXEMPLAR_SIZE // 2)
        self.points_lp = generate_points_lp(cfg.POINT.STRIDE_LP, cfg.POINT.STRIDE_LP, cfg.TRAIN.OUTPUT_SIZE_LP)
        self.model = model

    # grids as methods too (the homography tracker calls them through self, hdn_tracker.py:32-49)
    def generate_points(self, stride, size):
        return generate_points(stride, size)

    def generate_points_lp(self, stride_w, stride_h, size):
        return generate_points_lp(stride_w, stride_h, size)

    # ---- whole-map decoders (kept for API parity; the tracking loop uses the single-column forms above) -------
    def _convert_logpolar_simi_in_lp(self, delta, point, peak_idx, idx=0):
        d = delta.permute(1, 2, 3, 0).contiguous().view(4, -1).detach().cpu().numpy()
        d[2, :] = point[:, 1] - d[2, :] * cfg.POINT.STRIDE_LP
        d[3, :] = point[:, 1] + d[3, :] * cfg.POINT.STRIDE_LP
        d[0, :] = point[:, 0] - d[0, :] * cfg.POINT.STRIDE_LP
        d[1, :] = point[:, 0] + d[1, :] * cfg.POINT.STRIDE_LP
        return d

    def _convert_logpolar_simi(self, delta, point, peak_idx, idx=0):
        d = self._convert_logpolar_simi_in_lp(delta, point, peak_idx, idx)
        rotation = d[2, :] * (2 * np.pi / cfg.TRAIN.EXEMPLAR_SIZE)
        mag = np.log(cfg.TRAIN.EXEMPLAR_SIZE / 2) / cfg.TRAIN.EXEMPLAR_SIZE
        d[0, :] = np.exp(d[0, :] * mag)
        d[1, :] = d[0, :]
        d[2, :] = rotation
        return d

    def _convert_score(self, score):
        if self.cls_out_channels == 1:
            return score.permute(1, 2, 3, 0).contiguous().view(-1).sigmoid().detach().cpu().numpy()
        s = score.permute(1, 2, 3, 0).contiguous().view(self.cls_out_channels, -1).permute(1, 0)
        return s.softmax(1).detach()[:, 1].cpu().numpy()

    # ---- device epilogues ----------------------------------------------------------------------------------------
    def _stage1(self, x_crop):
        """-> (best_idx, pscore[best], score[best] as np.float32, pred_c[:, best] float32)."""
        if hasattr(self.model, "track_new_scored") and self.cls_out_channels == 2:
            idx, ps, sc, g = self.model.track_new_scored(x_crop, cfg.TRACK.WINDOW_INFLUENCE)
            i = int(idx[0])
            return i, ps[0], sc[0], decode_center(self.points, i, g[0])
        out = self.model.track_new(x_crop)
        score = self._convert_score(out["cls"])
        pred_c = self._convert_c(out["loc_c"], self.points)
        pscore = score * (1 - cfg.TRACK.WINDOW_INFLUENCE) + self.window * cfg.TRACK.WINDOW_INFLUENCE
        i = int(np.argmax(pscore))
        return i, pscore[i], score[i], pred_c[:, i]

    def _stage2(self, x_crop, fr_idx=0):
        """-> (best_idx_lp, score_lp[best], decoded [scale, scale, rot, .] float32)."""
        if hasattr(self.model, "track_new_lp_scored") and self.cls_out_channels == 2:
            idx, ps, sc, g = self.model.track_new_lp_scored(x_crop, [0, 0])
            i = int(idx[0])
            return i, sc[0], decode_logpolar(self.points_lp, i, g[0])
        out = self.model.track_new_lp(x_crop, [0, 0])
        score_lp = self._convert_score(out["cls_lp"])
        i = int(np.argmax(score_lp))
        return i, score_lp[i], self._convert_logpolar_simi(out["loc_lp"], self.points_lp, i, fr_idx)[:, i]

    # ---- misc helpers ------------------------------------------------------------------------------------------------
    def mask_img(self, img, points):
        mask = np.zeros([img.shape[0], img.shape[1]])
        cv2.drawContours(mask, [points.astype(np.int32)], 0, (1), -1)
        img[np.where(mask <= 0)] = 0
        return img

    def get_window_scale_coef(self, region):
        region = region.reshape(8, -1)
        xs, ys = region[0::2], region[1::2]
        quad = np.linalg.norm(region[0:2] - region[2:4]) * np.linalg.norm(region[2:4] - region[4:6])
        return np.sqrt(quad / ((max(xs) - min(xs)) * (max(ys) - min(ys))))

    def _start_state(self, bbox, poly, first_point):
        """State shared by both trackers' init (hdn_tracker.py:118-133 / proj_e2e:69-83)."""
        self.center_pos = np.array([poly[0], poly[1]])
        self.init_rot = self.rot = poly[4]
        polygon = transformPoly(cetner2poly(poly[:4]), getRotMatrix(poly[0], poly[1], poly[4]))
        d2 = (polygon - first_point) ** 2
        self.poly_shift_l = np.argmin(d2[:, 0] + d2[:, 1])
        self.scale = 1
        self.lp_shift = [0, 0]
        self.v = 0
        self.size = np.array([poly[2], poly[3]])
        self.align_size = np.array([bbox[2], bbox[3]])
        return polygon

    def _context_size(self, amount):
        w = self.size[0] + amount * np.sum(self.size)
        h = self.size[1] + amount * np.sum(self.size)
        return w, h, np.floor(np.sqrt(w * h))

    # ---- similarity-only tracking (cfg.TRACK.TYPE == 'hdnTracker') ---------------------------------------------------
    def init(self, img, bbox, poly, first_point):
        polygon = self._start_state(bbox, poly, first_point)
        self.scale_coeff = self.get_window_scale_coef(polygon)
        w_z, h_z, s_z = self._context_size(cfg.TRACK.CONTEXT_AMOUNT)
        self.channel_average = np.mean(img, axis=(0, 1))
        self.model.template(self.get_subwindow(img, self.center_pos, cfg.TRACK.EXEMPLAR_SIZE, s_z, self.channel_average, islog=1))
        self.init_img = img
        self.init_crop_size = np.array([w_z, h_z])
        self.init_size = self.size
        self.init_s_z = s_z
        self.init_pos = np.array([poly[0], poly[1]])
        self.window_scale_factor = 1.0
        self.lost, self.lost_count, self.last_lost = True, 0, False

    def update_template(self):
        img = img_rot_around_center(self.init_img, self.init_pos[0], self.init_pos[1], self.init_img.shape[1], self.init_img.shape[0],
                                    self.lp_shift[1])
        self.model.template(self.get_subwindow(img, self.init_pos, cfg.TRACK.EXEMPLAR_SIZE, self.init_s_z, self.channel_average, islog=1))

    def update_template_window(self, sc):
        self.model.template(self.get_subwindow(self.init_img, self.init_pos, cfg.TRACK.EXEMPLAR_SIZE, self.init_s_z * sc,
                                               self.channel_average, islog=1))

    def track_new(self, fr_idx, img, gt_box=None, gt_poly=None):
        """hdn_tracker.py:173-301: translation, then scale/rotation; the template is re-cropped (rotated) every frame."""
        w_z, h_z, s_z = self._context_size(cfg.TRACK.CONTEXT_AMOUNT)
        ratio = np.round(cfg.TRACK.INSTANCE_SIZE / cfg.TRACK.EXEMPLAR_SIZE)
        scale_z = cfg.TRACK.EXEMPLAR_SIZE / s_z
        s_x = np.floor(s_z * ratio * 1)  # the reference forces window_scale_factor to 1 before use (:186)
        self.window_scale_factor = s_x / (s_z * ratio)

        best_idx, pbest, best_score, pred_c = self._stage1(self.get_subwindow(img, self.center_pos, cfg.TRACK.INSTANCE_SIZE, s_x,
                                                                              self.channel_average))
        stop_update = pbest < 0.05
        center = [0, 0] if stop_update else pred_c / scale_z * self.window_scale_factor
        next_factor = 1
        if pbest < cfg.TRACK.SCALE_SCORE_THRESH:
            next_factor = 1.5
            if self.lost_count == 0:
                self.last_lost = True
            self.lost_count += 1
            if not self.last_lost and self.lost_count < 5:
                self.lost_count, self.last_lost = 0, False
        speed = math.sqrt(center[0] * center[0] + center[1] * center[1])
        self.v = speed if fr_idx == 1 else (self.v + speed) / 2
        cx, cy = center[0] + self.center_pos[0], center[1] + self.center_pos[1]
        self.center_pos = np.array([cx, cy])

        _, lp_score, sim_lp = self._stage2(self.get_subwindow(img, self.center_pos, cfg.TRACK.INSTANCE_SIZE, s_x, self.channel_average), fr_idx)
        if stop_update or lp_score < 0.25:
            sim_lp = [1, 1, 0, 0]
        width = self.size[0] * sim_lp[0] * self.window_scale_factor
        height = self.size[1] * sim_lp[1] * self.window_scale_factor
        width = max(10 * self.init_size[0] / self.init_size[1], min(width, img.shape[:2][1]))
        height = max(10, min(height, img.shape[:2][0]))
        self.size = np.array([width, height])
        self.lp_shift[1] += sim_lp[2]
        self.rot += sim_lp[2]
        self.scale = width / self.init_size[0]
        if self.rot >= 2 * math.pi:
            self.rot -= 2 * math.pi
            self.lp_shift[1] -= 2 * math.pi
        elif self.rot < -2 * math.pi:
            self.rot += 2 * math.pi
            self.lp_shift[1] += 2 * math.pi
        polygon = transformPoly(cetner2poly([cx, cy, width, height]), getRotMatrix(cx, cy, self.rot))
        polygon = np.roll(polygon, 4 - self.poly_shift_l, 0)
        lo, hi = np.min(polygon, 0), np.max(polygon, 0)
        aligned = [lo[0], lo[1], hi[0] - lo[0], hi[1] - lo[1]]
        self.align_size = [aligned[2], aligned[3]]
        self.update_template()
        self.window_scale_factor = next_factor
        return {"bbox": [cx - width / 2, cy - height / 2, width, height], "bbox_aligned": aligned, "best_score": best_score, "rot": self.rot,
                "polygon": polygon}
