"""Homography tracker -- mirror of hdn/tracker/hdn_tracker_proj_e2e.py (hdnTrackerHomo :21-285), the tracker the
shipped configuration selects (TRACK.TYPE = 'hdnTrackerHomoProje2e').

Per frame (track_new, :141-285), unchanged in meaning:
  0. un-warp the frame by inv(H_total) (cv2.warpPerspective, replicate border)                      :150-155
  1. translation: 255-crop -> model.track_new -> softmax / Hanning / arg-max -> centre shift          :156-185
  2. scale / rotation: re-crop at the moved centre -> model.track_new_lp -> arg-max -> (scale, rot)   :194-217
  3. residual homography: rotate frame by -rot, 127-crop at init_s_z_sm*scale, gray-normalise,
     model.track_proj -> H; conjugate back to frame coordinates; gate on homo_score > 2.5             :223-266
  4. H_total <- H_total @ H_sim @ H_homo / h33; polygon = perspectiveTransform(init points)           :262-285
The three gates (pscore < 0.05, lp score < 0.25, homo_score > 2.5) and the singular-H reset are kept verbatim.

What differs from the reference is only where work happens: stage epilogues run on the device (K6) with one packed
read-back each, the template-side head convolutions are cached, the crop for stage 3 is not bounced
device->host->device, and `get_mask_window` (whose result track_proj never reads, model_builder...py:161) is skipped.
With cfg.CUDA the per-frame pixel work (warpPerspective, the three crops + resizes, the cubic rotation, the gray
normalisation: ~45 ms of single-threaded OpenCV per 1280x720 frame in the reference) also runs on the device, bit-compatible
with OpenCV (hdn_b200/preproc.py): the frame is uploaded once and the host keeps only the 3x3 algebra.
HDN_B200_DEVICE_PREPROC=0 restores the host path.
"""
import os

import cv2
import numpy as np
import torch

from hdn.core.config import cfg
from hdn.tracker.base_tracker import crop_window
from hdn.tracker.hdn_tracker import decode_center, decode_logpolar, hdnTracker
from hdn.utils.point import Point
from hdn.utils.transform import img_rot_around_center, rot_scale_around_center_shift_tran
from homo_estimator.Deep_homography.Oneline_DLTv1.tools.get_img_info import get_search_info, get_template_info, merge_tmp_search

_EYE = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]


def stack_requests(payloads):
    """Batch the per-sequence payloads of one stage (host arrays or device tensors)."""
    if isinstance(payloads[0], torch.Tensor):
        return payloads[0] if len(payloads) == 1 else torch.cat(payloads, 0)
    return payloads[0] if len(payloads) == 1 else np.concatenate(payloads, 0)


def serve(model, kind, batch):
    """Answer one (possibly batched) stage request.  batch: float32 ndarray [B,...] or a device tensor (device-side
    pre-processing).  -> list of B per-item results."""
    if isinstance(batch, torch.Tensor):
        x = batch
    else:
        x = torch.from_numpy(np.ascontiguousarray(batch))
        if cfg.CUDA:
            x = x.pin_memory().cuda(non_blocking=True)
    B = x.shape[0]
    if kind == "template":
        model.template(x)
        return [x[i:i + 1] for i in range(B)]
    if kind == "stage1":
        idx, ps, sc, g = model.track_new_scored(x, cfg.TRACK.WINDOW_INFLUENCE)
    elif kind == "stage2":
        idx, ps, sc, g = model.track_new_lp_scored(x, [0, 0])
    elif kind == "stage3":
        if getattr(model, "_h4p_cache", None) is None or model._h4p_cache.shape[0] != B or model._h4p_cache.device != x.device:
            model._h4p_cache = torch.tensor([[0.0, 0.0, 0.0, 127.0, 127.0, 127.0, 127.0, 0.0]], device=x.device).repeat(B, 1)  # get_img_info.py:93-98
        out = model.track_proj_packed(x, model._h4p_cache)
        return [out[i] for i in range(B)]
    else:
        raise ValueError(kind)
    return [(idx[i], ps[i], sc[i], g[i]) for i in range(B)]


def serve_alone(model, steps):
    """Drive one tracker's step generator with batch-1 network calls; returns the generator's return value."""
    try:
        kind, payload = next(steps)
        while True:
            kind, payload = steps.send(serve(model, kind, payload)[0])
    except StopIteration as done:
        return done.value


class hdnTrackerHomo(hdnTracker):
    def __init__(self, model):
        super().__init__(model)
        self.points = self.generate_points(cfg.POINT.STRIDE, self.score_size)
        self.p = Point(cfg.POINT.STRIDE, cfg.TRAIN.OUTPUT_SIZE, cfg.TRAIN.EXEMPLAR_SIZE // 2)
        self.points_lp = self.generate_points_lp(cfg.POINT.STRIDE_LP, cfg.POINT.STRIDE_LP, cfg.TRAIN.OUTPUT_SIZE_LP)
        self.model.eval()
        self._pre = None
        self.device_preproc = os.environ.get("HDN_B200_DEVICE_PREPROC", "1") != "0"

    def _preproc(self):
        """The device-side pixel pipeline (None = host OpenCV path: CPU configuration or HDN_B200_DEVICE_PREPROC=0)."""
        if not (self.device_preproc and cfg.CUDA and torch.cuda.is_available()):
            return None
        if self._pre is None:
            from hdn_b200.preproc import FramePreproc
            self._pre = FramePreproc(next(self.model.parameters()).device)
        return self._pre

    # ------------------------------------------------------------------ stage 3 network call
    def homo_estimate(self, tmp, search, tmp_mask=None):
        """proj_e2e:43-56: pack (template, search) gray patches and run model.track_proj."""
        info = merge_tmp_search(tmp, search)

        def up(a):
            t = torch.Tensor(np.asarray(a)).float().unsqueeze(0)
            return t.cuda(non_blocking=True) if cfg.CUDA else t

        pair = up(info["org_imgs"])
        data = {"org_imgs": pair, "input_tensors": pair if info["input_tensors"].shape == info["org_imgs"].shape else up(info["input_tensors"]),
                "h4p": up(info["four_points"]), "patch_indices": None}  # identity indices (get_img_info.py:92): gather skipped
        return self.model.track_proj(data, tmp_mask)

    # ------------------------------------------------------------------ first frame
    def init(self, img, bbox, poly, gt_points, first_point):
        """proj_e2e:60-120.  Same arguments as the reference."""
        serve_alone(self.model, self._init_steps(img, bbox, poly, gt_points, first_point))

    def _init_steps(self, img, bbox, poly, gt_points, first_point):
        """Generator form of `init`: yields ('template', z_crop) so several trackers can be templated in one batch."""
        self._start_state(bbox, poly, first_point)
        w_z, h_z, s_z = self._context_size(cfg.TRACK.CONTEXT_AMOUNT)
        _, _, s_z_sm = self._context_size(0)  # no-context window: the homography estimator's crop
        self.channel_average = np.mean(img, axis=(0, 1))
        z_crop, self.z_crop_points = crop_window(img, self.center_pos, cfg.TRACK.EXEMPLAR_SIZE, s_z, self.channel_average, islog=1)
        z_sm, self.z_crop_points_sm = crop_window(img, self.center_pos, cfg.TRACK.EXEMPLAR_SIZE, s_z_sm, self.channel_average, islog=1)
        self.z_crop_sm = torch.from_numpy(z_sm)  # only read on the host (get_template_info): no upload needed
        self.z_crop = yield ("template", z_crop)
        self.init_img = img
        self.init_crop_size = np.array([w_z, h_z])
        self.init_size = self.size
        self.init_s_z, self.init_s_z_sm = s_z, s_z_sm
        self.init_pos = np.array([poly[0], poly[1]])
        self.window_scale_factor = 1.0
        self.lost, self.lost_count, self.last_lost = True, 0, False
        self.init_points = np.array(gt_points).astype(np.float32)
        self.init_homo_tmp, self.print_tmp_img = get_template_info(self.z_crop_sm[:, 0:3, :, :])
        self._tmpl_gray_dev = None  # float32 device copy of init_homo_tmp, made on first use by the device-side path
        self.H_total = np.array(_EYE, dtype=np.float32)
        self.H_total_sim = np.array(_EYE, dtype=np.float32)
        self.uncertain = 0
        self.recover_H = np.identity(3).astype("float")

    def update_template(self):
        img = img_rot_around_center(self.init_img, self.init_pos[0], self.init_pos[1], self.init_img.shape[1], self.init_img.shape[0],
                                    self.lp_shift[1])
        self.z_crop = self.get_subwindow(img, self.init_pos, cfg.TRACK.EXEMPLAR_SIZE, self.init_s_z, self.channel_average, islog=1)
        self.model.template(self.z_crop)

    def update_template_window(self, sc):
        self.model.template(self.get_subwindow(self.init_img, self.init_pos, cfg.TRACK.EXEMPLAR_SIZE, self.init_s_z * sc, self.channel_average,
                                               islog=1))

    def get_points_by_homo(self, uni_points, H):
        return np.vsplit(H @ uni_points, [2])[0].transpose([1, 0])

    # ------------------------------------------------------------------ every other frame
    def track_new(self, fr_idx, img, gt_box=None, gt_poly=None, gt_points=None):
        """proj_e2e:141-285.  Same arguments and result dict as the reference."""
        return serve_alone(self.model, self._track_steps(fr_idx, img))

    def _track_steps(self, fr_idx, img):
        """Generator form of `track_new`: yields ('stage1'|'stage2'|'stage3', host array) requests and receives the stage's
        read-back, so a driver can run many sequences in lock-step with ONE batched network call per stage
        (hdn_b200.batched.LockstepTrackers); `track_new` drives it alone."""
        # 0. bring the frame back into the template's pose
        if np.linalg.det(self.H_total) == 0:
            self.H_total = np.array(_EYE).astype(np.float32)
        pre = self._preproc()
        frame_w, frame_h = img.shape[1], img.shape[0]
        if pre is not None:  # one upload of the frame; every pixel operation below runs on the device, bit-compatible with OpenCV
            pre.upload(img)
            img = pre.warp_perspective(np.linalg.inv(self.H_total))
            crop = lambda src, pos, msz, osz: pre.crop(src, pos, msz, osz, self.channel_average)  # noqa: E731
        else:
            img = cv2.warpPerspective(img, np.linalg.inv(self.H_total), (frame_w, frame_h), borderMode=cv2.BORDER_REPLICATE)
            crop = lambda src, pos, msz, osz: crop_window(src, pos, msz, osz, self.channel_average)[0]  # noqa: E731
        init_points = self.init_points.reshape(-1, 2).astype(np.float32)
        s_z = cur_sz = self.init_s_z
        center_pos = self.init_pos
        ratio = np.round(cfg.TRACK.INSTANCE_SIZE / cfg.TRACK.EXEMPLAR_SIZE)
        scale_z = cfg.TRACK.EXEMPLAR_SIZE / s_z
        s_x = np.floor(s_z * ratio)

        # 1. translation
        idx, ps, sc, g = yield ("stage1", crop(img, center_pos, cfg.TRACK.INSTANCE_SIZE, s_x))
        best_idx, pbest, best_score, pred_c = int(idx), ps, sc, decode_center(self.points, int(idx), g)
        stop_update = pbest < 0.05
        center = [0, 0] if stop_update else pred_c / scale_z
        cx, cy = center[0] + center_pos[0], center[1] + center_pos[1]
        delta_cx, delta_cy = center[0], center[1]
        self.center_pos = np.array([cx, cy])

        # 2. scale / rotation in log-polar coordinates
        idx, _, lp_score, g = yield ("stage2", crop(img, self.center_pos, cfg.TRACK.INSTANCE_SIZE, s_x))
        sim_lp = decode_logpolar(self.points_lp, int(idx), g)
        if stop_update or lp_score < 0.25:
            sim_lp = [1, 1, 0, 0]
        scale_delta = sim_lp[0] * cur_sz / self.init_s_z
        rot_delta = sim_lp[2]
        H_sim = rot_scale_around_center_shift_tran(cx, cy, rot_delta, scale_delta, delta_cx, delta_cy)
        self.rot += rot_delta
        self.scale *= scale_delta

        # 3. residual homography on the de-rotated, re-scaled no-context crop
        crop_w = self.z_crop_points_sm[2] - self.z_crop_points_sm[0] + 1
        crop_h = self.z_crop_points_sm[3] - self.z_crop_points_sm[1] + 1
        if pre is not None:
            rot_img = pre.rotate(img, cx, cy, -rot_delta)
            search_gray = pre.crop(rot_img, self.center_pos, cfg.TRACK.EXEMPLAR_SIZE, self.init_s_z_sm * scale_delta, self.channel_average, gray=True)
            if self._tmpl_gray_dev is None or self._tmpl_gray_dev.device != search_gray.device:
                self._tmpl_gray_dev = torch.from_numpy(np.asarray(self.init_homo_tmp).astype(np.float32)).to(search_gray.device)
            pair = torch.cat((self._tmpl_gray_dev, search_gray), 0)[None]
        else:
            rot_img = img_rot_around_center(img, cx, cy, frame_w, frame_h, -rot_delta)
            x_homo, _ = crop_window(rot_img, self.center_pos, cfg.TRACK.EXEMPLAR_SIZE, self.init_s_z_sm * scale_delta, self.channel_average)
            search_gray, _ = get_search_info(torch.from_numpy(x_homo)[:, 0:3, :, :])
            pair = np.concatenate([self.init_homo_tmp, search_gray], 0).astype(np.float32)[np.newaxis]
        both = yield ("stage3", pair)  # H (9) + scores
        homo_score = both[9]
        H_hm = np.linalg.inv(both[:9].reshape(3, 3))
        H_hm = (1.0 / H_hm.item(8)) * H_hm
        H_comp = np.identity(3) @ H_hm
        to_square = np.array([[127 / crop_w, 0, 0], [0, 127 / crop_h, 0], [0, 0, 1]]).astype(np.float32)
        H_comp = np.linalg.inv(to_square) @ H_comp @ to_square
        to_crop = np.array([[1, 0, -self.z_crop_points_sm[0]], [0, 1, -self.z_crop_points_sm[1]], [0, 0, 1]]).astype(np.float32)
        H_homo = np.linalg.inv(to_crop) @ H_comp @ to_crop

        # 4. compose; a poor photometric score drops the residual
        H = self.H_total @ H_sim if homo_score > 2.5 else self.H_total @ H_sim @ H_homo
        self.H_total = (1.0 / H.item(8)) * H

        pred_points = cv2.perspectiveTransform(np.expand_dims(init_points, 0), self.H_total)[0]
        lo, hi = np.min(pred_points, 0), np.max(pred_points, 0)
        aligned = [lo[0], lo[1], hi[0] - lo[0], hi[1] - lo[1]]
        self.align_size = [aligned[2], aligned[3]]
        return {"bbox_aligned": aligned, "best_score": best_score, "polygon": pred_points, "points": pred_points, "bbox": aligned}
