#!/usr/bin/env python
"""Model- and tracker-level goldens from the REFERENCE ITSELF (CPU, this container).

TEST INFRASTRUCTURE ONLY.   python oracle/gen_golden_model.py [--calibrate] [what ...]
Writes tests/golden/model_*.npz, tests/golden/tracker_*.npz and tests/golden/state_dict_keys.json.

The reference's ModelBuilder / hdnTrackerHomo are imported unmodified from /root/reference (oracle/ref_import.py),
filled with the deterministic weight fixture of hdn_b200/synthetic.py (no checkpoint ships with the reference),
and run on seeded inputs from the same file.  The tests rebuild the same weights and inputs on the GPU box from the
same seeds, so only the (small) outputs are stored.

Reference call sites exercised:
  model  ModelBuilder.template / track_new / track_new_lp / track_proj   model_builder_e2e_unconstrained_v2.py:87-217
  tracker hdnTrackerHomo.init / track_new                                hdn_tracker_proj_e2e.py:60-285
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

ref_import.install()
ref_import.force_homo_backbone_offline()
import torch  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
YAML = os.path.join(ref_import.REF_ROOT, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml")


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# loaded BY PATH: in this process `hdn` is the reference, and the repo root is not importable (oracle/ref_import.py)
synth = load_by_path("hdn_b200_synthetic", os.path.join(REPO, "hdn_b200", "synthetic.py"))


class fixture:  # same names the rest of this script uses
    SCALES = synth.SCALES
    fill = staticmethod(synth.fill_weights)
t = torch.from_numpy


def build_reference_model(instance=255, exemplar=127, scales=None, variant=None):
    from hdn.core.config import cfg
    cfg.merge_from_file(YAML)
    cfg.TRACK.INSTANCE_SIZE, cfg.TRACK.EXEMPLAR_SIZE = instance, exemplar
    cfg.CUDA = False
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    torch.manual_seed(0)
    model = ModelBuilder()
    fixture.fill(model, scales, variant)
    return model, cfg


FEATURE_STRIDE = {255: 3, 512: 13}  # keeps the committed fixtures small (zf / zf_lp / x_lp samples, ~0.4 MB per golden)


def homo_inputs(seed, B=1):
    rng = np.random.default_rng(seed)
    pair = rng.standard_normal((B, 2, 127, 127)).astype(np.float32)
    h4p = np.tile(np.asarray([0, 0, 0, 127, 127, 127, 127, 0], np.float32), (B, 1))
    return pair, h4p


def run_model(model, instance, exemplar, seed):
    z = synth.crop_tensor(seed, (1, 6, exemplar, exemplar))
    x = synth.crop_tensor(seed + 1, (1, 3, instance, instance))
    pair, h4p = homo_inputs(seed + 2)
    out = {}
    with torch.no_grad():
        model.template(t(z))
        stride = FEATURE_STRIDE[instance]  # element-wise samples of the neck features: every stride-th value of the flattened map
        for i, f in enumerate(model.zf):
            out["zf%d_stats" % i] = np.asarray([f.mean().item(), f.std().item(), f.abs().max().item()], np.float32)
            out["zf%d_sample" % i] = f.reshape(-1)[::stride].numpy().copy()
        for i, f in enumerate(model.zf_lp):
            out["zf_lp%d_stats" % i] = np.asarray([f.mean().item(), f.std().item(), f.abs().max().item()], np.float32)
            out["zf_lp%d_sample" % i] = f.reshape(-1)[::stride].numpy().copy()
        out["feature_stride"] = np.int32(stride)
        r = model.track_new(t(x))
        out["cls"], out["loc_c"] = r["cls"].numpy(), r["loc_c"].numpy()
        r = model.track_new_lp(t(x), [0, 0])
        out["cls_lp"], out["loc_lp"] = r["cls_lp"].numpy(), r["loc_lp"].numpy()
        out["x_lp_stats"] = np.asarray([r["x_lp"].mean().item(), r["x_lp"].std().item()], np.float32)
        out["x_lp_sample"] = r["x_lp"].reshape(-1)[::stride].numpy().copy()
        data = {"org_imgs": t(pair), "input_tensors": t(pair), "h4p": t(h4p),
                "patch_indices": t(np.tile(np.arange(127 * 127, dtype=np.float32), (1, 1)))}
        H, s_homo, s_simi = model.track_proj(data, None)
        out["H"], out["homo_score"], out["simi_score"] = H.numpy(), np.float32(s_homo.item()), np.float32(s_simi.item())
        # the fc output itself (offsets), for a tighter check than H
        p1 = model.hm_net.ShareFeature(t(pair)[:, :1])
        p2 = model.hm_net.ShareFeature(t(pair)[:, 1:])
        y = model.hm_net.backbone(torch.cat((p1, p2), 1))
        out["offsets"] = model.hm_net.fc(model.hm_net.avgpool(y).flatten(1)).numpy()
    return out


def calibrate():
    model, _ = build_reference_model(scales={k: 1.0 for k in fixture.SCALES})
    o = run_model(model, 255, 127, 1000)
    want = {"head_cls": 2.0 / o["cls"].std(), "head_loc": 1.0 / o["loc_c"].std(), "head_lp_cls": 2.0 / o["cls_lp"].std(),
            "head_lp_loc": 0.5 / o["loc_lp"].std(), "fc": 4.0 / np.abs(o["offsets"]).mean()}
    print("raw stds:", {k: float(o[k].std()) for k in ("cls", "loc_c", "cls_lp", "loc_lp", "offsets")})
    print("SCALES =", {k: float("%.4g" % v) for k, v in want.items()})


def gen_keys():
    model, _ = build_reference_model()
    sd = model.state_dict()
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as fh:
        json.dump({k: list(v.shape) for k, v in sd.items()}, fh, indent=0, sort_keys=True)
    print("state_dict_keys.json:", len(sd), "tensors")


def gen_model(tag, instance, exemplar, seed):
    model, _ = build_reference_model(instance, exemplar)
    o = run_model(model, instance, exemplar, seed)
    o.update(seed=np.int64(seed), instance=np.int32(instance), exemplar=np.int32(exemplar))
    np.savez_compressed(os.path.join(OUT, "model_%s.npz" % tag), **o)
    print("model_%s: cls %s std %.3f | loc std %.3f | cls_lp %s std %.3f | loc_lp std %.3f | offsets %s | homo_score %.4f" % (
        tag, o["cls"].shape, o["cls"].std(), o["loc_c"].std(), o["cls_lp"].shape, o["cls_lp"].std(), o["loc_lp"].std(),
        np.round(o["offsets"], 2), o["homo_score"]))


def gen_tracker(seed=7, n_frames=8):
    model, cfg = build_reference_model()
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    torch.set_num_threads(8)
    tracker = build_tracker(model)
    frames, polys = synth.sequence(seed, n_frames)
    rec = {"polygon": [], "best_score": [], "H_total": [], "center_pos": [], "rot": [], "scale": []}
    with torch.no_grad():
        for idx, (img, gt) in enumerate(zip(frames, polys)):
            if idx == 0:  # exactly tools/test.py:118-130
                cx, cy, w, h = get_min_max_bbox(np.array(gt))
                gt_poly = get_w_h_from_poly(np.array(gt))
                tracker.init(img, [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], gt_poly, gt, np.array([gt[:2]]))
                continue
            o = tracker.track_new(idx, img, None, None, None)
            rec["polygon"].append(np.asarray(o["polygon"], np.float64))
            rec["best_score"].append(float(o["best_score"]))
            rec["H_total"].append(np.asarray(tracker.H_total, np.float64))
            rec["center_pos"].append(np.asarray(tracker.center_pos, np.float64))
            rec["rot"].append(float(tracker.rot))
            rec["scale"].append(float(tracker.scale))
            print("frame %d best_score %.4f polygon %s" % (idx, o["best_score"], np.round(o["polygon"].reshape(-1), 1)))
    np.savez_compressed(os.path.join(OUT, "tracker_seq%d.npz" % seed), seed=np.int64(seed), n_frames=np.int32(n_frames), gt=polys,
                        **{k: np.asarray(v) for k, v in rec.items()})


GATE_EVENTS = {6: "occlude", 12: "invert", 18: "flat", 24: "occlude", 27: "invert"}


def gen_tracker_gates(seed=13, n_frames=31, scales=None):
    """Second tracker golden ('gates' weight calibration, hdn_b200/synthetic.py GATES): every frame is ONE independent step of
    hdnTrackerHomo.track_new from a known state -- H_total is set to the ground-truth homography of the previous frame (what a
    perfect tracker would hold), so the steps stay realistic although the weights are untrained, and a last-bit difference cannot
    amplify over frames.  The log-polar head is confident on ordinary frames (non-zero rotation, scale != 1 -> decode_logpolar,
    H_sim, the rotated / re-scaled stage-3 crop) and the event frames trip `lp score < 0.25` and `homo_score > 2.5`
    (hdn_tracker_proj_e2e.py:203,261).  `pscore < 0.05` (:176) cannot fire with the shipped WINDOW_INFLUENCE: the Hanning term
    alone is 0.163 at the map centre."""
    import cv2
    model, cfg = build_reference_model(scales=scales, variant="gates")
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    torch.set_num_threads(8)
    tracker = build_tracker(model)
    frames, polys = synth.sequence(seed, n_frames, events=GATE_EVENTS)
    log = {}
    ref_new, ref_lp, ref_proj = model.track_new, model.track_new_lp, model.track_proj

    def spy_new(x):
        r = ref_new(x)
        sc = tracker._convert_score(r["cls"])
        ps = sc * (1 - cfg.TRACK.WINDOW_INFLUENCE) + tracker.window * cfg.TRACK.WINDOW_INFLUENCE
        log.update(idx=int(np.argmax(ps)), pscore=float(ps.max()))
        return r

    def spy_lp(x, d):
        r = ref_lp(x, d)
        sc = tracker._convert_score(r["cls_lp"])
        log.update(idx_lp=int(np.argmax(sc)), lp_score=float(sc.max()))
        return r

    def spy_proj(d, m):
        r = ref_proj(d, m)
        log.update(homo_score=float(r[1]))
        return r

    model.track_new, model.track_new_lp, model.track_proj = spy_new, spy_lp, spy_proj
    keys = ("polygon", "best_score", "H_pre", "H_total", "idx", "pscore", "idx_lp", "lp_score", "homo_score", "rot_delta", "scale_delta")
    rec = {k: [] for k in keys}
    init_pts = polys[0].reshape(4, 2).astype(np.float32)
    with torch.no_grad():
        gt = polys[0]
        cx, cy, w, h = get_min_max_bbox(np.array(gt))
        tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
        for idx in range(1, n_frames):
            H_pre = cv2.getPerspectiveTransform(init_pts, polys[idx - 1].reshape(4, 2).astype(np.float32))
            tracker.H_total = H_pre.copy()
            rot0, scale0 = tracker.rot, tracker.scale
            o = tracker.track_new(idx, frames[idx], None, None, None)
            rec["polygon"].append(np.asarray(o["polygon"], np.float64))
            rec["best_score"].append(float(o["best_score"]))
            rec["H_pre"].append(H_pre)
            rec["H_total"].append(np.asarray(tracker.H_total, np.float64))
            rec["rot_delta"].append(float(tracker.rot - rot0))
            rec["scale_delta"].append(float(tracker.scale / scale0))
            for k in ("idx", "pscore", "idx_lp", "lp_score", "homo_score"):
                rec[k].append(log[k])
            print("frame %2d %-8s idx %3d pscore %.4f | lp idx %3d score %.4f rot %+.4f scale %.4f | homo %.3f" % (
                idx, GATE_EVENTS.get(idx, ""), log["idx"], log["pscore"], log["idx_lp"], log["lp_score"], rec["rot_delta"][-1], rec["scale_delta"][-1],
                log["homo_score"]), flush=True)
    lp, hs = np.asarray(rec["lp_score"]), np.asarray(rec["homo_score"])
    print("lp gate fires on %d / %d frames, homo gate on %d, non-identity similarity on %d" % (
        (lp < 0.25).sum(), len(lp), (hs > 2.5).sum(), (np.asarray(rec["rot_delta"]) != 0).sum()))
    np.savez_compressed(os.path.join(OUT, "tracker_gates%d.npz" % seed), seed=np.int64(seed), n_frames=np.int32(n_frames), gt=polys,
                        events=np.asarray(sorted(GATE_EVENTS.items()), dtype="U16"), **{k: np.asarray(v) for k, v in rec.items()})


def gen_tracker_sim(seed=13, n_frames=7):
    """Similarity-only tracker (cfg.TRACK.TYPE = 'hdnTracker', hdn/tracker/hdn_tracker.py:110-301): init + free-running
    track_new (translation, scale / rotation, per-frame `update_template` from the rotated first frame), once with the default
    weight calibration (lp gate fires: identity similarity) and once with the 'gates' one (confident log-polar head: the box
    is re-scaled and rotated every frame; the tracker clamps the size to the frame, so the run stays bounded)."""
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    torch.set_num_threads(8)
    frames, polys = synth.sequence(seed, n_frames)
    out = {"seed": np.int64(seed), "n_frames": np.int32(n_frames), "gt": polys}
    for tag, variant in (("v0", None), ("v1", "gates")):
        model, cfg = build_reference_model(variant=variant)
        cfg.TRACK.TYPE = "hdnTracker"
        from hdn.tracker.tracker_builder import build_tracker
        tracker = build_tracker(model)
        assert type(tracker).__name__ == "hdnTracker"
        rec = {k: [] for k in ("polygon", "bbox", "best_score", "rot", "center_pos", "size")}
        with torch.no_grad():
            gt = polys[0]
            cx, cy, w, h = get_min_max_bbox(np.array(gt))
            tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), np.array([gt[:2]]))
            for idx in range(1, n_frames):
                o = tracker.track_new(idx, frames[idx], None, None)
                rec["polygon"].append(np.asarray(o["polygon"], np.float64))
                rec["bbox"].append(np.asarray(o["bbox"], np.float64))
                rec["best_score"].append(float(o["best_score"]))
                rec["rot"].append(float(o["rot"]))
                rec["center_pos"].append(np.asarray(tracker.center_pos, np.float64))
                rec["size"].append(np.asarray(tracker.size, np.float64))
                print("%s frame %d best_score %.4f rot %+.4f size %s centre %s" % (tag, idx, o["best_score"], o["rot"], np.round(tracker.size, 2),
                                                                                 np.round(tracker.center_pos, 2)), flush=True)
        out.update({"%s_%s" % (tag, k): np.asarray(v) for k, v in rec.items()})
        cfg.TRACK.TYPE = "hdnTrackerHomoProje2e"
    np.savez_compressed(os.path.join(OUT, "tracker_sim%d.npz" % seed), **out)


def gen_host():
    """Host-side (NumPy / OpenCV) helpers of the tracker: crops, gray packing, box / similarity algebra, point grids."""
    from hdn.core.config import cfg
    cfg.merge_from_file(YAML)
    cfg.CUDA = False
    from hdn.tracker.base_tracker import SiameseTracker
    from hdn.tracker.hdn_tracker import hdnTracker
    from hdn.utils import bbox as B, transform as T, point as P
    from homo_estimator.Deep_homography.Oneline_DLTv1.tools.get_img_info import get_search_info, get_template_info, merge_tmp_search
    frames, polys = synth.sequence(11, 2)
    img = frames[1]
    avg = np.mean(img, axis=(0, 1))
    st = SiameseTracker()
    o = {"avg": avg}
    cases = {"center": ([240.3, 180.9], 127, 181.0, 0), "border": ([20.0, 340.5], 255, 363.0, 0), "log": ([250.0, 170.0], 127, 150.0, 1),
             "same": ([200.0, 200.0], 127, 127, 0)}
    for name, (pos, msz, osz, islog) in cases.items():
        patch, box = st.get_subwindow_for_homo(img, np.array(pos), msz, osz, avg, islog)
        o["crop_%s" % name] = patch.numpy()
        o["crop_%s_box" % name] = np.asarray(box, np.float64)
        o["crop_%s_args" % name] = np.asarray(pos + [msz, osz, islog], np.float64)
    crop = st.get_subwindow(img, np.array([240.3, 180.9]), 127, 181.0, avg)
    g, shown = get_template_info(crop)
    g2, _ = get_search_info(crop)
    m = merge_tmp_search(g, g2)
    o.update(gray=g, shown=shown, merged_org=m["org_imgs"], merged_idx=np.asarray(m["patch_indices"], np.float64),
             merged_pts=np.asarray(m["four_points"], np.float64))
    poly = polys[1].astype(np.float64)
    o["poly"] = poly
    o["min_max"] = np.asarray(B.get_min_max_bbox(poly), np.float64)
    o["axis_aligned"] = np.asarray(B.get_axis_aligned_bbox(poly), np.float64)
    o["w_h_from_poly"] = np.asarray(B.get_w_h_from_poly(poly), np.float64)
    o["min_max_rect"] = np.asarray(B.get_min_max_bbox(np.array([10.0, 20.0, 30.0, 40.0])), np.float64)
    o["center2poly"] = B.cetner2poly([100.0, 80.0, 40.0, 20.0])
    o["rotmat"] = B.getRotMatrix(100.0, 80.0, 0.3)
    o["transform_poly"] = B.transformPoly(o["center2poly"], o["rotmat"])
    o["pts_xyxy"] = np.asarray(B.get_points_from_xyxy(np.array([10.0, 20.0, 30.0, 40.0])), np.float64)
    o["pts_xywh"] = np.asarray(B.get_points_from_xywh(np.array([10.0, 20.0, 30.0, 40.0])), np.float64)
    o["corner2center"] = np.asarray(B.corner2center(np.array([1.0, 2.0, 5.0, 10.0])), np.float64)
    o["center2corner"] = np.asarray(B.center2corner(np.array([3.0, 6.0, 4.0, 8.0])), np.float64)
    sims = [(100.0, 80.0, 0.0, 1.0, 3.0, -2.0), (100.0, 80.0, 0.2, 1.1, 3.0, -2.0), (100.0, 80.0, -0.4, 1.0, 0.0, 0.0),
            (100.0, 80.0, 0.0, 0.9, 1.5, 2.5), (100.0, 80.0, np.float32(0.13), np.float32(1.07), np.float64(2.0), np.float64(-1.0))]
    o["sim_args"] = np.asarray([[float(v) for v in s_] for s_ in sims], np.float64)
    o["sim_mats"] = np.asarray([T.rot_scale_around_center_shift_tran(*s_) for s_ in sims], np.float64)
    o["rot_img"] = T.img_rot_around_center(img, 240.0, 180.0, img.shape[1], img.shape[0], 0.25)[::4, ::4]
    o["mask_window"] = T.get_mask_window(60.7, 40.2, 0.3, 63.5, 63.5, 127, 127)
    o["points"] = P.generate_points(8, 25)
    o["points_lp"] = P.generate_points_lp(8, 8, 13)
    o["point_grid"] = P.Point(8, 25, 63).points

    class _S:
        cls_out_channels = 2
    rng = np.random.default_rng(5)
    loc4 = (rng.standard_normal((1, 4, 13, 13)) * 0.5).astype(np.float32)
    loc2 = rng.standard_normal((1, 2, 25, 25)).astype(np.float32)
    o["loc4"], o["loc2"] = loc4, loc2
    o["lp_decoded"] = hdnTracker._convert_logpolar_simi(_S(), t(loc4.copy()), o["points_lp"], 0, 0)
    o["c_decoded"] = SiameseTracker._convert_c(_S(), t(loc2.copy()), o["points"])
    from hdn.models.logpolar import getPolarImg
    o["polar_img"] = getPolarImg(crop[0].permute(1, 2, 0).numpy().astype(np.uint8))
    np.savez_compressed(os.path.join(OUT, "host_utils.npz"), **o)
    print("host_utils.npz:", len(o), "arrays, %.0f KB" % (os.path.getsize(os.path.join(OUT, "host_utils.npz")) / 1024))


if __name__ == "__main__":
    args = sys.argv[1:]
    if "--calibrate" in args:
        calibrate()
        sys.exit(0)
    which = args or ["keys", "native", "256", "tracker", "gates", "sim", "host"]
    if "keys" in which:
        gen_keys()
    if "native" in which:
        gen_model("native", 255, 127, 1000)
    if "256" in which:
        gen_model("256_512", 512, 256, 2000)
    if "tracker" in which:
        gen_tracker()
    if "gates" in which:
        gen_tracker_gates()
    if "sim" in which:
        gen_tracker_sim()
    if "host" in which:
        gen_host()
