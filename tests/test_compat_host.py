"""CPU-only: the mirror packages (hdn_b200/compat) against goldens produced by the real reference
(tests/golden/host_utils.npz, state_dict_keys.json; generator: oracle/gen_golden_model.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden
from hdn_b200 import compat

compat.activate()
YAML = os.path.join(ROOT, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml")


@pytest.fixture(scope="module")
def cfg():
    from hdn.core.config import cfg
    cfg.merge_from_file(YAML)
    return cfg


def test_config_surface(cfg):
    assert cfg.TRACK.TYPE == "hdnTrackerHomoProje2e" and cfg.TRACK.INSTANCE_SIZE == 255 and cfg.TRACK.EXEMPLAR_SIZE == 127
    assert cfg.BACKBONE.TYPE == "resnet50" and cfg.BACKBONE.KWARGS.used_layers == [2, 3, 4]
    assert cfg.BACKBONE_HOMO.TYPE == "resnet34" and cfg.BAN.KWARGS.cls_out_channels == 2 and cfg.BAN.KWARGS.weighted is True
    assert cfg.ADJUST.KWARGS.in_channels == [512, 1024, 2048] and cfg.POINT.STRIDE == 8 and cfg.TRAIN.OUTPUT_SIZE_LP == 13
    assert abs(cfg.TRACK.WINDOW_INFLUENCE - 0.1632532824922313) < 1e-15
    cfg.CUDA = False
    assert cfg.CUDA is False
    cfg.CUDA = True
    with pytest.raises(AttributeError):
        cfg.TRACK.NO_SUCH_KEY


def test_state_dict_keys_match_reference(cfg):
    """The reference checkpoint (hdn-simi-sup-hm-unsup.pth) must load: same 836 names and shapes."""
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    ours = {k: list(v.shape) for k, v in ModelBuilder().state_dict().items()}
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    assert len(ref) == 836
    assert sorted(ours) == sorted(ref)
    assert all(ours[k] == ref[k] for k in ref)


def test_load_pretrain_roundtrip(cfg, tmp_path):
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    from hdn.utils.model_load import load_pretrain
    a = ModelBuilder()
    path = str(tmp_path / "ckpt.pth")
    torch.save({"state_dict": {"module." + k: v for k, v in a.state_dict().items()}}, path)  # DataParallel-style checkpoint
    b = load_pretrain(ModelBuilder(), path)
    assert all(torch.equal(v, b.state_dict()[k]) for k, v in a.state_dict().items())


def test_crops_match_reference(cfg):
    from hdn.tracker.base_tracker import crop_window
    import synth
    g = load_golden("host_utils")
    frames, _ = synth.sequence(11, 2)
    img = frames[1]
    avg = np.mean(img, axis=(0, 1))
    assert np.array_equal(avg, g["avg"])
    for name in ("center", "border", "log", "same"):
        px, py, msz, osz, islog = g["crop_%s_args" % name]
        patch, box = crop_window(img, np.array([px, py]), int(msz), osz, avg, int(islog))
        assert np.array_equal(patch, g["crop_%s" % name]), name          # bit-exact: same OpenCV calls on the same pixels
        assert np.array_equal(np.asarray(box, np.float64), g["crop_%s_box" % name]), name


def test_gray_packing_matches_reference(cfg):
    from hdn.tracker.base_tracker import crop_window
    from homo_estimator.Deep_homography.Oneline_DLTv1.tools.get_img_info import get_search_info, get_template_info, merge_tmp_search
    import synth
    g = load_golden("host_utils")
    img = synth.sequence(11, 2)[0][1]
    crop = torch.from_numpy(crop_window(img, np.array([240.3, 180.9]), 127, 181.0, np.mean(img, axis=(0, 1)))[0])
    gray, shown = get_template_info(crop)
    assert np.array_equal(gray, g["gray"]) and np.array_equal(shown, g["shown"])
    m = merge_tmp_search(gray, get_search_info(crop)[0])
    assert np.array_equal(m["org_imgs"], g["merged_org"]) and np.array_equal(m["input_tensors"], g["merged_org"])
    assert np.array_equal(np.asarray(m["patch_indices"], np.float64), g["merged_idx"])
    assert np.array_equal(np.asarray(m["four_points"], np.float64), g["merged_pts"])


def test_geometry_helpers_match_reference(cfg):
    from hdn.utils import bbox as B, point as P, transform as T
    import synth
    g = load_golden("host_utils")
    poly = g["poly"]
    eq = lambda a, b: np.array_equal(np.asarray(a, np.float64), b)  # noqa: E731
    assert eq(B.get_min_max_bbox(poly), g["min_max"]) and eq(B.get_axis_aligned_bbox(poly), g["axis_aligned"])
    assert eq(B.get_w_h_from_poly(poly), g["w_h_from_poly"]) and eq(B.get_min_max_bbox(np.array([10.0, 20.0, 30.0, 40.0])), g["min_max_rect"])
    assert eq(B.cetner2poly([100.0, 80.0, 40.0, 20.0]), g["center2poly"]) and eq(B.getRotMatrix(100.0, 80.0, 0.3), g["rotmat"])
    assert eq(B.transformPoly(g["center2poly"], g["rotmat"]), g["transform_poly"])
    assert eq(B.get_points_from_xyxy(np.array([10.0, 20.0, 30.0, 40.0])), g["pts_xyxy"])
    assert eq(B.get_points_from_xywh(np.array([10.0, 20.0, 30.0, 40.0])), g["pts_xywh"])
    assert eq(B.corner2center(np.array([1.0, 2.0, 5.0, 10.0])), g["corner2center"]) and eq(B.center2corner(np.array([3.0, 6.0, 4.0, 8.0])), g["center2corner"])
    for args, ref in zip(g["sim_args"], g["sim_mats"]):
        assert np.array_equal(T.rot_scale_around_center_shift_tran(*args), ref)
    img = synth.sequence(11, 2)[0][1]
    assert np.array_equal(T.img_rot_around_center(img, 240.0, 180.0, img.shape[1], img.shape[0], 0.25)[::4, ::4], g["rot_img"])
    assert np.array_equal(T.get_mask_window(60.7, 40.2, 0.3, 63.5, 63.5, 127, 127), g["mask_window"])
    assert np.array_equal(P.generate_points(8, 25), g["points"]) and np.array_equal(P.generate_points_lp(8, 8, 13), g["points_lp"])
    assert np.array_equal(P.Point(8, 25, 63).points, g["point_grid"])


def test_single_column_decoders_equal_whole_map_decoders(cfg):
    """The tracker decodes only the arg-max column (after the K6 gather); it must equal the reference's whole-map decode there."""
    from hdn.tracker.hdn_tracker import decode_center, decode_logpolar
    g = load_golden("host_utils")
    loc4, loc2 = g["loc4"].reshape(4, -1), g["loc2"].reshape(2, -1)
    for idx in (0, 7, 84, 168):
        assert np.array_equal(decode_logpolar(g["points_lp"], idx, loc4[:, idx])[:3], g["lp_decoded"][:3, idx])
    for idx in (0, 312, 624):
        assert np.array_equal(decode_center(g["points"], idx, loc2[:, idx]), g["c_decoded"][:, idx])


def test_log_polar_template_image(cfg):
    from hdn.models.logpolar import getPolarImg
    from hdn.tracker.base_tracker import crop_window
    import synth
    g = load_golden("host_utils")
    img = synth.sequence(11, 2)[0][1]
    crop = crop_window(img, np.array([240.3, 180.9]), 127, 181.0, np.mean(img, axis=(0, 1)))[0]
    assert np.array_equal(getPolarImg(crop[0].transpose(1, 2, 0).astype(np.uint8)), g["polar_img"])


def test_tracker_builder_dispatch(cfg):
    from hdn.tracker.tracker_builder import TRACKS
    assert set(TRACKS) == {"hdnTracker", "hdnTrackerHomoProje2e"}


def _crop_full_canvas(im, pos, model_sz, original_sz, avg):
    """The reference's construction (base_tracker.py:76-118): pad the WHOLE frame, then slice.  Used to stress crop_window."""
    import cv2
    r, c, k = im.shape
    half = (original_sz - 1) / 2
    x0 = np.floor(pos[0] - half + 0.5); x1 = x0 + original_sz - 1
    y0 = np.floor(pos[1] - half + 0.5); y1 = y0 + original_sz - 1
    lp, tp = int(max(0., -x0)), int(max(0., -y0))
    rp, bp = int(max(0., x1 - c + 1)), int(max(0., y1 - r + 1))
    x0, x1, y0, y1 = x0 + lp, x1 + lp, y0 + tp, y1 + tp
    if any([tp, bp, lp, rp]):
        te = np.zeros((r + tp + bp, c + lp + rp, k), np.uint8)
        te[tp:tp + r, lp:lp + c, :] = im
        if tp: te[0:tp, lp:lp + c, :] = avg
        if bp: te[r + tp:, lp:lp + c, :] = avg
        if lp: te[:, 0:lp, :] = avg
        if rp: te[:, c + lp:, :] = avg
        patch = te[int(y0):int(y1 + 1), int(x0):int(x1 + 1), :]
    else:
        patch = im[int(y0):int(y1 + 1), int(x0):int(x1 + 1), :]
    if not np.array_equal(model_sz, original_sz):
        patch = cv2.resize(patch, (model_sz, model_sz))
    return patch.transpose(2, 0, 1)[np.newaxis].astype(np.float32), (x0, y0, x1 + 1, y1 + 1)


def test_crop_window_equals_full_canvas_construction(cfg):
    from hdn.tracker.base_tracker import crop_window
    import synth
    img = synth.texture(3, 90, 130)
    avg = np.mean(img, axis=(0, 1))
    rng = np.random.default_rng(0)
    for _ in range(300):
        pos = np.array([rng.uniform(-60, 190), rng.uniform(-60, 150)])
        osz = float(np.floor(rng.uniform(5, 140))) if rng.random() < 0.5 else float(rng.uniform(5, 140))  # stage 3 passes fractional sizes
        a, abox = crop_window(img, pos, 31, osz, avg)
        b, bbox = _crop_full_canvas(img, pos, 31, osz, avg)
        assert np.array_equal(a, b) and tuple(abox) == tuple(bbox)
