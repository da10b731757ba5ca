"""GPU: the mirrored ModelBuilder / hdnTrackerHomo (hdn_b200/compat) against goldens produced by running the
REFERENCE's ModelBuilder / tracker on CPU with the same seeded weights and inputs (oracle/gen_golden_model.py).

Tolerance: rtol 1e-3 / atol 1e-4*max|ref| on network outputs (cuDNN fp32 vs oneDNN fp32 through ~50 layers);
arg-max indices bit-exact; tracker polygons within 1e-3 relative (of the frame diagonal) per frame.
"""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_close, load_golden
from hdn_b200 import compat

compat.activate()
pytestmark = pytest.mark.gpu
YAML = os.path.join(ROOT, "experiments", "tracker_homo_config", "proj_e2e_GOT_unconstrained_v2.yaml")


def build_model(instance=255, exemplar=127, variant=None):
    from hdn.core.config import cfg
    import weights_fixture
    cfg.merge_from_file(YAML)
    cfg.TRACK.INSTANCE_SIZE, cfg.TRACK.EXEMPLAR_SIZE = instance, exemplar
    cfg.CUDA = True
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    model = weights_fixture.fill(ModelBuilder(), variant=variant).cuda().eval()
    return model, cfg


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["native", "256_512"])
def test_model_level_parity(tag):
    import synth
    from oracle import c_oracle
    g = load_golden("model_" + tag)
    inst, ex, seed = int(g["instance"]), int(g["exemplar"]), int(g["seed"])
    model, cfg = build_model(inst, ex)
    z = synth.crop_tensor(seed, (1, 6, ex, ex))
    x = synth.crop_tensor(seed + 1, (1, 3, inst, inst))
    model.template(cuda(z))
    stride = int(g["feature_stride"])
    for i, f in enumerate(model.zf):
        assert np.allclose([f.mean().item(), f.std().item(), f.abs().max().item()], g["zf%d_stats" % i], rtol=1e-3)
        assert_close(f.reshape(-1)[::stride].cpu().numpy(), g["zf%d_sample" % i], what="zf[%d] element-wise" % i)
    for i, f in enumerate(model.zf_lp):
        assert_close(f.reshape(-1)[::stride].cpu().numpy(), g["zf_lp%d_sample" % i], what="zf_lp[%d] element-wise" % i)
    out = model.track_new(cuda(x))
    assert_close(out["cls"].cpu().numpy(), g["cls"], what="cls")
    assert_close(out["loc_c"].cpu().numpy(), g["loc_c"], what="loc_c")
    lp = model.track_new_lp(cuda(x), [0, 0])
    assert_close(lp["cls_lp"].cpu().numpy(), g["cls_lp"], what="cls_lp")
    assert_close(lp["loc_lp"].cpu().numpy(), g["loc_lp"], what="loc_lp")
    assert np.allclose([lp["x_lp"].mean().item(), lp["x_lp"].std().item()], g["x_lp_stats"], rtol=1e-4)
    assert_close(lp["x_lp"].reshape(-1)[::stride].cpu().numpy(), g["x_lp_sample"], what="x_lp element-wise")
    # arg-max indices on the reference maps and on ours must be the same cell (bit-exact requirement)
    N = g["cls"].shape[-1]
    win = np.outer(np.hanning(N), np.hanning(N)).flatten()
    ref_idx = c_oracle.score_argmax(g["cls"], g["loc_c"], win, cfg.TRACK.WINDOW_INFLUENCE)[0]
    idx, ps, sc, gath = model.track_new_scored(cuda(x))
    assert int(idx[0]) == int(ref_idx[0])
    ref_idx_lp = c_oracle.score_argmax(g["cls_lp"], g["loc_lp"], None, 0.0)[0]
    idx_lp, _, _, _ = model.track_new_lp_scored(cuda(x))
    assert int(idx_lp[0]) == int(ref_idx_lp[0])
    # residual homography head
    rng = np.random.default_rng(seed + 2)
    pair = rng.standard_normal((1, 2, 127, 127)).astype(np.float32)
    h4p = np.asarray([[0, 0, 0, 127, 127, 127, 127, 0]], np.float32)
    data = {"org_imgs": cuda(pair), "input_tensors": cuda(pair), "h4p": cuda(h4p), "patch_indices": None}
    H, s_homo, s_simi = model.track_proj(data, None)
    with torch.no_grad():
        off, _, _ = model.hm_net.offsets(cuda(pair))
    assert np.allclose(off.cpu().numpy(), g["offsets"], rtol=1e-3, atol=1e-3)
    assert np.allclose(H.cpu().numpy(), g["H"], rtol=1e-3, atol=1e-5)
    assert abs(float(s_homo) - float(g["homo_score"])) <= 1e-3 * abs(float(g["homo_score"])) + 1e-5
    assert abs(float(s_simi) - float(g["simi_score"])) <= 1e-3 * abs(float(g["simi_score"])) + 1e-5


def test_template_kernels_are_cached_and_equal_recomputed():
    import synth
    model, _ = build_model()
    model.template(cuda(synth.crop_tensor(1000, (1, 6, 127, 127))))
    x = cuda(synth.crop_tensor(1001, (1, 3, 255, 255)))
    with torch.no_grad():  # inference: template() cached its kernels on the same (tensor-core) conv path
        xf = model.neck(model.backbone(x))
        a = model.head(model.zf, xf, model._k_sim)
        b = model.head(model.zf, xf, None)  # the reference's per-frame recomputation
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_tracker_trajectory_parity():
    """init + 7 frames of the homography tracker on a synthetic sequence: same polygons as the reference tracker."""
    import synth
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    g = load_golden("tracker_seq7")
    model, cfg = build_model()
    tracker = build_tracker(model)
    frames, polys = synth.sequence(int(g["seed"]), int(g["n_frames"]))
    assert np.array_equal(polys, g["gt"])
    diag = float(np.hypot(*frames[0].shape[:2]))
    for idx, (img, gt) in enumerate(zip(frames, polys)):
        if idx == 0:
            cx, cy, w, h = get_min_max_bbox(np.array(gt))
            tracker.init(img, [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
            continue
        o = tracker.track_new(idx, img, None, None, None)
        ref_poly = g["polygon"][idx - 1]
        err = np.abs(np.asarray(o["polygon"], np.float64) - ref_poly).max()
        assert err <= 1e-3 * diag, "frame %d: polygon off by %.4f px" % (idx, err)
        assert abs(float(o["best_score"]) - g["best_score"][idx - 1]) < 1e-3
        assert np.allclose(tracker.H_total, g["H_total"][idx - 1], rtol=1e-3, atol=1e-3 * np.abs(g["H_total"][idx - 1]).max())
        assert set(o) == {"bbox_aligned", "best_score", "polygon", "points", "bbox"}


def test_tracker_gate_steps_parity():
    """30 independent steps of hdnTrackerHomo.track_new from known states (H_total = ground-truth homography of the previous
    frame) with the 'gates' weight calibration: the log-polar head is confident on ordinary frames (non-zero rotation,
    scale != 1 -> decode_logpolar, H_sim, the rotated / re-scaled stage-3 crop) and the event frames trip `lp score < 0.25`
    and `homo_score > 2.5` (hdn_tracker_proj_e2e.py:203,261).  Arg-max indices and both gate decisions bit-exact; polygons
    and H_total within 1e-3.  (`pscore < 0.05` cannot fire with the shipped WINDOW_INFLUENCE: 0.163 * hanning(centre) > 0.05.)"""
    import synth
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    g = load_golden("tracker_gates13")
    events = {int(k): str(v) for k, v in g["events"]}
    model, cfg = build_model(variant="gates")
    tracker = build_tracker(model)
    frames, polys = synth.sequence(int(g["seed"]), int(g["n_frames"]), events=events)
    assert np.array_equal(polys, g["gt"])
    seen = {}
    s1, s2, s3 = model.track_new_scored, model.track_new_lp_scored, model.track_proj_packed

    def spy1(x, w=None):
        r = s1(x, w)
        seen.update(idx=int(r[0][0]), pscore=float(r[1][0]))
        return r

    def spy2(x, d=[0, 0]):
        r = s2(x, d)
        seen.update(idx_lp=int(r[0][0]), lp_score=float(r[2][0]))
        return r

    def spy3(pair, h4p):
        r = s3(pair, h4p)
        seen.update(homo_score=float(r[0][9]))
        return r

    model.track_new_scored, model.track_new_lp_scored, model.track_proj_packed = spy1, spy2, spy3
    gt = polys[0]
    cx, cy, w, h = get_min_max_bbox(np.array(gt))
    tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
    diag = float(np.hypot(*frames[0].shape[:2]))
    n_lp_gate = n_homo_gate = n_sim = 0
    for idx in range(1, len(frames)):
        i = idx - 1
        tracker.H_total = g["H_pre"][i].copy()
        rot0, scale0 = tracker.rot, tracker.scale
        o = tracker.track_new(idx, frames[idx], None, None, None)
        assert seen["idx"] == int(g["idx"][i]), "frame %d: stage-1 arg-max %d != %d" % (idx, seen["idx"], int(g["idx"][i]))
        assert seen["idx_lp"] == int(g["idx_lp"][i]), "frame %d: log-polar arg-max" % idx
        assert abs(seen["pscore"] - g["pscore"][i]) < 1e-3 and abs(seen["lp_score"] - g["lp_score"][i]) < 1e-3
        assert (seen["lp_score"] < 0.25) == (g["lp_score"][i] < 0.25), "frame %d: lp gate" % idx
        assert (seen["homo_score"] > 2.5) == (g["homo_score"][i] > 2.5), "frame %d: homo gate" % idx
        assert abs(seen["homo_score"] - g["homo_score"][i]) <= 2e-3 * g["homo_score"][i]
        assert abs((tracker.rot - rot0) - g["rot_delta"][i]) < 1e-4 and abs(tracker.scale / scale0 - g["scale_delta"][i]) < 1e-4 * g["scale_delta"][i]
        ref_poly = g["polygon"][i]
        scale = max(diag, float(np.abs(ref_poly).max()))
        assert np.abs(np.asarray(o["polygon"], np.float64) - ref_poly).max() <= 1e-3 * scale, "frame %d polygon" % idx
        assert np.allclose(tracker.H_total, g["H_total"][i], rtol=1e-3, atol=1e-3 * np.abs(g["H_total"][i]).max())
        n_lp_gate += g["lp_score"][i] < 0.25
        n_homo_gate += g["homo_score"][i] > 2.5
        n_sim += g["rot_delta"][i] != 0
    assert n_lp_gate >= 2 and n_homo_gate >= 5 and n_sim >= 20  # the golden really covers the branches


@pytest.mark.parametrize("tag,variant", [("v0", None), ("v1", "gates")])
def test_similarity_tracker_parity(tag, variant):
    """cfg.TRACK.TYPE = 'hdnTracker' (hdn/tracker/hdn_tracker.py:110-301): init + 6 free-running frames incl. the per-frame
    `update_template` from the rotated first frame, against the reference tracker's own trajectory; v1 = confident log-polar
    head (the box is rotated / re-scaled every frame and clamped to the frame size)."""
    import synth
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    g = load_golden("tracker_sim13")
    model, cfg = build_model(variant=variant)
    cfg.TRACK.TYPE = "hdnTracker"
    try:
        tracker = build_tracker(model)
        assert type(tracker).__name__ == "hdnTracker"
        frames, polys = synth.sequence(int(g["seed"]), int(g["n_frames"]))
        assert np.array_equal(polys, g["gt"])
        gt = polys[0]
        cx, cy, w, h = get_min_max_bbox(np.array(gt))
        tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), np.array([gt[:2]]))
        diag = float(np.hypot(*frames[0].shape[:2]))
        for idx in range(1, len(frames)):
            o = tracker.track_new(idx, frames[idx], None, None)
            i = idx - 1
            assert np.abs(np.asarray(o["polygon"], np.float64) - g[tag + "_polygon"][i]).max() <= 1e-3 * diag, "frame %d polygon" % idx
            assert np.abs(np.asarray(o["bbox"], np.float64) - g[tag + "_bbox"][i]).max() <= 1e-3 * diag
            assert abs(float(o["best_score"]) - g[tag + "_best_score"][i]) < 1e-3
            assert abs(float(o["rot"]) - g[tag + "_rot"][i]) < 1e-4
            assert np.abs(tracker.size - g[tag + "_size"][i]).max() <= 1e-3 * diag
            assert set(o) == {"bbox", "bbox_aligned", "best_score", "rot", "polygon"}
    finally:
        cfg.TRACK.TYPE = "hdnTrackerHomoProje2e"


def test_device_preprocessing_equals_host_opencv_path():
    """the frame is uploaded once and warpPerspective / crops / cubic rotation / gray normalisation run on the
    device, bit-compatible with OpenCV -- so the tracker's trajectory is IDENTICAL to the host-OpenCV path, frame by frame."""
    import synth
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    frames, polys = synth.sequence(41, 7, size=(720, 1280), obj=(240, 320))
    model, _ = build_model(variant="gates")  # confident log-polar head: non-zero rotation -> the cubic warpAffine really rotates

    def run(device_side):
        tracker = build_tracker(model)
        tracker.device_preproc = device_side
        gt = polys[0]
        cx, cy, w, h = get_min_max_bbox(np.array(gt))
        tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
        out = []
        for i in range(1, len(frames)):
            out.append(np.asarray(tracker.track_new(i, frames[i], None, None, None)["polygon"], np.float64))
        assert (tracker._pre is not None) == device_side
        return np.asarray(out), tracker.rot

    host, rot_h = run(False)
    dev, rot_d = run(True)
    assert rot_h != 0 and rot_h == rot_d
    assert np.array_equal(host, dev)


def test_cuda_graph_stages_equal_eager():
    """Replaying the three network stages from CUDA graphs must not change a single output value."""
    import synth
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    frames, polys = synth.sequence(21, 5)

    def run(graphs):
        model, _ = build_model()
        model.enable_graphs(graphs)
        tracker = build_tracker(model)
        gt = polys[0]
        cx, cy, w, h = get_min_max_bbox(np.array(gt))
        tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
        return np.asarray([tracker.track_new(i, frames[i], None, None, None)["polygon"] for i in range(1, len(frames))])

    eager, graphed = run(False), run(True)
    assert np.array_equal(eager, graphed)


def test_sanitizer_smoke_shapes():
    """Small staged-kernel cases kept cheap enough to run under compute-sanitizer (profiles/r01_sanitizer_*.log)."""
    from hdn_b200 import ops
    x = torch.randn(1, 256, 61, 61, device="cuda")
    k = torch.randn(1, 256, 29, 29, device="cuda")
    out = ops.xcorr_depthwise(x, k)
    ref = torch.nn.functional.conv2d(x.view(1, 256, 61, 61), k.view(256, 1, 29, 29), groups=256)
    assert torch.allclose(out, ref.view_as(out), rtol=1e-3, atol=1e-4 * float(ref.abs().max()))


def test_lockstep_batch_equals_single_sequence_tracking():
    """4 sequences advanced in lock-step (one batched network call per stage) == each tracked alone.  Every kernel computes a pair's
    outputs in the same order whatever the batch size -- except that the convolution DISPATCH depends on the launch size (batch-1
    layers split K over a cluster, large launches take the tensor-memory-operand kernel: other fp32 summation orders), and the
    free-running tracker of UNTRAINED weights turns a last-bit difference at a near-tie of the arg-max into pixels.  So: with the
    size-dependent kernels off the trajectories must agree within fp32 noise over all frames; with the default dispatch on the first."""
    import synth
    from hdn.tracker.tracker_builder import build_tracker
    from hdn_b200 import ops, runner
    model, _ = build_model()
    seqs = [synth.sequence(30 + s, 5) for s in range(4)]
    diag = float(np.hypot(*seqs[0][0][0].shape[:2]))
    try:
        ops.set_conv_splitk(False)
        ops.set_conv_ts(False)
        polys, _ = runner.track_lockstep(model, seqs)
        for s, (frames, gt) in enumerate(seqs):
            alone, _ = runner.track_sequence(build_tracker(model), frames, gt)
            assert np.abs(polys[s] - alone).max() <= 1e-3 * diag, s
    finally:
        ops.set_conv_splitk(True)
        ops.set_conv_ts(2)
    polys, _ = runner.track_lockstep(model, seqs)
    for s, (frames, gt) in enumerate(seqs):
        alone, _ = runner.track_sequence(build_tracker(model), frames, gt)
        assert np.abs(polys[s][:1] - alone[:1]).max() <= 1e-3 * diag, s
