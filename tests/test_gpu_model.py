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


def build_model(instance=255, exemplar=127):
    from hdn.core.config import cfg
    import weights_fixture
    cfg.merge_from_file(YAML)
    cfg.TRACK.INSTANCE_SIZE, cfg.TRACK.EXEMPLAR_SIZE = instance, exemplar
    cfg.CUDA = True
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    model = weights_fixture.fill(ModelBuilder()).cuda().eval()
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
    for i, f in enumerate(model.zf):
        assert np.allclose([f.mean().item(), f.std().item(), f.abs().max().item()], g["zf%d_stats" % i], rtol=1e-3)
    out = model.track_new(cuda(x))
    assert_close(out["cls"].cpu().numpy(), g["cls"], what="cls")
    assert_close(out["loc_c"].cpu().numpy(), g["loc_c"], what="loc_c")
    lp = model.track_new_lp(cuda(x), [0, 0])
    assert_close(lp["cls_lp"].cpu().numpy(), g["cls_lp"], what="cls_lp")
    assert_close(lp["loc_lp"].cpu().numpy(), g["loc_lp"], what="loc_lp")
    assert np.allclose([lp["x_lp"].mean().item(), lp["x_lp"].std().item()], g["x_lp_stats"], rtol=1e-4)
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
    """4 sequences advanced in lock-step (one batched network call per stage) == each tracked alone, within fp32 noise."""
    import synth
    from hdn.tracker.tracker_builder import build_tracker
    from hdn_b200 import runner
    model, _ = build_model()
    seqs = [synth.sequence(30 + s, 5) for s in range(4)]
    polys, _ = runner.track_lockstep(model, seqs)
    diag = float(np.hypot(*seqs[0][0][0].shape[:2]))
    for s, (frames, gt) in enumerate(seqs):
        alone, _ = runner.track_sequence(build_tracker(model), frames, gt)
        assert np.abs(polys[s] - alone).max() <= 1e-3 * diag, s
