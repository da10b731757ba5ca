"""Pin the oracle (C restatement + torch port) to goldens produced by the real reference.

CPU only.  The goldens in tests/golden/ops_*.npz were written by oracle/gen_golden.py,
which imports the unmodified reference from /root/reference.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close, golden_names, load_golden, regen_image
from oracle import c_oracle, torch_port

t = torch.from_numpy


@pytest.mark.parametrize("name", golden_names("ops_k1_") + golden_names("ops_k2_"))
def test_xcorr(name):
    g = load_golden(name)
    circ = bool(g["circular"])
    assert_close(c_oracle.xcorr_dw(g["x"], g["k"], circ), g["out"], what=name + " C")
    fn = torch_port.xcorr_depthwise_circular if circ else torch_port.xcorr_depthwise
    assert_close(fn(t(g["x"]), t(g["k"])).numpy(), g["out"], what=name + " torch")


def test_xcorr_shared_template_equals_tiled():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 4, 9, 9)).astype(np.float32)
    k = rng.standard_normal((1, 4, 3, 3)).astype(np.float32)
    a = c_oracle.xcorr_dw(x, k)
    b = c_oracle.xcorr_dw(x, np.tile(k, (3, 1, 1, 1)))
    assert np.array_equal(a, b)
    assert_close(torch_port.xcorr_depthwise(t(x), t(k)).numpy(), a)


@pytest.mark.parametrize("name", golden_names("ops_k3_"))
def test_logpolar(name):
    g = load_golden(name)
    img = regen_image(g)
    S = int(g["inst"]) // 2
    rot = float(g["delta"][1])
    assert_close(c_oracle.logpolar(img, g["polar"], rot, S), g["out"], what=name + " C")
    assert_close(torch_port.logpolar(t(img), t(g["polar"]), rot, S).numpy(), g["out"], what=name + " torch")


def reproject(H, src):
    """corners (x,y) -> H (x,y,1), [B,4,2]"""
    p = np.concatenate([src.reshape(-1, 4, 2), np.ones((src.shape[0], 4, 1))], 2).astype(np.float64)
    q = np.einsum("bij,bpj->bpi", H.astype(np.float64), p)
    return q[..., :2] / q[..., 2:3]


def test_dlt():
    g = load_golden("ops_k5_dlt")
    ref = g["H"][:, 0]
    for name, got in (("C", c_oracle.dlt4(g["src"], g["off"])),
                      ("torch", torch_port.dlt_solve(t(g["src"]), t(g["off"])).numpy()[:, 0])):
        # H with rtol 1e-3 / atol 1e-5, and reprojected corners within 1e-3 * 127 px
        assert np.allclose(got, ref, rtol=1e-3, atol=1e-5), name
        assert np.max(np.abs(reproject(got, g["src"]) - reproject(ref, g["src"]))) <= 1e-3 * 127, name
        # the defining property: H maps the (re-ordered) source corners onto src + off
        assert np.max(np.abs(reproject(got, g["src"]) - (g["src"] + g["off"]).reshape(-1, 4, 2))) < 2e-3, name


def _warp_mismatch_ok(got, ref, name):
    """The warp has jump discontinuities where a sample lands exactly on x == 0 / W-1 (weights from
    clamped corners); a last-bit difference in the sample coordinate flips those pixels.  Everything
    else must meet the tolerance; flipped pixels must be few and sit on such a boundary (one side is 0)."""
    atol = 1e-4 * np.max(np.abs(ref))
    bad = np.abs(got - ref) > atol + 1e-3 * np.abs(ref)
    assert bad.mean() < 2e-3, "%s: %.4f%% mismatching" % (name, 100 * bad.mean())
    assert np.all((np.abs(got[bad]) <= atol) | (np.abs(ref[bad]) <= atol)), name


@pytest.mark.parametrize("name", ["ops_k4_warp", "ops_k4_warp_small"])
def test_homo_warp(name):
    g = load_golden(name)
    M, Minv = (g["M"], g["M_inv"]) if "M" in g else c_oracle.default_M(127, 127)
    _warp_mismatch_ok(c_oracle.homo_warp(g["img"], g["H"], M, Minv), g["out"], name + " C")
    _warp_mismatch_ok(torch_port.homo_warp(t(g["img"]), t(g["H"]), t(np.asarray(M)), t(np.asarray(Minv))).numpy(), g["out"],
                      name + " torch")


def test_homo_warp_quirks():
    """identity H is NOT an identity warp (stretch by W/(W-1)); last row/col are 0."""
    g = load_golden("ops_k4_warp")
    out = g["out"][0]  # H[0] is the DLT of zero offsets
    tiny = 1e-6  # (A + B) - A - B in fp32: the cancelling weights leave rounding residue, not exact zeros
    assert np.all(np.abs(out[0, -1, :]) <= tiny) and np.all(np.abs(out[0, :, -1]) <= tiny)
    mine = c_oracle.homo_warp(g["img"][:1], g["H"][:1])[0]
    assert np.all(np.abs(mine[0, -1, :]) <= tiny) and np.all(np.abs(mine[0, :, -1]) <= tiny)
    # interior: sample of output col j sits at x = j * W/(W-1), i.e. a stretched copy, not a copy
    assert not np.allclose(mine[0, 5:100, 5:100], g["img"][0, 0, 5:100, 5:100], atol=1e-3)


def test_linspace_matches_torch():
    for n in (127, 24, 40, 5):
        ref = torch.linspace(-1.0, 1.0, n).numpy()
        step = np.float32(2.0) / np.float32(n - 1)
        mine = np.asarray([np.float32(-1.0) + step * np.float32(i) if i < n // 2 else np.float32(1.0) - step * np.float32(n - 1 - i)
                           for i in range(n)], np.float32)
        assert np.max(np.abs(mine - ref)) <= 1.2e-7


def test_score_argmax():
    g = load_golden("ops_k6_score")
    for name, fn in (("C", c_oracle.score_argmax), ("torch", lambda *a: torch_port.score_argmax(t(a[0]), t(a[1]), *a[2:]))):
        idx, ps, sc, gath = fn(g["cls"], g["loc"], g["window"], float(g["win_infl"]))
        assert np.array_equal(idx, g["idx"]), name  # bit-exact index
        assert np.allclose(ps, g["pscore"], rtol=1e-6, atol=1e-7), name
        assert np.allclose(sc, g["score"], rtol=1e-6, atol=1e-7), name
        # _convert_c (base_tracker.py:54-59): centre = point - 8 * loc at the arg-max
        centre = g["points"][idx] - 8.0 * gath
        assert np.allclose(centre, g["center"], rtol=1e-6, atol=1e-5), name
    assert g["idx"][1] == 312  # all-tie scores: the Hanning window alone decides -> centre cell
    assert g["pscore"][2] < 0.05 + 0.17  # background everywhere


def test_score_argmax_lp():
    g = load_golden("ops_k6_score_lp")
    idx, ps, sc, gath = c_oracle.score_argmax(g["cls"], g["loc"], None, 0.0)
    assert np.array_equal(idx, g["idx"])
    assert np.allclose(sc, g["score"], rtol=1e-6, atol=1e-7)
    # _convert_logpolar_simi (hdn_tracker.py:51-67): scale = exp((px - 8 loc0) * ln(E/2)/E), rot = (py - 8 loc2) * 2pi/E
    E = float(g["exemplar"])
    st = float(g["stride_lp"])
    pts = g["points"][idx]
    scale = np.exp((pts[:, 0] - gath[:, 0] * st) * (np.log(E / 2) / E))
    rot = (pts[:, 1] - gath[:, 2] * st) * (2 * np.pi / E)
    assert np.allclose(scale, g["sim"][:, 0], rtol=1e-5)
    assert np.allclose(rot, g["sim"][:, 2], rtol=1e-5, atol=1e-6)


def test_oracle_argmax_nan_semantics_match_numpy():
    """np.argmax (the reference's arg-max) treats NaN as the maximum and returns the first one; so does the oracle."""
    rng = np.random.default_rng(9)
    N, L, B = 13, 4, 3
    cls = (rng.standard_normal((B, 2, N, N)) * 2).astype(np.float32)
    loc = rng.standard_normal((B, L, N, N)).astype(np.float32)
    cls[0] = np.nan
    cls[1, 1, 3, 7] = np.nan
    cls[1, 0, 9, 1] = np.nan
    e = np.exp(cls - cls.max(1, keepdims=True))
    score = (e[:, 1] / e.sum(1)).reshape(B, -1)
    idx = c_oracle.score_argmax(cls, loc, None, 0.0)[0]
    assert np.array_equal(idx, np.argmax(score, 1))
    assert list(idx[:2]) == [0, 3 * N + 7]


def test_cpu_shim_runs_the_mirror_like_the_reference():
    """oracle/cpu_shim.py (bench.py's model / tracker-level CPU arm): the mirrored ModelBuilder + hdnTrackerHomo on CPU tensors,
    operators backed by torch_port, reproduce the REFERENCE's own outputs (model_native golden, first frames of tracker_seq7)."""
    import subprocess
    import sys
    from conftest import ROOT
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from conftest import load_golden, assert_close
from oracle import cpu_shim
from hdn_b200 import synthetic
model, cfg = cpu_shim.build_cpu_model()
g = load_golden("model_native")
seed = int(g["seed"])
with torch.no_grad():
    model.template(torch.from_numpy(synthetic.crop_tensor(seed, (1, 6, 127, 127))))
    out = model.track_new(torch.from_numpy(synthetic.crop_tensor(seed + 1, (1, 3, 255, 255))))
assert_close(out["cls"].numpy(), g["cls"], what="cls"); assert_close(out["loc_c"].numpy(), g["loc_c"], what="loc_c")
from hdn.tracker.tracker_builder import build_tracker
from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
t = load_golden("tracker_seq7")
tracker = build_tracker(model)
frames, polys = synthetic.sequence(int(t["seed"]), 3)
gt = polys[0]; cx, cy, w, h = get_min_max_bbox(np.array(gt))
tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
for i in (1, 2):
    o = tracker.track_new(i, frames[i], None, None, None)
    assert np.abs(np.asarray(o["polygon"], np.float64) - t["polygon"][i - 1]).max() <= 1e-3 * 600, i
print("OK")
''' % (ROOT, ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]
