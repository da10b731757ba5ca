"""GPU parity: the CUDA path (through the C ABI, via hdn_b200.ops) against
  (1) goldens produced by the real reference (tests/golden/ops_*.npz),
  (2) the C oracle on seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances are SURVEY.md 8(d): rtol 1e-3, atol 1e-4*max|ref|; arg-max indices bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close, golden_names, load_golden, regen_image
from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g2d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def ops():
    from hdn_b200 import ops as o
    return o


# ------------------------------------------------------------------ K1 / K2
@pytest.fixture(params=["fft", "direct"])
def algo(request):
    """Both correlation algorithms: 'fft' = FFT kernel wherever one exists (29x29 / 15x15 templates), 'direct' = direct sum everywhere."""
    ops().set_xcorr_algo(request.param)
    yield request.param
    ops().set_xcorr_algo("auto")


def uses_fft(C, Hx, Wx, Hk, Wk, circ):
    from hdn_b200 import _lib
    return bool(_lib.lib().hdn_xcorr_uses_fft(C, Hx, Wx, Hk, Wk, int(circ), C * Hk * Wk))


def test_fft_kernel_is_the_default_for_the_fma_bound_shapes():
    ops().set_xcorr_algo("auto")
    for Hx, Hk, circ in ((61, 29, 0), (29, 29, 1), (39, 15, 0)):
        assert uses_fft(256, Hx, Hx, Hk, Hk, circ)
    for Hx, Hk, circ in ((29, 5, 0), (13, 13, 1)):
        assert not uses_fft(256, Hx, Hx, Hk, Hk, circ)  # HBM-bound / tiny: direct sum
    ops().set_xcorr_algo("direct")
    assert not uses_fft(256, 61, 61, 29, 29, 0)
    ops().set_xcorr_algo("auto")


@pytest.mark.parametrize("name", golden_names("ops_k1_") + golden_names("ops_k2_"))
def test_xcorr_golden(name, algo):
    g = load_golden(name)
    fn = ops().xcorr_depthwise_circular if int(g["circular"]) else ops().xcorr_depthwise
    out = fn(g2d(g["x"]), g2d(g["k"])).cpu().numpy()
    assert_close(out, g["out"], what=name)


FAST_SHAPES = [  # (B, C, Hx, Wx, Hk, Wk, circular)  -- the shapes with staged TMA kernels, C = 256 as in the network
    (2, 256, 29, 29, 5, 5, 0),
    (3, 256, 13, 13, 13, 13, 1),
    (1, 256, 61, 61, 29, 29, 0),
    (1, 256, 29, 29, 29, 29, 1),
    (1, 256, 39, 39, 15, 15, 0),
    (5, 24, 29, 29, 5, 5, 0),      # C % G == 0 with small C
    (2, 20, 29, 29, 5, 5, 0),      # C % G != 0 -> generic kernel
]


@pytest.mark.parametrize("shape", FAST_SHAPES)
@pytest.mark.parametrize("shared", [False, True])
def test_xcorr_vs_oracle(shape, shared, algo):
    B, C, Hx, Wx, Hk, Wk, circ = shape
    rng = np.random.default_rng(hash(shape) % 2**31)
    x = rng.standard_normal((B, C, Hx, Wx)).astype(np.float32)
    if algo == "fft":
        x += 0.5  # post-ReLU-like features with a DC component: the hard case for a transform-domain product
    k = (rng.standard_normal((1 if shared else B, C, Hk, Wk)) * 0.1).astype(np.float32)
    ref = c_oracle.xcorr_dw(x, k, bool(circ))
    fn = ops().xcorr_depthwise_circular if circ else ops().xcorr_depthwise
    out = fn(g2d(x), g2d(k)).cpu().numpy()
    assert_close(out, ref, what=str(shape))
    if algo == "fft" and uses_fft(C, Hx, Wx, Hk, Wk, circ):  # the FFT route is fp32-accurate, not merely within 1e-3
        assert np.abs(out - ref).max() <= 5e-6 * np.abs(ref).max()


def test_xcorr_multi_equals_single():
    rng = np.random.default_rng(11)
    xs = [g2d(rng.standard_normal((2, 256, 29, 29)).astype(np.float32)) for _ in range(6)]
    ks = [g2d(rng.standard_normal((2, 256, 5, 5)).astype(np.float32)) for _ in range(6)]
    multi = ops().xcorr_depthwise_multi(xs, ks)
    for x, k, m in zip(xs, ks, multi):
        assert torch.equal(ops().xcorr_depthwise(x, k), m)  # same kernel, same order -> bit-identical
    xs = [g2d(rng.standard_normal((2, 256, 13, 13)).astype(np.float32)) for _ in range(6)]
    multi = ops().xcorr_depthwise_multi(xs, xs, circular=True)
    for x, m in zip(xs, multi):
        assert torch.equal(ops().xcorr_depthwise_circular(x, x), m)


def test_xcorr_unaligned_views_fall_back_correctly():
    rng = np.random.default_rng(12)
    buf = g2d(rng.standard_normal(2 * 256 * 29 * 29 + 1).astype(np.float32))
    x = buf[1:].reshape(2, 256, 29, 29)  # 4-byte aligned only -> generic kernel
    k = g2d(rng.standard_normal((2, 256, 5, 5)).astype(np.float32))
    ref = c_oracle.xcorr_dw(x.cpu().numpy(), k.cpu().numpy())
    assert_close(ops().xcorr_depthwise(x, k).cpu().numpy(), ref)


@pytest.mark.parametrize("shape", [(64, 256, 61, 61, 29, 29, 0), (64, 256, 29, 29, 29, 29, 1), (64, 256, 29, 29, 5, 5, 0),
                                   (64, 256, 13, 13, 13, 13, 1), (256, 256, 39, 39, 15, 15, 0)])
def test_xcorr_full_size_properties(shape, algo):
    """BASELINE.json sizes (batch 64 / 256): one-hot kernels make the correlation an exact shifted crop,
    and the operator is linear in x."""
    B, C, Hx, Wx, Hk, Wk, circ = shape
    fft = algo == "fft" and uses_fft(C, Hx, Wx, Hk, Wk, circ)
    fn = ops().xcorr_depthwise_circular if circ else ops().xcorr_depthwise
    gen = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn((B, C, Hx, Wx), device=DEV, generator=gen)
    u0, v0 = Hk // 3, Wk - 2
    k = torch.zeros((B, C, Hk, Wk), device=DEV)
    k[:, :, u0, v0] = 1.0
    out = fn(x, k)
    Ho, Wo = ops().xcorr_out_hw(Hx, Wx, Hk, Wk, circ)
    assert tuple(out.shape) == (B, C, Ho, Wo)
    if circ:
        rows = (torch.arange(Ho, device=DEV) + u0 - Hx // 2) % Hx
        cols = (torch.arange(Wo, device=DEV) + v0 - Wx // 2).clamp(0, Wx - 1)
        expect = x[:, :, rows][:, :, :, cols]
    else:
        expect = x[:, :, u0:u0 + Ho, v0:v0 + Wo]
    if fft:  # through the transform domain the crop is reproduced to fp32 rounding, not bit for bit
        assert float((out - expect).abs().max()) <= 1e-5 * float(x.abs().max())
    else:
        assert torch.equal(out, expect)  # 1.0 * x + 0 * ... is exact
    # linearity with a dense kernel
    k = torch.randn((B, C, Hk, Wk), device=DEV, generator=gen) * 0.1
    x2 = torch.randn((B, C, Hx, Wx), device=DEV, generator=gen)
    lhs = fn(2.0 * x + x2, k)
    rhs = 2.0 * fn(x, k) + fn(x2, k)
    assert torch.allclose(lhs, rhs, rtol=1e-3, atol=1e-4 * float(rhs.abs().max()))
    # shared template == tiled template
    assert torch.equal(fn(x, k[:1]), fn(x, k[:1].expand(B, C, Hk, Wk).contiguous()))
    if fft:  # the two algorithms agree far inside the parity tolerance
        ops().set_xcorr_algo("direct")
        direct = fn(x, k)
        ops().set_xcorr_algo("fft")
        assert float((fn(x, k) - direct).abs().max()) <= 5e-6 * float(direct.abs().max())


@pytest.mark.parametrize("shape", [(64, 256, 61, 61, 29, 29, 0), (64, 256, 29, 29, 29, 29, 1), (256, 256, 39, 39, 15, 15, 0)])
def test_xcorr_full_size_random_slice_vs_oracle(shape, algo):
    """BASELINE batch sizes against the C oracle itself: two randomly chosen pairs of the batch-64 / batch-256 output are
    recomputed on the CPU (the oracle finishes a 2-pair slice in seconds) -- not only self-consistency properties."""
    B, C, Hx, Wx, Hk, Wk, circ = shape
    fn = ops().xcorr_depthwise_circular if circ else ops().xcorr_depthwise
    gen = torch.Generator(device=DEV).manual_seed(B + Hx)
    x = torch.randn((B, C, Hx, Wx), device=DEV, generator=gen) + 0.5
    k = torch.randn((B, C, Hk, Wk), device=DEV, generator=gen) * 0.1
    out = fn(x, k)
    pick = np.random.default_rng(Hx * 7 + Hk).choice(B, 2, replace=False)
    for b in pick:
        ref = c_oracle.xcorr_dw(x[b:b + 1].cpu().numpy(), k[b:b + 1].cpu().numpy(), bool(circ))
        assert_close(out[b:b + 1].cpu().numpy(), ref, what="%s pair %d" % (shape, b))
        if algo == "fft" and uses_fft(C, Hx, Wx, Hk, Wk, circ):
            assert np.abs(out[b:b + 1].cpu().numpy() - ref).max() <= 5e-6 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(1, 4, 61, 61, 29, 29, 0), (1, 256, 61, 61, 29, 29, 0), (5, 256, 61, 61, 29, 29, 0), (3, 256, 29, 29, 29, 29, 1),
                                   (1, 8, 29, 29, 29, 29, 1), (7, 256, 39, 39, 15, 15, 0), (13, 64, 61, 61, 29, 29, 0), (64, 256, 61, 61, 29, 29, 0), (64, 256, 29, 29, 29, 29, 1)])
@pytest.mark.parametrize("variant", ["fft_pipe", "fft_ws", "fft"])
def test_fft_kernel_variants_equal_the_phased_kernel_bitwise(shape, variant):
    """The software-pipelined (O(g) next to R(g+1)) and the warp-specialised (roles + mbarriers, no CTA barrier) transform-domain
    kernels do the arithmetic of the three-phase kernel in the same order, so all agree bit for bit -- at one group, fewer groups than
    CTAs, one / several groups per CTA, odd counts, dense and shared templates, and through the multi-problem launch."""
    B, C, Hx, Wx, Hk, Wk, circ = shape
    fn = ops().xcorr_depthwise_circular if circ else ops().xcorr_depthwise
    gen = torch.Generator(device=DEV).manual_seed(B * 31 + C + Hx)
    x = torch.randn((B, C, Hx, Wx), device=DEV, generator=gen)
    k = torch.randn((B, C, Hk, Wk), device=DEV, generator=gen) * 0.1
    try:
        ops().set_xcorr_algo("fft_phased")
        want, want_shared = fn(x, k), fn(x, k[:1])
        want_multi = ops().xcorr_depthwise_multi([x, x * 0.5, x + 1.0], [k, k, k * 2.0], circular=bool(circ))
        ops().set_xcorr_algo(variant)
        for _ in range(3):  # repeated: a race between the rounds would not repeat itself
            assert torch.equal(fn(x, k), want)
        assert torch.equal(fn(x, k[:1]), want_shared)
        got_multi = ops().xcorr_depthwise_multi([x, x * 0.5, x + 1.0], [k, k, k * 2.0], circular=bool(circ))
        assert all(torch.equal(a, b) for a, b in zip(got_multi, want_multi))
        ops().set_xcorr_algo("direct")
        direct = fn(x, k)
        assert float((want - direct).abs().max()) <= 5e-6 * float(direct.abs().max())
    finally:
        ops().set_xcorr_algo("auto")


@pytest.mark.parametrize("shape", [(1, 256, 61, 61, 0), (5, 256, 61, 61, 0), (64, 256, 61, 61, 0), (1, 8, 29, 29, 1), (7, 256, 29, 29, 1), (64, 256, 29, 29, 1)])
@pytest.mark.parametrize("variant", ["auto", "fft_phased", "fft_pipe"])
def test_shared_template_with_cached_spectra_equals_the_per_pair_path(shape, variant):
    """A shared template's row spectra taken once (hdn_xcorr_template_spectra_f32) and consumed by hdn_xcorr_dw_multi_spec_f32 give the
    result of the per-pair kernels (which re-transform the template for every pair) to fp32 rounding, the C oracle's within the parity
    tolerance, for both kernel schedules, 3 problems per launch, repeated (no race between the spectra landing and the column stage)."""
    B, C, Hx, Wx, circ = shape
    Hk = Wk = 29
    fn = ops().xcorr_depthwise_circular if circ else ops().xcorr_depthwise
    gen = torch.Generator(device=DEV).manual_seed(B * 17 + C + Hx)
    xs = [torch.randn((B, C, Hx, Wx), device=DEV, generator=gen) + 0.3 * i for i in range(3)]
    ks = [torch.randn((1, C, Hk, Wk), device=DEV, generator=gen) * 0.1 for _ in range(3)]
    spectra = ops().xcorr_template_spectra(ks, Hx, Wx, bool(circ))
    assert spectra is not None and len(spectra) == 3
    try:
        ops().set_xcorr_algo("fft_phased")
        want = [fn(x, k) for x, k in zip(xs, ks)]
        ops().set_xcorr_algo(variant)
        first = None
        for _ in range(3):
            got = ops().xcorr_depthwise_multi_spec(xs, spectra, Hk, Wk, bool(circ))
            for g, w_ in zip(got, want):
                assert float((g - w_).abs().max()) <= 2e-6 * float(w_.abs().max())
            if first is None:
                first = [g.clone() for g in got]
            assert all(torch.equal(a, b) for a, b in zip(got, first))
    finally:
        ops().set_xcorr_algo("auto")
    b = B // 2
    ref = c_oracle.xcorr_dw(xs[1][b:b + 1].cpu().numpy(), ks[1].cpu().numpy(), bool(circ))
    assert_close(first[1][b:b + 1].cpu().numpy(), ref, what="cached spectra %s" % (shape,))
    assert np.abs(first[1][b:b + 1].cpu().numpy() - ref).max() <= 5e-6 * np.abs(ref).max()


def test_m1_engine_shared_template_takes_the_cached_spectra_path():
    """M1Engine.bind gives a batch-shared template its row spectra once; the chain's results equal the per-pair path's (arg-max indices
    bit-exact, maps to fp32 rounding) and a per-pair-template batch never takes that path."""
    from hdn_b200 import engine
    host = engine.make_inputs("256/512", 4, seed=3, shared_template=True)
    dev_in = {k: ([t.to(DEV) for t in v] if isinstance(v, list) else v.to(DEV)) for k, v in host.items()}
    eng = engine.M1Engine("256/512", 4, DEV, shared_template=True, use_graph=False)
    eng.bind(dev_in)
    assert "ks_spec" in eng.inp and "kl_spec" in eng.inp
    got = {k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in eng.run().items()}
    try:
        ops().TEMPLATE_SPECTRA = False
        eng.bind(dev_in)
        assert "ks_spec" not in eng.inp
        want = eng.run()
    finally:
        ops().TEMPLATE_SPECTRA = True
    for key in ("corr", "corr_lp"):
        for a, b in zip(got[key], want[key]):
            assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
    assert torch.equal(got["idx"], want["idx"]) and torch.equal(got["idx_lp"], want["idx_lp"]) and torch.equal(got["H"], want["H"])
    per_pair = engine.make_inputs("256/512", 4, seed=3, shared_template=False)
    eng.bind({k: ([t.to(DEV) for t in v] if isinstance(v, list) else v.to(DEV)) for k, v in per_pair.items()})
    assert "ks_spec" not in eng.inp


def test_template_spectra_only_where_a_kernel_exists():
    k = torch.randn((1, 256, 5, 5), device=DEV)
    assert ops().xcorr_template_spectra([k], 29, 29, False) is None            # native 127/255: direct kernels, nothing to cache
    assert ops().xcorr_template_spectra([torch.randn((1, 256, 15, 15), device=DEV)], 39, 39, False) is None
    with pytest.raises((ValueError, RuntimeError)):
        ops().xcorr_template_spectra([torch.randn((2, 256, 29, 29), device=DEV)], 61, 61, False)  # per-pair templates have no shared spectra


def test_xcorr_untiled_shape_runs_generic_kernel_and_is_counted():
    """A crop size outside the tiled table (INSTANCE_SIZE 287 -> 33x33 search features) still gives the reference's result,
    and the library counts it (hdn_xcorr_generic_launches) instead of degrading silently."""
    from hdn_b200 import _lib
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 32, 33, 33)).astype(np.float32)
    k = rng.standard_normal((2, 32, 5, 5)).astype(np.float32)
    assert _lib.lib().hdn_xcorr_is_staged(32, 33, 33, 5, 5, 0, 32 * 25) == 0
    n0 = _lib.lib().hdn_xcorr_generic_launches()
    out = ops().xcorr_depthwise(g2d(x), g2d(k)).cpu().numpy()
    assert _lib.lib().hdn_xcorr_generic_launches() == n0 + 1
    assert_close(out, c_oracle.xcorr_dw(x, k, False))


# ------------------------------------------------------------------ K3
@pytest.mark.parametrize("name", golden_names("ops_k3_"))
def test_logpolar_golden(name):
    g = load_golden(name)
    img = regen_image(g)
    S = int(g["inst"]) // 2
    stn = ops().STN_Polar(int(g["inst"]))
    out, grid = stn(g2d(img), g2d(g["polar"]), [0, float(g["delta"][1])])
    assert grid is None and tuple(out.shape) == g["out"].shape
    assert_close(out.cpu().numpy(), g["out"], what=name)
    assert_close(out.cpu().numpy(), c_oracle.logpolar(img, g["polar"], float(g["delta"][1]), S), what=name + " vs C")


def test_logpolar_batch_chunked_path_equals_per_item_path():
    """Large batches share the sampling position across 4 items per thread (or recompute it per item when `polar` is given):
    bit-identical to the one-item-per-thread path small batches take."""
    B, H, W, S = 24, 512, 512, 256
    gen = torch.Generator(device=DEV).manual_seed(9)
    img = torch.rand((B, 3, H, W), device=DEV, generator=gen) * 255.0
    for polar in (None, (torch.rand((B, 2), device=DEV, generator=gen) - 0.5) * 40.0):
        big = ops().logpolar_sample(img, polar, 0.25, S)
        small = torch.cat([ops().logpolar_sample(img[b:b + 2], None if polar is None else polar[b:b + 2], 0.25, S) for b in range(0, B, 2)])
        assert torch.equal(big, small)
    ref = c_oracle.logpolar(img[:2].cpu().numpy(), polar[:2].cpu().numpy(), 0.25, S)
    assert_close(big[:2].cpu().numpy(), ref, what="chunked log-polar vs C oracle")


def test_logpolar_full_size_affine_reproduction():
    """[64,3,512,512] -> [64,3,256,256]: bilinear sampling reproduces an affine image exactly at the sample point."""
    B, H, W, S = 64, 512, 512, 256
    yy, xx = torch.meshgrid(torch.arange(H, device=DEV, dtype=torch.float32), torch.arange(W, device=DEV, dtype=torch.float32), indexing="ij")
    img = torch.stack([0.5 * xx + 0.25 * yy, xx, yy]).unsqueeze(0).expand(B, 3, H, W).contiguous()
    out = ops().logpolar_sample(img, None, 0.0, S)
    assert tuple(out.shape) == (B, 3, S, S)
    j = torch.arange(S, dtype=torch.float64)
    rho = torch.exp(np.log(S / 2) / S * j) - 1
    th = torch.arange(S, dtype=torch.float64) * 2 * np.pi / S
    ix = (((rho[None] * torch.cos(th)[:, None]) / (H // 2) + 1) * W - 1) / 2
    iy = (((rho[None] * torch.sin(th)[:, None]) / (W // 2) + 1) * H - 1) / 2
    ix, iy = ix.clamp(0, W - 1), iy.clamp(0, H - 1)
    got = out[7].double().cpu()
    assert float((got[1] - ix).abs().max()) < 2e-3
    assert float((got[2] - iy).abs().max()) < 2e-3
    assert float((got[0] - (0.5 * ix + 0.25 * iy)).abs().max()) < 2e-3
    assert torch.equal(out[0], out[63])
    # grid option reproduces the reference's second return value shape
    stn = ops().STN_Polar(512, return_grid=True)
    _, grid = stn(img[:2], torch.zeros(2, 2, device=DEV))
    assert tuple(grid.shape) == (2, S, S, 2)


# ------------------------------------------------------------------ K5
def test_dlt_golden_and_property():
    from test_oracle_golden import reproject
    g = load_golden("ops_k5_dlt")
    H = ops().DLT_solve(g2d(g["src"]), g2d(g["off"]))
    assert tuple(H.shape) == (16, 1, 3, 3)
    got, ref = H[:, 0].cpu().numpy(), g["H"][:, 0]
    assert np.allclose(got, ref, rtol=1e-3, atol=1e-5)
    assert np.max(np.abs(reproject(got, g["src"]) - reproject(ref, g["src"]))) <= 1e-3 * 127
    assert np.array_equal(got, c_oracle.dlt4(g["src"], g["off"]))  # same fp64 elimination order -> bit-identical
    # full batch: H maps every source corner onto src + off
    rng = np.random.default_rng(3)
    B = 4096 + 3
    src = np.tile(np.asarray([0, 0, 0, 127, 127, 127, 127, 0], np.float32), (B, 1))
    off = rng.uniform(-8, 8, (B, 8)).astype(np.float32)
    got = ops().DLT_solve(g2d(src), g2d(off))[:, 0].cpu().numpy()
    assert np.max(np.abs(reproject(got, src) - (src + off).reshape(-1, 4, 2))) < 2e-3
    assert np.all(got[:, 2, 2] == 1.0)


# ------------------------------------------------------------------ K4
@pytest.mark.parametrize("name", ["ops_k4_warp", "ops_k4_warp_small"])
def test_homo_warp_golden(name):
    from test_oracle_golden import _warp_mismatch_ok
    g = load_golden(name)
    M, Minv = (g["M"], g["M_inv"]) if "M" in g else c_oracle.default_M(127, 127)
    out = ops().homo_warp(g2d(g["img"]), g2d(g["H"]), list(np.asarray(M).reshape(9)), list(np.asarray(Minv).reshape(9))).cpu().numpy()
    _warp_mismatch_ok(out, g["out"], name)
    ref = c_oracle.homo_warp(g["img"], g["H"], M, Minv)
    assert np.mean(out != ref) < 1e-3, "GPU and C oracle follow the same fp32 operation order"
    # the reference-signature entry point, with explicit (identity) patch indices -> general gather path
    B, Ch, H, W = g["img"].shape
    Mt = g2d(np.asarray(M, np.float32)).unsqueeze(0).expand(B, 3, 3)
    Mit = g2d(np.asarray(Minv, np.float32)).unsqueeze(0).expand(B, 3, 3)
    pidx = torch.arange(H * W, device=DEV, dtype=torch.float32).unsqueeze(0).expand(B, H * W)
    bidx = (torch.arange(B, device=DEV) * H * W).unsqueeze(1).expand(B, H * W).reshape(-1)
    via = ops().transform(H, W, Mit, g2d(g["H"]), Mt, g2d(g["img"]), pidx, bidx)
    fast = ops().transform(H, W, Mit, g2d(g["H"]), Mt, g2d(g["img"]), None, None)
    assert torch.equal(via, fast) and np.array_equal(fast.cpu().numpy(), out)


def test_dlt_warp_fused_equals_two_step():
    rng = np.random.default_rng(9)
    B = 64
    src = g2d(np.tile(np.asarray([0, 0, 0, 127, 127, 127, 127, 0], np.float32), (B, 1)))
    off = g2d(rng.uniform(-8, 8, (B, 8)).astype(np.float32))
    img = g2d(rng.standard_normal((B, 1, 127, 127)).astype(np.float32))
    H1 = ops().DLT_solve(src, off)[:, 0]
    w1 = ops().homo_warp(img, H1)
    H2, w2 = ops().dlt_warp(src, off, img)
    assert torch.equal(H1, H2) and torch.equal(w1, w2)
    # zero offsets: identity H is a stretch by W/(W-1), last row / col ~ 0
    H0, w0 = ops().dlt_warp(src[:1], torch.zeros_like(off[:1]), img[:1])
    assert float(w0[0, 0, -1, :].abs().max()) < 1e-6 and float(w0[0, 0, :, -1].abs().max()) < 1e-6


# ------------------------------------------------------------------ K6
def test_score_argmax_golden():
    g = load_golden("ops_k6_score")
    idx, ps, sc, gath = ops().score_argmax(g2d(g["cls"]), g2d(g["loc"]), g2d(g["window"]), float(g["win_infl"]))
    assert idx.dtype == torch.int64 and ps.dtype == torch.float64
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    assert np.allclose(ps.cpu().numpy(), g["pscore"], rtol=1e-6, atol=1e-7)
    assert np.allclose(sc.cpu().numpy(), g["score"], rtol=1e-6, atol=1e-7)
    centre = g["points"][g["idx"]] - 8.0 * gath.cpu().numpy()
    assert np.allclose(centre, g["center"], rtol=1e-6, atol=1e-5)
    g = load_golden("ops_k6_score_lp")
    idx, ps, sc, gath = ops().score_argmax(g2d(g["cls"]), g2d(g["loc"]), None, 0.0)
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    assert np.allclose(sc.cpu().numpy(), g["score"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("N,L,windowed", [(25, 2, True), (13, 4, False), (33, 2, True), (40, 3, True)])
def test_score_argmax_vs_oracle_batch512(N, L, windowed):
    rng = np.random.default_rng(N)
    B = 512
    cls = (rng.standard_normal((B, 2, N, N)) * 2).astype(np.float32)
    loc = rng.standard_normal((B, L, N, N)).astype(np.float32)
    cls[5] = 0.25  # exact ties everywhere
    win = np.outer(np.hanning(N), np.hanning(N)).flatten() if windowed else None
    w = 0.1632532824922313 if windowed else 0.0
    ridx, rps, rsc, rg = c_oracle.score_argmax(cls, loc, win, w)
    idx, ps, sc, gath = ops().score_argmax(g2d(cls), g2d(loc), g2d(win) if windowed else None, w)
    assert np.array_equal(idx.cpu().numpy(), ridx)  # bit-exact index
    assert np.array_equal(gath.cpu().numpy(), rg)
    assert np.allclose(ps.cpu().numpy(), rps, rtol=1e-6, atol=1e-7)
    assert np.allclose(sc.cpu().numpy(), rsc, rtol=1e-6, atol=1e-7)
    if not windowed:
        assert int(idx[5]) == 0  # first maximum wins a full tie


def test_score_argmax_nan_scores_follow_numpy():
    """Non-finite scores (a NaN frame): np.argmax returns the first NaN; the kernel must do the same and never gather out of range."""
    rng = np.random.default_rng(9)
    N, L, B = 25, 2, 4
    cls = (rng.standard_normal((B, 2, N, N)) * 2).astype(np.float32)
    loc = rng.standard_normal((B, L, N, N)).astype(np.float32)
    cls[0] = np.nan                  # every score NaN -> index 0
    cls[1, 1, 3, 7] = np.nan         # a single NaN -> that cell
    cls[2, 0, 20, 2] = np.nan
    cls[2, 1, 4, 4] = np.nan         # two NaNs -> the first in scan order
    win = np.outer(np.hanning(N), np.hanning(N)).flatten()
    w = 0.1632532824922313
    idx, ps, sc, gath = ops().score_argmax(g2d(cls), g2d(loc), g2d(win), w)
    e = np.exp(cls - cls.max(1, keepdims=True))
    score = (e[:, 1] / e.sum(1)).reshape(B, -1).astype(np.float32)
    pscore = score * np.float32(1 - w) + win * w   # hdn_tracker_proj_e2e.py:172-173
    expect = np.argmax(pscore, 1)
    assert list(expect[:3]) == [0, 3 * N + 7, 4 * N + 4]
    assert np.array_equal(idx.cpu().numpy(), expect)
    assert np.array_equal(gath.cpu().numpy(), np.stack([loc[b].reshape(L, -1)[:, expect[b]] for b in range(B)]))
    ridx = c_oracle.score_argmax(cls, loc, win, w)[0]
    assert np.array_equal(ridx, expect)


# ------------------------------------------------------------------ device-side frame pre-processing
def _frame(rng, h, w):
    from hdn_b200 import synthetic
    img = synthetic.texture(int(rng.integers(1 << 30)), h, w)
    return np.ascontiguousarray(np.clip(img.astype(np.int32) + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8))


def test_device_warp_perspective_is_bit_exact_with_opencv():
    import cv2
    from hdn_b200.preproc import FramePreproc
    from oracle import cv_port
    rng = np.random.default_rng(21)
    pre = FramePreproc(DEV)
    for t, (h, w) in enumerate([(360, 480), (720, 1280), (97, 131), (15, 40), (804, 1920)]):
        img = _frame(rng, h, w)
        src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32)
        Hm = cv2.getPerspectiveTransform(src, src + rng.normal(0, 12 + 10 * t, (4, 2)).astype(np.float32))
        M = np.linalg.inv(Hm)
        pre.upload(img)
        got = pre.warp_perspective(M).cpu().numpy()
        assert np.array_equal(got, cv2.warpPerspective(img, M, (w, h), borderMode=cv2.BORDER_REPLICATE)), (h, w)
        if h * w < 200000:
            assert np.array_equal(got, cv_port.warp_perspective_u8(img, M))
    img = _frame(rng, 120, 160)
    pre.upload(img)
    assert np.array_equal(pre.warp_perspective(np.eye(3)).cpu().numpy(), img)  # identity reproduces the frame


def test_device_rotation_is_bit_exact_with_opencv():
    import cv2
    from hdn_b200.preproc import FramePreproc
    rng = np.random.default_rng(22)
    pre = FramePreproc(DEV)
    for h, w in [(360, 480), (720, 1280), (97, 131)]:
        img = _frame(rng, h, w)
        for rot in (0.0, 0.013, -0.4, 1.3):
            cx, cy = float(rng.uniform(0, w)), float(rng.uniform(0, h))
            cc, ss = np.cos(rot), np.sin(rot)
            M = np.array([[cc, -ss, cx - cx * cc + cy * ss], [ss, cc, cy - cy * cc - cx * ss]])
            ref = cv2.warpAffine(img, M, (w, h), flags=2, borderMode=cv2.BORDER_REPLICATE)
            got = pre.rotate(pre.upload(img), cx, cy, rot).cpu().numpy()
            assert np.array_equal(got, ref), (h, w, rot)


def test_device_crops_are_bit_exact_with_the_trackers_crop_window():
    """300 random windows (inside, across every border, fully outside, exact 2x, no resize) + the gray normalisation."""
    from hdn_b200 import compat
    compat.activate()
    from hdn.tracker.base_tracker import crop_window
    from hdn_b200.preproc import FramePreproc
    from oracle import cv_port
    rng = np.random.default_rng(23)
    img = _frame(rng, 360, 480)
    avg = np.mean(img, axis=(0, 1))
    pre = FramePreproc(DEV)
    dev = pre.upload(img)
    cases = [([240.3, 180.9], 127, 181.0), ([20.0, 340.5], 255, 363.0), ([200.0, 200.0], 127, 127.0), ([230.0, 170.0], 255, 510.0),
             ([-300.0, -300.0], 127, 90.0), ([100.0, 100.0], 127, 254.0)]
    cases += [([float(rng.uniform(-60, 540)), float(rng.uniform(-60, 420))], int(rng.choice([127, 255])), float(np.floor(rng.uniform(40, 700))))
              for _ in range(294)]
    for pos, msz, osz in cases:
        ref, _ = crop_window(img, np.array(pos), msz, osz, avg)
        got = pre.crop(dev, pos, msz, osz, avg).cpu().numpy()
        assert np.array_equal(got, ref), (pos, msz, osz)
    for pos, msz, osz in cases[:40]:
        ref = cv_port.gray_normalise(crop_window(img, np.array(pos), 127, osz, avg)[0]).astype(np.float32)
        got = pre.crop(dev, pos, 127, osz, avg, gray=True).cpu().numpy()
        assert np.array_equal(got, ref), (pos, osz)
