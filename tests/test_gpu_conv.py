"""GPU: the tcgen05 3xTF32 implicit-GEMM convolution (hdn_conv_gemm_f32) against torch fp64 / cuDNN fp32.

The kernel must be fp32-ACCURATE (that is the point of the hi/lo operand split and of the chunked TMEM accumulation):
its error against an fp64 reference has to stay within a small factor of cuDNN's own fp32 error, far below plain TF32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # B, Cin, Cout, H, W, k, dilation, valid
    (1, 64, 128, 9, 9, 1, 1, False),          # smallest legal K (2 blocks)
    (1, 512, 512, 31, 31, 3, 4, False),       # layer4 conv2: the layer cuDNN runs at 1.7 TFLOP/s
    (1, 1024, 2048, 31, 31, 3, 2, False),     # layer4 projection shortcut
    (1, 1024, 256, 31, 31, 1, 1, False),      # bottleneck conv1
    (2, 256, 256, 15, 15, 3, 2, False),       # template-side layer3 block
    (1, 256, 256, 31, 31, 3, 1, True),        # head conv_search: 3x3 VALID -> 29x29
    (3, 256, 256, 7, 7, 3, 1, True),          # head conv_kernel: 7x7 -> 5x5
    (1, 128, 512, 63, 63, 1, 1, False),       # 3969 pixels: many pixel tiles, ragged last tile
    (40, 512, 256, 31, 31, 1, 1, False),      # large batch -> the 128-pixel tile configuration
]


@pytest.mark.parametrize("case", CASES)
def test_conv_gemm_is_fp32_accurate(case):
    from hdn_b200 import ops
    B, Cin, Cout, H, W, k, d, valid = case
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(Cin * 7 + Cout + H)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = 1 + 0.1 * torch.randn(Cout, device="cuda", generator=g)
    shift = 0.1 * torch.randn(Cout, device="cuda", generator=g)
    pad = 0 if valid else d * (k // 2)

    def ref(dt):
        y = F.conv2d(x.to(dt), w.to(dt), padding=pad, dilation=d) * scale.to(dt).view(1, -1, 1, 1) + shift.to(dt).view(1, -1, 1, 1)
        return y

    y64 = ref(torch.float64)
    res = torch.randn(y64.shape, device="cuda", generator=g)
    want = F.relu(y64 + res.double())
    got = ops.conv_gemm(x, ops.pack_conv_weight(w), scale, shift, res, ksize=k, dilation=d, relu=True, valid=valid)
    assert tuple(got.shape) == tuple(want.shape)
    den = float(want.abs().max())
    err = float((got.double() - want).abs().max()) / den
    err_cudnn = float((F.relu(ref(torch.float32) + res).double() - want).abs().max()) / den
    assert err < 1e-5 and err < 8 * max(err_cudnn, 3e-7), (err, err_cudnn)
    # options off: plain convolution
    plain = ops.conv_gemm(x, ops.pack_conv_weight(w), ksize=k, dilation=d, valid=valid)
    want_plain = F.conv2d(x.double(), w.double(), padding=pad, dilation=d)
    assert float((plain.double() - want_plain).abs().max()) / float(want_plain.abs().max()) < 1e-5


def test_conv_gemm_rejects_unsupported_shapes():
    from hdn_b200 import ops
    assert ops.conv_gemm_supported(256, 256, 3, 2) and ops.conv_gemm_supported(64, 128, 1)
    assert not ops.conv_gemm_supported(3, 64, 7) and not ops.conv_gemm_supported(256, 2, 1) and not ops.conv_gemm_supported(48, 128, 1)
    x = torch.randn(1, 48, 8, 8, device="cuda")
    with pytest.raises((ValueError, RuntimeError)):
        ops.conv_gemm(x, torch.zeros(2 * 128 * 48, device="cuda"), ksize=1)


def test_fused_block_equals_module_path():
    """conv_bn_act (one tensor-core launch) == bn(conv(x)) (+ residual) (ReLU) of the plain modules."""
    from hdn_b200 import convs
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(256, 512, 3, padding=2, dilation=2, bias=False).cuda()
    bn = torch.nn.BatchNorm2d(512).cuda().eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        x = torch.randn(2, 256, 15, 15, device="cuda")
        r = torch.randn(2, 512, 15, 15, device="cuda")
        fused = convs.conv_bn_act(conv, bn, x, residual=r, relu=True)
        plain = torch.relu(bn(conv(x)) + r)
    assert torch.allclose(fused, plain, rtol=1e-4, atol=1e-5 * float(plain.abs().max()))


def test_conv_gemm_multi_equals_single_launches():
    """6 same-shape problems in one launch (blockIdx.z = problem x image) == 6 single launches, bit for bit."""
    from hdn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = [torch.randn(2, 256, 17, 17, device="cuda", generator=g) for _ in range(3)]
    ws = [torch.randn(256, 256, 3, 3, device="cuda", generator=g) * 0.03 for _ in range(6)]
    sc = [1 + 0.1 * torch.randn(256, device="cuda", generator=g) for _ in range(6)]
    sh = [0.1 * torch.randn(256, device="cuda", generator=g) for _ in range(6)]
    packs = [ops.pack_conv_weight(w) for w in ws]
    ops.set_conv_splitk(False)  # (a lone small launch would split K over a cluster and sum in another order: same result to fp32 rounding only)
    try:
        multi = ops.conv_gemm_multi([xs[i // 2] for i in range(6)], packs, sc, sh, ksize=3, relu=True, valid=True)
        for i in range(6):
            single = ops.conv_gemm(xs[i // 2], packs[i], sc[i], sh[i], ksize=3, relu=True, valid=True)
            assert tuple(single.shape) == (2, 256, 15, 15) and torch.equal(single, multi[i])
    finally:
        ops.set_conv_splitk(True)
    split = ops.conv_gemm(xs[0], packs[0], sc[0], sh[0], ksize=3, relu=True, valid=True)
    assert float((split - multi[0]).abs().max()) <= 5e-6 * float(multi[0].abs().max())


@pytest.mark.parametrize("B,N,L", [(1, 25, 2), (3, 33, 2), (2, 13, 4), (2, 29, 4)])
def test_fused_head_tail_equals_reference_ops(B, N, L):
    """hdn_head_project_multi_f32 + hdn_head_score_f32 == the reference's tail of MultiBAN.forward (ban.py:62-66, 102-127):
    per level 1x1 + BN + ReLU + 1x1 (+bias), loc * loc_scale, softmax-weighted level sums, then softmax / window / arg-max."""
    from hdn_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 100 + N + L)
    C, nlev = 256, 3
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    feats = [rnd(B, C, N, N) for _ in range(2 * nlev)]                     # cls2, loc2, cls3, loc3, ...
    w1 = [rnd(C, C, 1, 1) * (2.0 / C) ** 0.5 for _ in range(2 * nlev)]
    sc = [1 + 0.1 * rnd(C) for _ in range(2 * nlev)]
    sh = [0.1 * rnd(C) for _ in range(2 * nlev)]
    w2 = [rnd(2 if i % 2 == 0 else L, C) * 0.1 for i in range(2 * nlev)]
    b2 = [rnd(2 if i % 2 == 0 else L) * 0.1 for i in range(2 * nlev)]
    cls_w = torch.softmax(rnd(nlev), 0).tolist()
    loc_w = torch.softmax(rnd(nlev), 0).tolist()
    loc_scale = (1 + 0.2 * rnd(nlev)).tolist()
    packs = [ops.pack_conv_weight(w) for w in w1]
    cls_parts = ops.head_project_multi(feats[0::2], packs[0::2], sc[0::2], sh[0::2], w2[0::2])
    loc_parts = ops.head_project_multi(feats[1::2], packs[1::2], sc[1::2], sh[1::2], w2[1::2])
    assert tuple(cls_parts[0].shape) == (2, B, 2, N * N) and tuple(loc_parts[0].shape) == (2, B, L, N * N)
    win = torch.from_numpy(np.outer(np.hanning(N), np.hanning(N)).flatten()).cuda()
    w_infl = 0.1632532824922313
    cls, loc, buf = ops.head_score(cls_parts, loc_parts, b2[0::2], b2[1::2], cls_w, loc_scale, loc_w, N, win, w_infl)
    idx, ps, scr, gath = ops.unpack_scores(buf.cpu().numpy(), B, L)

    def level(i, dt):
        h = F.relu(F.conv2d(feats[i].to(dt), w1[i].to(dt)) * sc[i].to(dt).view(1, -1, 1, 1) + sh[i].to(dt).view(1, -1, 1, 1))
        return F.conv2d(h, w2[i].to(dt).view(-1, C, 1, 1), b2[i].to(dt))

    def combined(dt):
        c = sum(level(2 * l, dt) * cls_w[l] for l in range(nlev))
        o = sum(level(2 * l + 1, dt) * loc_scale[l] * loc_w[l] for l in range(nlev))
        return c, o

    c64, o64 = combined(torch.float64)
    c32, o32 = combined(torch.float32)
    for got, want, lib in ((cls, c64, c32), (loc, o64, o32)):
        den = float(want.abs().max())
        err = float((got.double() - want).abs().max()) / den
        err_lib = float((lib.double() - want).abs().max()) / den
        assert err < 1e-5 and err < 8 * max(err_lib, 3e-7), (err, err_lib)
    # the fused K6 on the combined maps == the stand-alone K6 kernel on the same maps (bit for bit)
    idx2, ps2, sc2, g2 = ops.score_argmax(cls, loc, win, w_infl)
    assert np.array_equal(idx, idx2.cpu().numpy()) and np.array_equal(ps, ps2.cpu().numpy()) and np.array_equal(scr, sc2.cpu().numpy())
    assert np.array_equal(gath, g2.cpu().numpy())
    # maps not requested: same scores
    _, _, buf2 = ops.head_score(cls_parts, loc_parts, b2[0::2], b2[1::2], cls_w, loc_scale, loc_w, N, win, w_infl, want_maps=False)
    assert torch.equal(buf[:20 * B + 4 * B * L], buf2[:20 * B + 4 * B * L])  # (the buffer's tail is alignment padding)


@pytest.mark.parametrize("workload,B,chunk,shared,u8", [("127/255", 5, 2, False, False), ("127/255", 3, 4, True, True), ("256/512", 2, 1, False, True), ("256/512", 3, 2, True, True)])
def test_head_engine_equals_cpu_port(workload, B, chunk, shared, u8):
    """The fused chain from neck features (HeadEngine: conv_search -> correlation -> 1x1 tail -> level sum + K6, K3, K5 + K4),
    device-resident and through the pinned-host pipeline, against the CPU port of the reference's own torch calls
    (oracle/torch_port.fused_chain).  Maps within 1e-3; arg-max indices bit-exact."""
    from hdn_b200 import head_engine as he
    from hdn_b200.engine import WIN_INFL
    from oracle import c_oracle, torch_port as tp
    dev = torch.device("cuda")
    host = he.make_inputs(workload, B, seed=5, shared_template=shared, u8_crop=u8)
    eng = he.HeadEngine(workload, B, dev, chunk=chunk)
    up = lambda v: [t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)  # noqa: E731
    eng.set_template(up(host["zf"]), up(host["zf_lp"]))
    eng.bind({k: up(host[k]) for k in he.FRAME_KEYS})
    out = {k: v.cpu() for k, v in eng.run().items()}
    w_sim, w_lp = ({k: ([t.cpu() for t in v] if isinstance(v, list) and isinstance(v[0], torch.Tensor) else v) for k, v in w.raw.items()}
                   for w in (eng.w_sim, eng.w_lp))
    M, Mi = c_oracle.default_M(127, 127)
    feats = dict(host, S=he.NECK[workload]["S"], M=torch.from_numpy(M), Minv=torch.from_numpy(Mi), img=host["img"].float())
    win = np.outer(np.hanning(eng.N), np.hanning(eng.N)).flatten()
    ref = tp.fused_chain(feats, w_sim, w_lp, win, WIN_INFL)
    for k in ("cls", "loc", "cls_lp", "loc_lp", "x_lp"):
        r = ref[k].numpy()
        atol = 1e-4 * np.abs(r).max()
        assert np.all(np.abs(out[k].numpy() - r) <= atol + 1e-3 * np.abs(r)), k
    assert np.array_equal(out["idx"].numpy(), ref["idx"]) and np.array_equal(out["idx_lp"].numpy(), ref["idx_lp"])
    assert np.allclose(out["center"].numpy(), ref["center"], rtol=1e-3, atol=1e-4) and np.allclose(out["sim_lp"].numpy(), ref["sim_lp"], rtol=1e-3, atol=1e-4)
    assert np.allclose(out["H"].numpy(), ref["H"].numpy(), rtol=1e-3, atol=1e-5)
    # end to end through pinned host buffers: same results as the device-resident run
    pinned = he.make_inputs(workload, B, seed=5, shared_template=shared, pin=True, u8_crop=u8)
    h2d, d2h = eng.alloc_host_io(pinned)
    nb = lambda t: t.numel() * t.element_size()  # noqa: E731
    assert h2d == sum(sum(nb(t) for t in pinned[k]) if isinstance(pinned[k], list) else nb(pinned[k]) for k in he.FRAME_KEYS)
    res = eng.run_host(pinned)
    torch.cuda.synchronize()
    for k in he.HeadEngine.HOST_OUT:
        assert torch.equal(res[k], out[k]), k
    # a stream of batches (wait=False): steps overlap, result sets alternate, every set equals the synchronous result
    sets = [eng.run_host(pinned, wait=False) for _ in range(3)]
    eng.finish()
    torch.cuda.synchronize()
    assert sets[0] is sets[2] and sets[0] is not sets[1]
    for k in he.HeadEngine.HOST_OUT:
        assert torch.equal(sets[1][k], out[k]) and torch.equal(sets[2][k], out[k]), k


@pytest.mark.parametrize("case", [(1, 256, 256, 15, 15, 3, 2), (1, 1024, 256, 31, 31, 1, 1), (1, 512, 512, 31, 31, 3, 4), (2, 256, 1024, 15, 15, 1, 1)])
def test_conv_gemm_cluster_split_k(case):
    """Tracking-batch-size layers split K over a thread-block cluster (partial tiles added through distributed shared memory in
    rank order): same fp32 accuracy as the single-CTA kernel, deterministic (bitwise repeatable), residual / ReLU epilogue intact."""
    from hdn_b200 import ops
    B, Cin, Cout, H, W, k, d = case
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout + H + k)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = 1 + 0.1 * torch.randn(Cout, device="cuda", generator=g)
    shift = 0.1 * torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(B, Cout, H, W, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    try:
        ops.set_conv_splitk(False)
        one = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True)
        ops.set_conv_splitk(True)
        a = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True)
        b = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True)
    finally:
        ops.set_conv_splitk(True)
    assert torch.equal(a, b)
    want = F.relu(F.conv2d(x.double(), w.double(), padding=d * (k // 2), dilation=d) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
                  + res.double())
    den = float(want.abs().max())
    assert float((a.double() - want).abs().max()) / den < 1e-5 and float((one.double() - want).abs().max()) / den < 1e-5
    assert float((a - one).abs().max()) / den < 5e-6


@pytest.mark.parametrize("case", [(2, 256, 256, 63, 63), (3, 256, 256, 31, 31), (1, 64, 128, 9, 12), (2, 256, 256, 7, 7), (1, 512, 256, 15, 15), (5, 32, 128, 3, 3)])
def test_conv3x3_valid_shifted_window_kernel(case):
    """conv_shift.cu (activations staged once per channel block, the nine taps = shifted windows of one shared-memory tile):
    fp32-accurate against fp64, equal to the generic implicit GEMM up to summation order, multi-problem launches, BN + residual + ReLU."""
    from hdn_b200 import ops
    B, Cin, Cout, H, W = case
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout + H * 7 + W)
    xs = [torch.randn(B, Cin, H, W, device="cuda", generator=g) + 0.3 for _ in range(2)]
    ws = [torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (Cin * 9)) ** 0.5 for _ in range(3)]
    sc = [1 + 0.1 * torch.randn(Cout, device="cuda", generator=g) for _ in range(3)]
    sh = [0.1 * torch.randn(Cout, device="cuda", generator=g) for _ in range(3)]
    packs = [ops.pack_conv_weight(w) for w in ws]
    res = torch.randn(B, Cout, H - 2, W - 2, device="cuda", generator=g)
    try:
        ops.set_conv_shift(False)
        old = ops.conv_gemm(xs[0], packs[0], sc[0], sh[0], res, ksize=3, relu=True, valid=True)
        ops.set_conv_shift(True)
        new = ops.conv_gemm(xs[0], packs[0], sc[0], sh[0], res, ksize=3, relu=True, valid=True)
        again = ops.conv_gemm(xs[0], packs[0], sc[0], sh[0], res, ksize=3, relu=True, valid=True)
        multi = ops.conv_gemm_multi([xs[0], xs[1], xs[1]], packs, sc, sh, ksize=3, relu=False, valid=True)
    finally:
        ops.set_conv_shift(True)
    assert torch.equal(new, again)
    want = F.relu(F.conv2d(xs[0].double(), ws[0].double()) * sc[0].double().view(1, -1, 1, 1) + sh[0].double().view(1, -1, 1, 1) + res.double())
    den = float(want.abs().max())
    err_new, err_old = float((new.double() - want).abs().max()) / den, float((old.double() - want).abs().max()) / den
    assert err_new < 1e-5 and err_new < 4 * max(err_old, 3e-7), (err_new, err_old)
    for i, xi in enumerate([0, 1, 1]):
        w64 = F.conv2d(xs[xi].double(), ws[i].double()) * sc[i].double().view(1, -1, 1, 1) + sh[i].double().view(1, -1, 1, 1)
        assert float((multi[i].double() - w64).abs().max()) / float(w64.abs().max()) < 1e-5, i


STRIDED = [  # B, Cin, Cout, H, W, k, stride, pad, dil   -- the remaining layer geometries of the two ResNets
    (2, 64, 64, 63, 63, 3, 1, 1, 1),      # ResNet-50 layer1 conv2 / ResNet-34 layer1: 64-wide, zero-padded 128-row tile
    (2, 64, 256, 63, 63, 1, 1, 0, 1),     # layer1 conv3 / 1x1 projection
    (1, 256, 64, 63, 63, 1, 1, 0, 1),     # layer1 conv1 of the later blocks
    (2, 128, 128, 63, 63, 3, 2, 0, 1),    # ResNet-50 layer2 block 0: 3x3 stride 2, padding 0 (resnet_atrous.py:68-80)
    (1, 256, 512, 63, 63, 3, 2, 0, 1),    # layer2 projection shortcut: 3x3 stride 2 padding 0
    (3, 64, 128, 32, 32, 3, 2, 1, 1),     # ResNet-34 layer2 block 0: 3x3 stride 2 padding 1
    (3, 64, 128, 32, 32, 1, 2, 0, 1),     # ResNet-34 1x1 stride-2 projection
    (2, 256, 512, 8, 8, 3, 2, 1, 1),      # ResNet-34 layer4 block 0 (tiny maps: split-K)
    (1, 128, 192, 17, 23, 3, 1, 1, 1),    # Cout = 192 (64-multiple, two tiles, the second half empty), non-square
]


@pytest.mark.parametrize("case", STRIDED)
def test_conv_gemm_strides_paddings_and_64_wide_layers(case):
    from hdn_b200 import ops
    B, Cin, Cout, H, W, k, s, p, d = case
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout * 3 + H + k + s)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g) + 0.2
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = 1 + 0.1 * torch.randn(Cout, device="cuda", generator=g)
    shift = 0.1 * torch.randn(Cout, device="cuda", generator=g)
    want = F.conv2d(x.double(), w.double(), stride=s, padding=p, dilation=d) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    res = torch.randn(want.shape, device="cuda", generator=g)
    want = F.relu(want + res.double())
    got = ops.conv_gemm(x, ops.pack_conv_weight(w), scale, shift, res, ksize=k, dilation=d, relu=True, stride=s, padding=p, cout=Cout)
    assert tuple(got.shape) == tuple(want.shape)
    den = float(want.abs().max())
    err = float((got.double() - want).abs().max()) / den
    lib = F.relu(F.conv2d(x, w, stride=s, padding=p, dilation=d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    err_lib = float((lib.double() - want).abs().max()) / den
    assert err < 1e-5 and err < 8 * max(err_lib, 3e-7), (err, err_lib)


def test_every_convolution_of_the_path_runs_on_our_kernels():
    """tcgen05 for every 1x1 / 3x3 layer of the two ResNets, the direct few-channel kernel for the 7x7 stems and ShareFeature: no cuDNN
    convolution is left on the path."""
    from hdn_b200 import compat, convs
    compat.activate()
    from hdn.models.backbone.resnet_atrous import resnet50
    from homo_estimator.Deep_homography.Oneline_DLTv1.backbone.resnet import resnet34
    from homo_estimator.Deep_homography.Oneline_DLTv1.preprocess.input_feature_extractor import PreShareFeature
    x = torch.zeros(1, 64, 8, 8, device="cuda")
    with torch.no_grad():
        for net in (resnet50(used_layers=[2, 3, 4]), resnet34(used_layers=[4])):
            left = [n for n, m in net.named_modules() if isinstance(m, torch.nn.Conv2d) and not convs.tensor_core_eligible(m, x)]
            assert left == ["conv1"], left
            assert convs.small_eligible(net.conv1, x)
        assert all(convs.small_eligible(m, x) for m in PreShareFeature().modules() if isinstance(m, torch.nn.Conv2d))


@pytest.mark.parametrize("cin,cout,k,s,p,hw", [(3, 64, 7, 2, 0, (255, 255)), (3, 64, 7, 2, 0, (127, 127)), (2, 64, 7, 2, 3, (127, 127)),
                                                (1, 4, 3, 1, 1, (127, 127)), (4, 8, 3, 1, 1, (127, 127)), (8, 1, 3, 1, 1, (127, 127)),
                                                (3, 32, 7, 2, 3, (40, 77)), (5, 12, 3, 1, 0, (9, 35)), (8, 8, 3, 1, 1, (1, 1))])
def test_conv_small_matches_torch(cin, cout, k, s, p, hw):
    """hdn_conv_small_f32 (stems, ShareFeature) against torch's fp32 convolution + BatchNorm affine + ReLU; the sums have at most
    8 * 49 terms, so the two orders agree to a few ulp of the accumulated magnitude."""
    from hdn_b200 import ops
    g = torch.Generator().manual_seed(cin * 100 + cout)
    x = torch.randn(3, cin, *hw, generator=g).cuda()
    w = torch.randn(cout, cin, k, k, generator=g).cuda() * 0.2
    scale, shift = (torch.rand(cout, generator=g) + 0.5).cuda(), torch.randn(cout, generator=g).cuda()
    torch.backends.cudnn.allow_tf32 = False
    ref64 = torch.nn.functional.conv2d(x.double(), w.double(), stride=s, padding=p) * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
    for relu in (True, False):
        got = ops.conv_small(x, w, scale, shift, stride=s, padding=p, relu=relu)
        want = torch.relu(ref64) if relu else ref64
        assert got.shape == want.shape
        err = (got.double() - want).abs().max().item()
        assert err <= 2e-6 * max(1.0, ref64.abs().max().item()), err
    plain = ops.conv_small(x, w, stride=s, padding=p)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), stride=s, padding=p)
    assert (plain.double() - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())


def test_conv_small_rejects_what_it_does_not_cover():
    from hdn_b200 import ops
    x = torch.zeros(1, 16, 8, 8, device="cuda")
    with pytest.raises((ValueError, RuntimeError)):
        ops.conv_small(x, torch.zeros(4, 16, 3, 3, device="cuda"), padding=1)       # too many input channels
    with pytest.raises((ValueError, RuntimeError)):
        ops.conv_small(x[:, :3], torch.zeros(4, 3, 5, 5, device="cuda"), padding=1)   # 5x5
    with pytest.raises((ValueError, RuntimeError)):
        ops.conv_small(x[:, :3], torch.zeros(4, 3, 3, 3, device="cuda"), padding=2)   # padding beyond k // 2
    assert not ops.conv_small_supported(3, 48, 7, 2) and ops.conv_small_supported(2, 64, 7, 2) and ops.conv_small_supported(8, 1, 3, 1)


def test_share_feature_and_stems_equal_the_module_path():
    """PreShareFeature.forward and the two stems through conv_bn_act == the plain nn.Module evaluation (cuDNN) of the same layers."""
    from hdn_b200 import compat, convs, synthetic
    compat.activate()
    from hdn.models.backbone.resnet_atrous import resnet50
    from homo_estimator.Deep_homography.Oneline_DLTv1.backbone.resnet import resnet34
    from homo_estimator.Deep_homography.Oneline_DLTv1.preprocess.input_feature_extractor import PreShareFeature
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        sf = PreShareFeature().cuda().eval()
        for m in sf.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.3), m.running_var.uniform_(0.5, 2.0), m.weight.uniform_(0.5, 1.5), m.bias.normal_(0, 0.2)
        x = torch.randn(2, 1, 127, 127, device="cuda")
        got, want = sf(x), sf.ShareFeature(x)
        assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
        for net, xin in ((resnet50(used_layers=[2, 3, 4]), torch.randn(2, 3, 255, 255)), (resnet34(used_layers=[4]), torch.randn(2, 2, 127, 127))):
            net = net.cuda().eval()
            net.bn1.running_mean.normal_(0, 0.3), net.bn1.running_var.uniform_(0.5, 2.0)
            xin = xin.cuda()
            got = convs.conv_bn_act(net.conv1, net.bn1, xin, relu=True)
            want = torch.relu(net.bn1(net.conv1(xin)))
            assert got.shape == want.shape and (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("batch", [1, 3])
def test_programmatic_dependent_launch_chain_equals_plain_stream_order(batch):
    """Consecutive tcgen05 convolutions are chained by programmatic dependent launch (the next kernel's prologue and first weight
    records overlap the previous kernel's tail; activation reads and all writes wait for its completion).  A ResNet-50 forward --
    53 dependent launches with residual reads and split-K clusters -- must give bit-identical features with the chaining on and
    off, eagerly (cold: weights packed right before their first use) and as a replayed CUDA graph."""
    from hdn_b200 import compat, ops, synthetic
    compat.activate()
    from hdn.models.backbone.resnet_atrous import resnet50
    torch.manual_seed(11)
    with torch.no_grad():
        net = synthetic.fill_weights(resnet50(used_layers=[2, 3, 4])).cuda().eval()
        x = torch.randn(batch, 3, 255, 255, device="cuda") * 50
        try:
            ops.set_conv_pdl(True)
            cold = [f.clone() for f in net(x)]          # first call: every layer packs its weight, then launches
            ops.set_conv_pdl(False)
            want = [f.clone() for f in net(x)]
            ops.set_conv_pdl(True)
            assert all(torch.equal(a, b) for a, b in zip(cold, want))
            for _ in range(5):
                assert all(torch.equal(a, b) for a, b in zip(net(x), want))
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                net(x)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = net(x)
            for _ in range(3):
                x.copy_(x)  # same input; replay must reproduce the eager result
                graph.replay()
                torch.cuda.synchronize()
                assert all(torch.equal(a, b) for a, b in zip(outs, want))
        finally:
            ops.set_conv_pdl(True)


TS_CASES = [  # B, Cin, Cout, H, W, k, stride, pad, dilation  -- all large enough (>= 2 x 148 tiles of 128 x 128) to take the big-launch path
    (40, 512, 256, 31, 31, 1, 1, 0, 1),
    (24, 256, 256, 31, 31, 3, 1, 2, 2),
    (12, 512, 512, 31, 31, 3, 1, 4, 4),
    (48, 128, 128, 63, 63, 3, 2, 1, 1),
    (12, 64, 64, 63, 63, 3, 1, 1, 1),
    (10, 64, 256, 63, 63, 1, 1, 0, 1),
    (40, 256, 512, 31, 31, 1, 2, 0, 1),
    (3, 1024, 2048, 63, 63, 1, 1, 0, 1),
]


@pytest.mark.parametrize("case", TS_CASES)
def test_conv_gemm_with_activations_in_tensor_memory(case):
    """conv_gemm_ts.cu (A operand = split activations written by tcgen05.st into TMEM, B = packed weight records in shared memory,
    pixel-major accumulator, transpose-free epilogue) against fp64 and against the shared-memory-operand kernel it replaces for large
    launches: same 3xTF32 terms, same chunking, so fp32-accurate and equal to the other kernel up to the MMA's internal summation."""
    from hdn_b200 import ops
    B, Cin, Cout, H, W, k, stride, pad, d = case
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout + H + k)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = 1 + 0.1 * torch.randn(Cout, device="cuda", generator=g)
    shift = 0.1 * torch.randn(Cout, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    y64 = F.conv2d(x[:2].double(), w.double(), stride=stride, padding=pad, dilation=d) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    res = torch.randn((B,) + tuple(y64.shape[1:]), device="cuda", generator=g)
    want64 = F.relu(y64 + res[:2].double())
    try:
        ops.set_conv_ts(False)
        old = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True, stride=stride, padding=pad, cout=Cout)
        ops.set_conv_ts(2)
        new = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True, stride=stride, padding=pad, cout=Cout)
        again = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True, stride=stride, padding=pad, cout=Cout)
        plain = ops.conv_gemm(x, wp, ksize=k, dilation=d, stride=stride, padding=pad, cout=Cout)
    finally:
        ops.set_conv_ts(2)
    den = float(want64.abs().max())
    assert tuple(new.shape) == tuple(old.shape)
    assert torch.equal(new, again)
    assert float((new[:2].double() - want64).abs().max()) / den < 1e-5
    assert float((new - old).abs().max()) / den < 5e-6
    want_plain = F.conv2d(x[:1].double(), w.double(), stride=stride, padding=pad, dilation=d)
    assert float((plain[:1].double() - want_plain).abs().max()) / float(want_plain.abs().max()) < 1e-5


@pytest.mark.parametrize("case", [(1, 256, 256, 31, 31, 3, 1, 2, 2), (1, 1024, 256, 31, 31, 1, 1, 0, 1), (2, 512, 512, 15, 15, 3, 1, 1, 1), (1, 128, 128, 63, 63, 3, 2, 1, 1),
                                  (1, 512, 2048, 31, 31, 1, 1, 0, 1), (3, 256, 256, 7, 7, 3, 1, 0, 1), (1, 64, 64, 32, 32, 3, 1, 1, 1)])
def test_conv_gemm_ts_split_k_clusters(case):
    """The tensor-memory-operand kernel at tracking batch sizes (set_conv_ts(2)): K split over a cluster of 2 / 4 / 8 CTAs, the leader adds
    the peers' partial tiles to its registers through distributed shared memory.  fp32-accurate, deterministic, equal to the default
    dispatch up to the summation order."""
    from hdn_b200 import ops
    B, Cin, Cout, H, W, k, stride, pad, d = case
    g = torch.Generator(device="cuda").manual_seed(Cin * 3 + Cout + H + k)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = 1 + 0.1 * torch.randn(Cout, device="cuda", generator=g)
    shift = 0.1 * torch.randn(Cout, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    y64 = F.conv2d(x.double(), w.double(), stride=stride, padding=pad, dilation=d) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    res = torch.randn(y64.shape, device="cuda", generator=g)
    want = F.relu(y64 + res.double())
    try:
        ops.set_conv_ts(0)
        ref = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True, stride=stride, padding=pad, cout=Cout)
        ops.set_conv_ts(2)
        new = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True, stride=stride, padding=pad, cout=Cout)
        again = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True, stride=stride, padding=pad, cout=Cout)
        ops.set_conv_splitk(False)
        nosplit = ops.conv_gemm(x, wp, scale, shift, res, ksize=k, dilation=d, relu=True, stride=stride, padding=pad, cout=Cout)
    finally:
        ops.set_conv_ts(2)
        ops.set_conv_splitk(True)
    den = float(want.abs().max())
    assert torch.equal(new, again)
    assert float((new.double() - want).abs().max()) / den < 1e-5
    assert float((nosplit.double() - want).abs().max()) / den < 1e-5
    assert float((new - ref).abs().max()) / den < 5e-6
