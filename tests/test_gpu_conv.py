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
