"""Dense-convolution helpers for the backbones (library GEMM territory: SURVEY 2.3 K7).

cuDNN has no fast fp32 kernel for the stride-8 backbone's DILATED 3x3 convolutions at tracking batch sizes: with TF32
off it falls back to `conv2d_grouped_direct_kernel`, 2.7 ms for one 512->512 31x31 dilation-4 layer on a B200
(1.7 TFLOP/s) -- three such layers were 61 % of the whole ResNet-50 forward (13.6 ms at batch 1; profiles/README.md).

`dilated_conv3x3` computes the same convolution as 9 plain fp32 GEMMs (cuBLAS, TF32 off): with the input zero-padded
by the dilation d and each plane flattened, tap (ky,kx) reads the contiguous slice starting at ky*d*Wp + kx*d, so
    out_flat[:, q] = sum_tap  W[:, :, ky, kx] @ xpad_flat[:, off_tap + q],   q = r*Wp + c,
and the valid outputs are the columns c < W of that Wp-strided grid (W/Wp of the work is useful: 31/39 at d = 4).
0.30 ms for the same layer; results differ from cuDNN's by fp32 summation order only (~3e-6 relative).
"""
import os

import torch
import torch.nn.functional as F

def _tap_major(conv):
    """[Cout,Cin,3,3] -> [9,Cout,Cin] contiguous.  Cached ON THE MODULE (it lives and dies with it; a module-level dict keyed by id()
    could hand a new model the packed weights of a freed one) and re-derived when the parameter changes (load_state_dict,
    .cuda(), in-place updates bump data_ptr / _version)."""
    weight = conv.weight
    ver = (weight.data_ptr(), weight._version, weight.device)
    hit = conv.__dict__.get("_hdn_tap_major")
    if hit is None or hit[0] != ver:
        hit = (ver, weight.detach().permute(2, 3, 0, 1).reshape(9, weight.shape[0], weight.shape[1]).contiguous())
        conv.__dict__["_hdn_tap_major"] = hit
    return hit[1]


def wants_shifted_gemm(conv, x):
    return (x.is_cuda and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.groups == 1 and conv.bias is None
            and conv.dilation[0] == conv.dilation[1] > 1 and conv.padding == conv.dilation and conv.in_channels >= 512
            and x.dtype == torch.float32)


def dilated_conv3x3(x, conv):
    """conv(x) for a stride-1 3x3 convolution with padding == dilation, as 9 shifted GEMMs."""
    d = conv.dilation[0]
    wt = _tap_major(conv)
    B, Cin, H, W = x.shape
    Wp = W + 2 * d
    xp = F.pad(x, (d, d, d, d)).reshape(B, Cin, -1)
    L = (H - 1) * Wp + W
    out = x.new_empty((B, wt.shape[1], H * Wp))
    acc = out[:, :, :L]
    for tap in range(9):
        off = (tap // 3) * d * Wp + (tap % 3) * d
        xs = xp[:, :, off:off + L]
        if tap == 0:
            torch.matmul(wt[0], xs, out=acc) if B == 1 else torch.bmm(wt[0].expand(B, -1, -1), xs, out=acc)
        else:
            acc.baddbmm_(wt[tap].expand(B, -1, -1), xs)
    return out.view(B, -1, H, Wp)[:, :, :, :W]


def _folded(conv, bn):
    """(tensor-core packed weight, scale, shift) of conv -> eval-mode BatchNorm.  Stored on the conv module itself (see _tap_major)
    together with the identity and version of every tensor it was derived from."""
    ver = (id(bn), conv.weight.data_ptr(), conv.weight._version, bn.weight.data_ptr(), bn.weight._version, bn.bias._version,
           bn.running_mean._version, bn.running_var._version, bn.running_mean.data_ptr(), bn.running_var.data_ptr(), float(bn.eps))
    hit = conv.__dict__.get("_hdn_folded")
    if hit is None or hit[0] != ver:
        from hdn_b200 import ops
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
        shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
        hit = (ver, ops.pack_conv_weight(conv.weight), scale, shift)
        conv.__dict__["_hdn_folded"] = hit
    return hit[1:]


def tensor_core_eligible(conv, x):
    """The tcgen05 3xTF32 kernels cover 1x1 and 3x3 convolutions with stride 1 or 2, zero padding 0 <= p <= dilation * (k // 2),
    Cin % 32 == 0 and Cout % 64 == 0 -- every convolution of the two ResNets and the heads except the 7x7 stems."""
    k = conv.kernel_size[0]
    return (x.is_cuda and x.dtype == torch.float32 and conv.kernel_size[0] == conv.kernel_size[1] and k in (1, 3) and conv.stride in ((1, 1), (2, 2))
            and conv.groups == 1 and conv.bias is None and conv.dilation[0] == conv.dilation[1] and conv.padding[0] == conv.padding[1]
            and isinstance(conv.padding[0], int) and 0 <= conv.padding[0] <= conv.dilation[0] * (k // 2) and conv.padding_mode == "zeros"
            and conv.in_channels % 32 == 0 and conv.out_channels % 64 == 0 and not torch.is_grad_enabled())


def small_eligible(conv, x):
    """The direct few-channel kernel (hdn_conv_small_f32): the 7x7 stride-2 stems and PreShareFeature's 3x3 layers."""
    k = conv.kernel_size[0]
    return (x.is_cuda and x.dtype == torch.float32 and conv.kernel_size[0] == conv.kernel_size[1] and conv.stride[0] == conv.stride[1]
            and (k, conv.stride[0]) in ((7, 2), (3, 1)) and conv.groups == 1 and conv.bias is None and conv.dilation == (1, 1)
            and conv.padding[0] == conv.padding[1] and isinstance(conv.padding[0], int) and 0 <= conv.padding[0] <= k // 2
            and conv.padding_mode == "zeros" and conv.in_channels <= 8 and not torch.is_grad_enabled()
            and ((k == 7 and conv.out_channels % 32 == 0) or (k == 3 and (conv.out_channels == 1 or conv.out_channels % 4 == 0))))


def _folded_bn(conv, bn):
    """(scale, shift) of an eval-mode BatchNorm, cached on the conv module like _folded."""
    ver = (id(bn), bn.weight.data_ptr(), bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           bn.running_mean.data_ptr(), bn.running_var.data_ptr(), float(bn.eps))
    hit = conv.__dict__.get("_hdn_folded_bn")
    if hit is None or hit[0] != ver:
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
        shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
        hit = (ver, scale, shift)
        conv.__dict__["_hdn_folded_bn"] = hit
    return hit[1:]


def conv_bn_act(conv, bn, x, residual=None, relu=False):
    """relu?(bn(conv(x)) + residual?) for an eval-mode block.  Eligible layers run as ONE tcgen05 launch (implicit GEMM, fp32-accurate
    3xTF32, BatchNorm / residual / ReLU in the epilogue: hdn_conv_gemm_ex_f32), the few-channel stems and PreShareFeature layers as one
    direct-sum launch (hdn_conv_small_f32); anything else (training mode, CPU tensors, other shapes) falls back to cuDNN."""
    if USE_TENSOR_CORES and not bn.training and tensor_core_eligible(conv, x):
        from hdn_b200 import ops
        wt, scale, shift = _folded(conv, bn)
        return ops.conv_gemm(x, wt, scale, shift, residual, ksize=conv.kernel_size[0], dilation=conv.dilation[0], relu=relu,
                             stride=conv.stride[0], padding=conv.padding[0], cout=conv.out_channels)
    if USE_TENSOR_CORES and not bn.training and residual is None and small_eligible(conv, x):
        from hdn_b200 import ops
        scale, shift = _folded_bn(conv, bn)
        return ops.conv_small(x, conv.weight.detach(), scale, shift, stride=conv.stride[0], padding=conv.padding[0], relu=relu)
    y = bn(conv3x3(conv, x))
    if residual is not None:
        y = y + residual
    return torch.relu_(y) if relu else y


USE_TENSOR_CORES = os.environ.get("HDN_B200_TCGEN05", "1") != "0"


def conv3x3(conv, x):
    """Inference path only: with autograd recording, defer to the module (training is out of scope)."""
    if wants_shifted_gemm(conv, x) and not (torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad)):
        return dilated_conv3x3(x, conv)
    return conv(x)
