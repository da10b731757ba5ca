"""Stride-8 ("atrous") ResNet feature pyramid -- mirror of hdn/models/backbone/resnet_atrous.py.

Same module tree and therefore the same state-dict keys (conv1, bn1, layer{1..4}.{i}.conv{1,2,3}/bn{1,2,3}/
downsample.{0,1}), same geometry:
  * stem: 7x7 stride-2 conv with padding 0 (resnet_atrous.py:117), 3x3 stride-2 max-pool padding 1;
  * layer2 keeps stride 2; layer3 / layer4 run at stride 1 with dilation 2 / 4 (:131-139);
  * a stage's first block (the one with a projection shortcut) uses HALF the stage dilation, and its 3x3 pads by
    `2 - stride` when undilated -- i.e. layer2's stride-2 3x3 has padding 0 (:68-80);
  * projection shortcuts are 1x1 only for layer1; every other stage uses a 3x3 (padding 0 at stride 2, else the
    halved dilation) (:150-173).
Every 1x1 / 3x3 convolution (+ BatchNorm, residual, ReLU) with Cin % 32 == 0 and Cout % 64 == 0 -- all of layer1..layer4, strided
and dilated alike -- is one fused tcgen05 launch (hdn_b200.convs.conv_bn_act); the 7x7 stem (3 input channels) stays on cuDNN.
"""
import math

import torch.nn as nn

from hdn_b200.convs import conv_bn_act

__all__ = ["ResNet", "resnet18", "resnet34", "resnet50"]


def _conv_geometry(stride, dilation, has_shortcut):
    """(dilation, padding) of a block's strided/dilated 3x3 given the stage parameters."""
    if has_shortcut and dilation > 1:
        dilation //= 2
        return dilation, dilation
    if dilation > 1:
        return dilation, dilation
    return 1, 2 - stride


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super().__init__()
        dil, pad = _conv_geometry(stride, dilation, downsample is not None)
        if stride != 1 and dil != 1:
            raise ValueError("a block is either strided or dilated, not both")
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=pad, dilation=dil, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        # Each conv + BN (+ residual) (+ ReLU) is one fused launch on the tensor cores where the layer is eligible
        # (stride 1, Cin % 32 == 0, Cout % 128 == 0: hdn_b200.convs.conv_bn_act); otherwise cuDNN / shifted GEMMs.
        y = conv_bn_act(self.conv1, self.bn1, x, relu=True)
        y = conv_bn_act(self.conv2, self.bn2, y, relu=True)
        shortcut = x if self.downsample is None else conv_bn_act(self.downsample[0], self.downsample[1], x)
        return conv_bn_act(self.conv3, self.bn3, y, residual=shortcut, relu=True)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super().__init__()
        # resnet_atrous.py:21-31: the first 3x3 takes the halved dilation of a projecting block, the second the full one
        dil1, pad1 = _conv_geometry(stride, dilation, downsample is not None)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=pad1, dilation=dil1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = conv_bn_act(self.conv1, self.bn1, x, relu=True)
        shortcut = x if self.downsample is None else conv_bn_act(self.downsample[0], self.downsample[1], x)
        return conv_bn_act(self.conv2, self.bn2, y, residual=shortcut, relu=True)


class ResNet(nn.Module):
    # (planes, stride, dilation) of layer1..layer4
    STAGES = ((64, 1, 1), (128, 2, 1), (256, 1, 2), (512, 1, 4))

    def __init__(self, block, layers, used_layers):
        super().__init__()
        self.inplanes = 64
        self.used_layers = used_layers
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=0, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        deepest = max(used_layers)
        for idx, ((planes, stride, dilation), depth) in enumerate(zip(self.STAGES, layers), start=1):
            if idx <= 2 or idx <= deepest:
                stage = self._stage(block, planes, depth, stride, dilation)
                self.feature_size = planes * block.expansion
            else:
                stage = nn.Identity()  # registers no parameters, like the reference's lambda (:134,141)
            setattr(self, "layer%d" % idx, stage)
        for m in self.modules():  # resnet_atrous.py:143-149
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _stage(self, block, planes, depth, stride, dilation):
        out_ch = planes * block.expansion
        shortcut = None
        if stride != 1 or self.inplanes != out_ch:
            if stride == 1 and dilation == 1:
                proj = nn.Conv2d(self.inplanes, out_ch, 1, stride=stride, bias=False)
            else:
                dd = dilation // 2 if dilation > 1 else 1
                proj = nn.Conv2d(self.inplanes, out_ch, 3, stride=stride, padding=dd if dilation > 1 else 0, dilation=dd, bias=False)
            shortcut = nn.Sequential(proj, nn.BatchNorm2d(out_ch))
        blocks = [block(self.inplanes, planes, stride, shortcut, dilation=dilation)]
        self.inplanes = out_ch
        blocks += [block(out_ch, planes, dilation=dilation) for _ in range(1, depth)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        stem = conv_bn_act(self.conv1, self.bn1, x, relu=True)
        feats = [stem]
        y = self.maxpool(stem)
        for idx in range(1, 5):
            y = getattr(self, "layer%d" % idx)(y)
            feats.append(y)
        picked = [feats[i] for i in self.used_layers]
        return picked[0] if len(picked) == 1 else picked


def resnet18(**kwargs):
    return ResNet(BasicBlock, [2, 2, 2, 2], **kwargs)


def resnet34(**kwargs):
    return ResNet(BasicBlock, [3, 4, 6, 3], **kwargs)


def resnet50(**kwargs):
    return ResNet(Bottleneck, [3, 4, 6, 3], **kwargs)
