"""Plain ResNet trunk of the homography estimator -- mirror of Oneline_DLTv1/backbone/resnet.py:60-220.

Differences from torchvision's: the stem takes the 2-channel (template, search) gray pair
(`Conv2d(2, 64, 7, stride 2, padding 3)`, resnet.py:143) and there is no avgpool/fc inside (the model builder
owns those); `forward` returns the feature levels listed in `used_layers`.  Same state-dict keys.
"""
import torch.nn as nn

from hdn_b200.convs import conv_bn_act


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = conv_bn_act(self.conv1, self.bn1, x, relu=True)
        shortcut = x if self.downsample is None else conv_bn_act(self.downsample[0], self.downsample[1], x)  # 1x1 stride-2 projection
        return conv_bn_act(self.conv2, self.bn2, y, residual=shortcut, relu=True)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = conv_bn_act(self.conv1, self.bn1, x, relu=True)
        y = conv_bn_act(self.conv2, self.bn2, y, relu=True)
        shortcut = x if self.downsample is None else conv_bn_act(self.downsample[0], self.downsample[1], x)
        return conv_bn_act(self.conv3, self.bn3, y, residual=shortcut, relu=True)


class ResNet(nn.Module):
    def __init__(self, block, layers, used_layers):
        super().__init__()
        self.inplanes = 64
        self.used_layers = used_layers
        self.conv1 = nn.Conv2d(2, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        for idx, (planes, depth) in enumerate(zip((64, 128, 256, 512), layers), start=1):
            setattr(self, "layer%d" % idx, self._stage(block, planes, depth, 1 if idx == 1 else 2))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _stage(self, block, planes, depth, stride):
        out_ch = planes * block.expansion
        shortcut = None
        if stride != 1 or self.inplanes != out_ch:
            shortcut = nn.Sequential(nn.Conv2d(self.inplanes, out_ch, 1, stride=stride, bias=False), nn.BatchNorm2d(out_ch))
        blocks = [block(self.inplanes, planes, stride, shortcut)]
        self.inplanes = out_ch
        blocks += [block(out_ch, planes) for _ in range(1, depth)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        y = self.maxpool(conv_bn_act(self.conv1, self.bn1, x, relu=True))
        feats = [y]
        for idx in range(1, 5):
            y = getattr(self, "layer%d" % idx)(y)
            feats.append(y)
        picked = [feats[i] for i in self.used_layers]
        return picked[0] if len(picked) == 1 else picked


_DEPTHS = {"resnet18": (BasicBlock, [2, 2, 2, 2]), "resnet34": (BasicBlock, [3, 4, 6, 3]), "resnet50": (Bottleneck, [3, 4, 6, 3]),
           "resnet101": (Bottleneck, [3, 4, 23, 3]), "resnet152": (Bottleneck, [3, 8, 36, 3])}


def build(name, **kwargs):
    block, layers = _DEPTHS[name]
    return ResNet(block, layers, **kwargs)


def resnet18(pretrained=False, **kwargs):
    return build("resnet18", **kwargs)


def resnet34(pretrained=False, **kwargs):
    return build("resnet34", **kwargs)


def resnet50(pretrained=False, **kwargs):
    return build("resnet50", **kwargs)
