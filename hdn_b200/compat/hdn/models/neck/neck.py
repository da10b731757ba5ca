"""Adjust ("neck") layers -- mirror of hdn/models/neck/neck.py.

AdjustLayer = 1x1 conv + BN; when `cut` is set and the map is narrower than 20 it keeps the central
[cut_left : cut_left + cut_num] window (7x7 of a 15x15 template map; neck.py:22-29).  The log-polar neck is built
with cut=False (model_builder...py:50-51).  Key names: downsample{2,3,4}.downsample.{0,1}.
"""
import torch.nn as nn

from hdn_b200.convs import conv_bn_act


class AdjustLayer(nn.Module):
    def __init__(self, in_channels, out_channels, cut=True, cut_left=4, cut_num=7):
        super().__init__()
        self.downsample = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False), nn.BatchNorm2d(out_channels))
        self.cut, self.cut_left, self.cut_num = cut, cut_left, cut_num

    def forward(self, x):
        y = conv_bn_act(self.downsample[0], self.downsample[1], x)
        if self.cut and y.size(3) < 20:
            lo, hi = self.cut_left, self.cut_left + self.cut_num
            y = y[:, :, lo:hi, lo:hi]
        return y


class AdjustAllLayer(nn.Module):
    def __init__(self, in_channels, out_channels, cut=True, cut_left=4, cut_num=7):
        super().__init__()
        self.num = len(out_channels)
        if self.num == 1:
            self.downsample = AdjustLayer(in_channels[0], out_channels[0], cut, cut_left, cut_num)
        else:
            for level, (cin, cout) in enumerate(zip(in_channels, out_channels), start=2):
                self.add_module("downsample%d" % level, AdjustLayer(cin, cout, cut, cut_left, cut_num))

    def forward(self, features):
        if self.num == 1:
            return self.downsample(features)
        return [getattr(self, "downsample%d" % (i + 2))(f) for i, f in enumerate(features)]
