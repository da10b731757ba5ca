"""`PreShareFeature` -- mirror of Oneline_DLTv1/preprocess/input_feature_extractor.py:3-29: three 3x3 conv + BN + ReLU
layers (1 -> 4 -> 8 -> 1 channels, padding 1) applied to each gray patch before the homography trunk."""
import torch.nn as nn

from hdn_b200.convs import conv_bn_act


class PreShareFeature(nn.Module):
    def __init__(self):
        super().__init__()
        layers = []
        for cin, cout in ((1, 4), (4, 8), (8, 1)):
            layers += [nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
        self.ShareFeature = nn.Sequential(*layers)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def forward(self, x):
        # conv -> BatchNorm -> ReLU triples: one direct-sum launch each on the device (hdn_conv_small_f32), the modules themselves otherwise
        layers = self.ShareFeature
        for i in range(0, len(layers), 3):
            x = conv_bn_act(layers[i], layers[i + 1], x, relu=True)
        return x
