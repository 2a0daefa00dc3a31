"""ASPP and SelfAttention (models/model_parts.py) on the sm_100a kernels; sub-module layout matches
the reference + torchvision's ASPPConv / ASPPPooling so the state_dict keys are identical
(`convs.0.0`, `convs.{1..3}.0/1`, `convs.4.1/2`, `project.0/1`)."""
import torch
from torch import nn

from .. import _cabi as A
from .. import ops
from .layers import BatchNorm2d, Conv2d, Dropout, conv_bn


class ASPPConv(nn.Sequential):
    def __init__(self, in_channels, out_channels, dilation):
        super().__init__(Conv2d(in_channels, out_channels, 3, padding=dilation, dilation=dilation, bias=False),
                         BatchNorm2d(out_channels), nn.ReLU())

    def forward(self, x):
        return conv_bn(self[0], self[1], x, act=A.ACT_RELU)


class ASPPPooling(nn.Sequential):
    def __init__(self, in_channels, out_channels):
        super().__init__(nn.AdaptiveAvgPool2d(1), Conv2d(in_channels, out_channels, 1, bias=False),
                         BatchNorm2d(out_channels), nn.ReLU())

    def forward(self, x):
        h, w = x.shape[-2:]
        g = self[2](self[1](ops.spatial_mean(x)), act=A.ACT_RELU)
        return ops.broadcast_hw(g, h, w)       # bilinear resize from 1x1 == broadcast


class ASPP(nn.Module):
    def __init__(self, in_channels, atrous_rates, aspp_pooling=True, out_channels=256):
        super().__init__()
        modules = [nn.Sequential(Conv2d(in_channels, out_channels, 1, bias=False), BatchNorm2d(out_channels),
                                 nn.ReLU())]
        for r in atrous_rates:
            modules.append(ASPPConv(in_channels, out_channels, r))
        if aspp_pooling:
            modules.append(ASPPPooling(in_channels, out_channels))
        self.convs = nn.ModuleList(modules)
        self.project = nn.Sequential(
            Conv2d((1 + int(aspp_pooling) + len(atrous_rates)) * out_channels, out_channels, 1, bias=False),
            BatchNorm2d(out_channels), nn.ReLU(), Dropout(0.5))

    def forward(self, x):
        first = self.convs[0]
        res = [conv_bn(first[0], first[1], x, act=A.ACT_RELU)]
        for conv in list(self.convs)[1:]:
            res.append(conv(x))
        y = conv_bn(self.project[0], self.project[1], ops.cat_channels(res), act=A.ACT_RELU)
        return self.project[3](y)


class SelfAttention(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv2d(in_channels, out_channels, 3, padding=1, bias=False)
        self.attention = Conv2d(in_channels, out_channels, 3, padding=1, bias=False)
        with torch.no_grad():
            self.attention.weight.zero_()

    def forward(self, x):
        return ops.gate(self.conv(x), self.attention(x))
