"""DepthDecoder — drop-in for models/depth_decoder.py:22-116 on the sm_100a kernels.

Per stage i (coarse to fine): `upconv(i,0)` (an ASPP at the bottleneck when `intermediate_aspp`), then — fused into the
loader of `upconv(i,1)`, never materialised — the nearest x2 upsample and the concat with skip feature i-1, then
`upconv(i,1)`; stages listed in `scales` also emit a sigmoid disparity head.  Layer construction order fixes the
`decoder.<n>` indices of the reference's state_dict, so it is produced by one ordered plan (`_layer_plan`).
"""
from collections import OrderedDict

import numpy as np
from torch import nn

from .. import _cabi as A
from .layers import BatchNorm2d, Conv2d
from .model_parts import ASPP
from .monodepth_layers import Conv3x3, ConvBlock


class _SkipProj(nn.Sequential):
    """1x1 conv -> BN -> ReLU projection of a skip feature (reference :56-65); BN + ReLU run as one kernel."""

    def forward(self, x):
        conv, bn = self[0], self[1]
        return bn(conv(x), act=A.ACT_RELU)


class DepthDecoder(nn.Module):
    def __init__(self, num_ch_enc, scales, max_scale_size, num_output_channels=1, use_skips=True,
                 intermediate_aspp=False, aspp_rates=[6, 12, 18], num_ch_dec=[16, 32, 64, 128, 256],
                 n_upconv=4, batch_norm=False, dropout=0.0, n_project_skip_ch=-1,
                 aspp_pooling=True):
        super().__init__()
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array(num_ch_dec)
        self.n_upconv = n_upconv
        self.scales = scales
        self.use_skips = use_skips
        self.num_output_channels = num_output_channels
        self.max_scale_size = np.asarray(max_scale_size)
        self.upsample_mode = "nearest"
        self.enable_disparity = True          # PAD switches the heads of its segmentation decoder off
        self.sigmoid = nn.Sigmoid()           # attribute kept for surface compatibility; the heads fuse their sigmoid
        block = dict(bn=batch_norm, dropout=dropout)
        self.convs = OrderedDict(self._layer_plan(intermediate_aspp, aspp_rates, aspp_pooling, n_project_skip_ch, block))
        self.decoder = nn.ModuleList(self.convs.values())

    # ------------------------------------------------------------------------------------------ construction
    def _layer_plan(self, intermediate_aspp, aspp_rates, aspp_pooling, n_project_skip_ch, block):
        """(key, module) pairs in state_dict order: per stage upconv(i,0) [, skip_proj(i)], upconv(i,1); then the heads."""
        enc = [int(c) for c in self.num_ch_enc]
        dec = [int(c) for c in self.num_ch_dec]
        top = self.n_upconv
        for i in range(top, -1, -1):
            width = dec[i]
            fan_in = enc[-1] if i == top else dec[i + 1]
            if i == top and intermediate_aspp:
                yield ("upconv", i, 0), ASPP(fan_in, aspp_rates, aspp_pooling, width)
            else:
                yield ("upconv", i, 0), ConvBlock(fan_in, width, **block)
            skip_ch = 0
            if self.use_skips and i > 0:
                if n_project_skip_ch == -1:
                    skip_ch = enc[i - 1]
                    yield ("skip_proj", i), nn.Identity()
                else:
                    skip_ch = n_project_skip_ch
                    yield ("skip_proj", i), _SkipProj(Conv2d(enc[i - 1], skip_ch, kernel_size=1), BatchNorm2d(skip_ch),
                                                      nn.ReLU(inplace=True))
            yield ("upconv", i, 1), ConvBlock(width + skip_ch, width, **block)
        for s in self.scales:
            yield ("dispconv", s), Conv3x3(dec[s], self.num_output_channels)

    # ------------------------------------------------------------------------------------------------ forward
    def _stage(self, i, x, feats, out):
        x = self.convs[("upconv", i, 0)](x)
        skip = self.convs[("skip_proj", i)](feats[i - 1]) if (self.use_skips and i > 0) else None
        # upsample iff the running map is narrower than the skip it is fused with, and always at the finest stage
        # (reference :93; with a dilated encoder the coarsest two features share a resolution)
        upsample = bool(i == 0 or x.shape[-1] < feats[i - 1].shape[-1])
        x = self.convs[("upconv", i, 1)](x, x2=skip, up1=upsample)
        out[("upconv", i)] = x
        if self.enable_disparity and i in self.scales:
            out[("disp", i)] = self.convs[("dispconv", i)](x, act=A.ACT_SIGMOID)
        return x

    def forward(self, input_features, x=None, exec_layer=None):
        """`x` injects the running feature map (default: the deepest encoder feature), `exec_layer` restricts the
        stages that run (PAD drives the decoder in two halves, joint_segmentation_depth_decoder.py:140-170)."""
        stages = range(self.n_upconv, -1, -1)
        if exec_layer is not None and exec_layer != "all":
            stages = [i for i in stages if i in exec_layer]
        x = input_features[-1] if x is None else x
        self.outputs = {}
        for i in stages:
            x = self._stage(i, x, input_features, self.outputs)
        return self.outputs
