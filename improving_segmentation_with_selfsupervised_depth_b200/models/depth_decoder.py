"""DepthDecoder — drop-in for models/depth_decoder.py:22-116 on the sm_100a kernels.
The nearest x2 upsample and the skip concat are never materialised: both are resolved by the loader
of the following convolution (`up1`, second source)."""
from collections import OrderedDict

import numpy as np
from torch import nn

from .. import _cabi as A
from .layers import BatchNorm2d, Conv2d
from .model_parts import ASPP
from .monodepth_layers import Conv3x3, ConvBlock


class _SkipProj(nn.Sequential):
    def forward(self, x):
        return self[1](self[0](x), act=A.ACT_RELU)


class DepthDecoder(nn.Module):
    first_iter = True

    def __init__(self, num_ch_enc, scales, max_scale_size, num_output_channels=1, use_skips=True,
                 intermediate_aspp=False, aspp_rates=[6, 12, 18], num_ch_dec=[16, 32, 64, 128, 256],
                 n_upconv=4, batch_norm=False, dropout=0.0, n_project_skip_ch=-1,
                 aspp_pooling=True):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.use_skips = use_skips
        self.upsample_mode = 'nearest'
        self.scales = scales
        self.enable_disparity = True
        self.max_scale_size = np.asarray(max_scale_size)
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array(num_ch_dec)
        self.n_upconv = n_upconv

        self.convs = OrderedDict()
        for i in range(self.n_upconv, -1, -1):
            cin = int(self.num_ch_enc[-1] if i == self.n_upconv else self.num_ch_dec[i + 1])
            cout = int(self.num_ch_dec[i])
            if i == self.n_upconv and intermediate_aspp:
                self.convs[("upconv", i, 0)] = ASPP(cin, aspp_rates, aspp_pooling, cout)
            else:
                self.convs[("upconv", i, 0)] = ConvBlock(cin, cout, bn=batch_norm, dropout=dropout)
            cin = int(self.num_ch_dec[i])
            if self.use_skips and i > 0:
                if n_project_skip_ch == -1:
                    cin += int(self.num_ch_enc[i - 1])
                    self.convs[("skip_proj", i)] = nn.Identity()
                else:
                    cin += n_project_skip_ch
                    self.convs[("skip_proj", i)] = _SkipProj(
                        Conv2d(int(self.num_ch_enc[i - 1]), n_project_skip_ch, kernel_size=1),
                        BatchNorm2d(n_project_skip_ch), nn.ReLU(inplace=True))
            self.convs[("upconv", i, 1)] = ConvBlock(cin, cout, bn=batch_norm, dropout=dropout)
        for s in self.scales:
            self.convs[("dispconv", s)] = Conv3x3(int(self.num_ch_dec[s]), self.num_output_channels)
        self.decoder = nn.ModuleList(list(self.convs.values()))
        self.sigmoid = nn.Sigmoid()

    def forward(self, input_features, x=None, exec_layer=None):
        self.outputs = {}
        if x is None:
            x = input_features[-1]
        if exec_layer is None:
            exec_layer = "all"
        verbose = DepthDecoder.first_iter
        if verbose:
            print(f"bottleneck shape {x.shape}")
        for i in range(self.n_upconv, -1, -1):
            if exec_layer != "all" and i not in exec_layer:
                continue
            x = self.convs[("upconv", i, 0)](x)
            if verbose:
                print(f"upconv{i}-0 shape: {x.shape}")
            up = bool(x.shape[-1] < input_features[i - 1].shape[-1] or i == 0)
            skip = None
            if self.use_skips and i > 0:
                skip = self.convs[("skip_proj", i)](input_features[i - 1])
            x = self.convs[("upconv", i, 1)](x, x2=skip, up1=up)
            self.outputs[("upconv", i)] = x
            if verbose:
                print(f"upconv{i}-1 shape: {x.shape}")
            if i in self.scales and self.enable_disparity:
                self.outputs[("disp", i)] = self.convs[("dispconv", i)](x, act=A.ACT_SIGMOID)
                if verbose:
                    print(f"disp{i} shape: {self.outputs[('disp', i)].shape}, expected {self.max_scale_size // (2 ** i)}")
            if i == 0:
                DepthDecoder.first_iter = False
        return self.outputs
