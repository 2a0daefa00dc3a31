"""PoseDecoder — drop-in for models/pose_decoder.py:18-58 on the sm_100a kernels."""
from collections import OrderedDict

from torch import nn

from .. import _cabi as A
from .. import ops
from .layers import Conv2d


class PoseDecoder(nn.Module):
    def __init__(self, num_ch_enc, num_input_features, num_frames_to_predict_for=None, stride=1):
        super().__init__()
        self.num_ch_enc = num_ch_enc
        self.num_input_features = num_input_features
        if num_frames_to_predict_for is None:
            num_frames_to_predict_for = num_input_features - 1
        self.num_frames_to_predict_for = num_frames_to_predict_for
        self.convs = OrderedDict()
        self.convs[("squeeze")] = Conv2d(int(self.num_ch_enc[-1]), 256, 1)
        self.convs[("pose", 0)] = Conv2d(num_input_features * 256, 256, 3, stride, 1)
        self.convs[("pose", 1)] = Conv2d(256, 256, 3, stride, 1)
        self.convs[("pose", 2)] = Conv2d(256, 6 * num_frames_to_predict_for, 1)
        self.relu = nn.ReLU()
        self.net = nn.ModuleList(list(self.convs.values()))

    def forward(self, input_features):
        last = [f[-1] for f in input_features]
        cat = [self.convs["squeeze"](f, act=A.ACT_RELU) for f in last]
        out = cat[0] if len(cat) == 1 else ops.cat_channels(cat)
        out = self.convs[("pose", 0)](out, act=A.ACT_RELU)
        out = self.convs[("pose", 1)](out, act=A.ACT_RELU)
        out = self.convs[("pose", 2)](out)
        out = ops.spatial_mean(out, scale=0.01)            # 0.01 * mean over W then H
        out = out.reshape(-1, self.num_frames_to_predict_for, 1, 6)
        return out[..., :3], out[..., 3:]
