"""ResnetEncoder — drop-in for models/resnet_encoder.py:64-101 on the sm_100a kernels."""
import numpy as np
from torch import nn

from .. import _cabi as A
from .. import ops
from .layers import conv_bn
from .resnet import CONFIGS, ResNet


class ResNetMultiImageInput(ResNet):
    """ResNet whose stem takes `num_input_images` stacked frames (reference :19-41)."""

    def __init__(self, num_layers, num_input_images=1):
        super().__init__(num_layers, in_channels=3 * num_input_images)


def resnet_multiimage_input(num_layers, pretrained=False, num_input_images=1):
    assert num_layers in [18, 50], "Can only run with 18 or 50 layer resnet"
    if pretrained:
        raise RuntimeError("ImageNet weights cannot be downloaded here (no network); load a state_dict instead")
    return ResNetMultiImageInput(num_layers, num_input_images)


class ResnetEncoder(nn.Module):
    def __init__(self, num_layers, pretrained, num_input_images=1, **kwargs):
        super().__init__()
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        if num_layers not in CONFIGS:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        if pretrained:
            raise RuntimeError("pretrained=True needs a torchvision download (no network); construct with "
                               "pretrained=False and load_state_dict() the ImageNet weights")
        if num_input_images > 1:
            self.encoder = resnet_multiimage_input(num_layers, pretrained, num_input_images)
        else:
            self.encoder = ResNet(num_layers, **kwargs)
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4

    def _trunk(self, first, second=None):
        e = self.encoder
        kw = {"nchw_norm_in": True}
        if second is not None:
            kw["x2"] = second
        self.features = [conv_bn(e.conv1, e.bn1, first, act=A.ACT_RELU, **kw)]
        x = ops.maxpool3x3s2(self.features[-1])
        for layer in (e.layer1, e.layer2, e.layer3, e.layer4):
            for blk in layer:
                x = blk(x)
            self.features.append(x)
        return self.features

    def forward(self, input_image):
        # (x - 0.45) / 0.225 is applied inside the stem kernel while it reads the NCHW image
        A.require_cuda(input_image)
        return self._trunk(input_image.float())

    def forward_pair(self, first, second):
        """Stem over the channel concat of two frames without materialising the concat
        (joint_segmentation_depth.py:44: `torch.cat(pose_inputs, 1)`)."""
        A.require_cuda(first, second)
        return self._trunk(first.float(), second.float())
