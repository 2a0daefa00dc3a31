"""Layer/geometry API of models/monodepth_layers.py on the sm_100a kernels."""
import ctypes as C

import torch
from torch import nn

from .. import _cabi as A
from .. import ops
from .layers import BatchNorm2d, Conv2d, Dropout2d


def disp_to_depth(disp, min_depth, max_depth):
    """Reference :18-27. Pure scalar affine + reciprocal on a tensor the caller owns (kept in torch:
    it is only used off the training path, e.g. evaluation scripts)."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    return scaled_disp, 1 / scaled_disp


def transformation_from_parameters(axisangle, translation, invert=False):
    """Reference :30-47: (B,1,3),(B,1,3) -> (B,4,4)."""
    vec = torch.cat([axisangle.reshape(-1, 3), translation.reshape(-1, 3)], 1)
    return ops.pose_matrix(vec, invert)


def get_translation_matrix(translation_vector):
    """Reference :50-63."""
    z = torch.zeros_like(translation_vector.reshape(-1, 3))
    return ops.pose_matrix(torch.cat([z, translation_vector.reshape(-1, 3)], 1), False)


def rot_from_axisangle(vec):
    """Reference :66-105."""
    v = vec.reshape(-1, 3)
    return ops.pose_matrix(torch.cat([v, torch.zeros_like(v)], 1), False)


class Conv3x3(nn.Module):
    """Reflection- (or zero-) padded 3x3 convolution, reference :127-142; the padding is resolved by
    the convolution kernel's loader instead of materialising a padded copy."""

    def __init__(self, in_channels, out_channels, use_refl=True):
        super().__init__()
        self.pad = nn.ReflectionPad2d(1) if use_refl else nn.ZeroPad2d(1)
        self.use_refl = use_refl
        self.conv = Conv2d(int(in_channels), int(out_channels), 3)

    def forward(self, x, x2=None, up1=False, act=A.ACT_NONE):
        return self.conv(x, x2=x2, up1=up1, act=act, pad=1,
                         pad_mode=A.PAD_REFLECT if self.use_refl else A.PAD_ZERO)


class ConvBlock(nn.Module):
    """conv3x3(reflect) -> [BN] -> ELU -> [Dropout2d], reference :108-124. Without BN the ELU runs in
    the convolution epilogue."""

    def __init__(self, in_channels, out_channels, bn=False, dropout=0.0):
        super().__init__()
        self.block = nn.Sequential(
            Conv3x3(in_channels, out_channels),
            BatchNorm2d(out_channels) if bn else nn.Identity(),
            nn.ELU(inplace=True),
            Dropout2d(dropout) if dropout > 0 else nn.Identity())
        self.has_bn, self.has_dropout = bn, dropout > 0

    def forward(self, x, x2=None, up1=False):
        if self.has_bn:
            out = self.block[1](self.block[0](x, x2=x2, up1=up1))
            out = ops.activation(out, A.ACT_ELU)
        else:
            out = self.block[0](x, x2=x2, up1=up1, act=A.ACT_ELU)
        if self.has_dropout:
            out = self.block[3](out)
        return out


def upsample(x):
    """Nearest x2 (reference :202-205). Inside the decoder this is fused into the next convolution's
    loader; the standalone form is a 1x1 identity-free copy through the resize kernel."""
    n, c, h, w = x.shape
    x = ops.as_cl(x)
    y = ops.cl_empty(n, c, 2 * h, 2 * w, x.device)
    A.call("segsde_upsample2x_nearest", C.byref(ops.view(x)), C.byref(ops.view(y)), A.stream_ptr())
    return y


class _ReprojPiece(nn.Module):
    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width


class BackprojectDepth(_ReprojPiece):
    """Reference :145-174: depth (B,1,H,W), inv_K (B,4,4) -> camera points (B,4,H*W)."""

    def forward(self, depth, inv_K):
        A.require_cuda(depth, inv_K)
        B = depth.shape[0]
        out = torch.empty(B, 4, self.height * self.width, device=depth.device, dtype=torch.float32)
        A.call("segsde_backproject", A.ptr(depth.contiguous().float()), A.ptr(inv_K.contiguous().float()),
               C.c_int(B), C.c_int(self.height), C.c_int(self.width), A.ptr(out), A.stream_ptr())
        return out


class Project3D(_ReprojPiece):
    """Reference :177-199: points (B,4,H*W), K, T -> normalised sampling grid (B,H,W,2)."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__(batch_size, height, width)
        self.eps = eps

    def forward(self, points, K, T):
        A.require_cuda(points, K, T)
        B = points.shape[0]
        out = torch.empty(B, self.height, self.width, 2, device=points.device, dtype=torch.float32)
        A.call("segsde_project3d", A.ptr(points.contiguous().float()), A.ptr(K.contiguous().float()),
               A.ptr(T.contiguous().float()), C.c_int(B), C.c_int(self.height), C.c_int(self.width),
               C.c_float(self.eps), A.ptr(out), A.stream_ptr())
        return out


def get_smooth_loss(disp, img):
    """Reference :208-221 (no mean normalisation: the caller passes the normalised disparity)."""
    A.require_cuda(disp, img)
    B, _, h, w = disp.shape
    d, im = disp.detach().contiguous().float(), img.contiguous().float()
    ones = torch.full((B,), 1.0 - 1e-7, device=d.device)        # kernel divides by (mean + 1e-7)
    acc = torch.zeros(2 + B, device=d.device)
    A.call("segsde_smooth_fused", A.ptr(d), A.ptr(im), A.ptr(ones), C.c_int(B), C.c_int(h), C.c_int(w),
           A.ptr(acc), None, A.stream_ptr())
    return acc[0] + acc[1]


class SSIM(nn.Module):
    """Reference :224-254: per-pixel SSIM loss map of two images (B,C,H,W), 3x3 mean, reflection pad."""

    def __init__(self):
        super().__init__()
        self.C1, self.C2 = 0.01 ** 2, 0.03 ** 2

    def forward(self, x, y):
        A.require_cuda(x, y)
        out = torch.empty_like(x, dtype=torch.float32, memory_format=torch.contiguous_format)
        B, Cc, H, W = x.shape
        A.call("segsde_ssim_map", A.ptr(x.contiguous().float()), A.ptr(y.contiguous().float()),
               C.c_int(B * Cc), C.c_int(H), C.c_int(W), A.ptr(out), A.stream_ptr())
        return out
