"""On-device input-pipeline ops: drop-ins for `loader/transformsgpu.py:color_jitter` /
`gaussian_blur` (kornia-free) and the loader's per-scale area pyramid.  kornia itself is not available in this image, so
the arithmetic follows kornia 0.4's documented definitions (see csrc/augment.cu) and is pinned against a PyTorch
restatement in tests/test_augment.py — "parity unpinned" w.r.t. kornia."""
import ctypes as C
import math

import numpy as np
import torch

from . import _cabi as A


def gaussian_taps(k, sigma, device):
    """kornia.filters.get_gaussian_kernel1d: exp(-x^2 / (2 sigma^2)) on x = arange(k) - k // 2, normalised."""
    x = torch.arange(k, dtype=torch.float32) - k // 2
    if k % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2) / (2.0 * float(sigma) ** 2))
    return (g / g.sum()).to(device)


def gaussian_blur2d(data, kernel_size, sigma):
    """kornia.filters.GaussianBlur2d(kernel_size=(ky, kx), sigma=(sy, sx)), border_type='reflect' on B x C x H x W."""
    A.require_cuda(data)
    x = data.detach().float().contiguous()
    b, c, h, w = x.shape
    ky, kx = int(kernel_size[0]), int(kernel_size[1])
    ty, tx = gaussian_taps(ky, sigma[0], x.device), gaussian_taps(kx, sigma[1], x.device)
    tmp, y = torch.empty_like(x), torch.empty_like(x)
    A.call("segsde_gaussian_blur", A.ptr(x), A.ptr(tmp), A.ptr(y), C.c_int(b * c), C.c_int(h), C.c_int(w), C.c_int(ky),
           C.c_int(kx), A.ptr(ty), A.ptr(tx), A.stream_ptr())
    return y


def gaussian_blur(blur, data=None, target=None):
    """loader/transformsgpu.py:21-30: with blur > 0.5, sigma ~ U(0.15, 1.15) (numpy RNG, as there) and a kernel of ~10 % of
    the image size."""
    if data is not None and data.shape[1] == 3 and blur > 0.5:
        sigma = np.random.uniform(0.15, 1.15)
        ky = int(np.floor(np.ceil(0.1 * data.shape[2]) - 0.5 + np.ceil(0.1 * data.shape[2]) % 2))
        kx = int(np.floor(np.ceil(0.1 * data.shape[3]) - 0.5 + np.ceil(0.1 * data.shape[3]) % 2))
        data = gaussian_blur2d(data, (ky, kx), (sigma, sigma))
    return data, target


def apply_color_jitter(data, brightness=0.0, contrast=1.0, saturation=1.0, hue=0.0, order=(0, 1, 2, 3)):
    """The fused jitter kernel with explicit parameters: additive brightness, multiplicative contrast / saturation, hue
    shift in radians; `order` = permutation of (0 brightness, 1 contrast, 2 saturation, 3 hue)."""
    A.require_cuda(data)
    x = data.detach().float().contiguous()
    b, c, h, w = x.shape
    if c != 3:
        raise ValueError("apply_color_jitter: B x 3 x H x W images expected")
    y = torch.empty_like(x)
    o = (C.c_int * 4)(*[int(v) for v in order])
    A.call("segsde_color_jitter", A.ptr(x), A.ptr(y), C.c_int(b), C.c_int64(h * w), C.c_float(brightness), C.c_float(contrast),
           C.c_float(saturation), C.c_float(hue), o, A.stream_ptr())
    return y


def color_jitter(jitter, data=None, target=None, s=0.25):
    """loader/transformsgpu.py:10-18: with jitter > 0.2, kornia.augmentation.ColorJitter(s, s, s, s): one draw per batch
    of brightness in U(-s, s) (additive), contrast / saturation in U(1-s, 1+s), hue in U(-s, s) * 2 pi, applied in a
    random order (torch CPU RNG)."""
    if data is not None and data.shape[1] == 3 and jitter > 0.2:
        u = torch.rand(4)
        data = apply_color_jitter(data, brightness=float((u[0] * 2 - 1) * s), contrast=float(1 + (u[1] * 2 - 1) * s),
                                  saturation=float(1 + (u[2] * 2 - 1) * s), hue=float((u[3] * 2 - 1) * s * 2 * math.pi),
                                  order=torch.randperm(4).tolist())
    return data, target


def area_pyramid(img):
    """Scales 1..3 of an image batch (B x C x H x W, H and W multiples of 8) as exact box means — what
    `F.interpolate(img, (H >> s, W >> s), mode="area")` computes — in one launch."""
    A.require_cuda(img)
    x = img.detach().float().contiguous()
    b, c, h, w = x.shape
    ys = [torch.empty(b, c, h >> s, w >> s, device=x.device, dtype=torch.float32) for s in (1, 2, 3)]
    A.call("segsde_area_pyramid", A.ptr(x), C.c_int(b * c), C.c_int(h), C.c_int(w), A.ptr(ys[0]), A.ptr(ys[1]), A.ptr(ys[2]),
           A.stream_ptr())
    return ys
