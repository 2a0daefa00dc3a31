"""Segmentation losses — drop-in for loss/loss.py on the sm_100a kernels."""
import ctypes as C

import torch

from .. import _cabi as A
from .. import ops


def cross_entropy2d(input, target, class_weight=None, pixel_weights=None):
    """Reference :17-37: ignore_index 250, mean over valid pixels; with pixel_weights the mean runs over
    all pixels.  A NaN anywhere in the pixel weights disables the weighting (reference :31-32; detected on the
    device, no host round trip).  Labels outside [0, C) other than 250 contribute nothing and are counted
    (ops.CHECK_LABELS turns them into an error)."""
    if class_weight is not None:
        raise NotImplementedError("class_weight is unused by every reference config (train.py:503,649)")
    n, c, h, w = input.size()
    nt, ht, wt = target.size()
    if h != ht and w != wt:
        input = ops.bilinear(input, (ht, wt), align_corners=True)
    return ops.cross_entropy(input, target, pixel_weights=pixel_weights, ignore_index=250)


class _BerhuFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, mask, apply_log, threshold):
        A.require_cuda(x, target)
        x = x.contiguous().float()
        target = target.detach().contiguous().float()
        m = mask.detach().expand_as(x).contiguous().float() if mask is not None else None
        if target.shape != x.shape:
            raise ValueError("berhu: input %s and target %s differ in shape" % (tuple(x.shape), tuple(target.shape)))
        dev, st = x.device, A.stream_ptr()
        maxbits = torch.zeros(1, device=dev, dtype=torch.int32)
        acc = ops.zeros_f64(1, dev)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        A.call("segsde_berhu_fwd", A.ptr(x), A.ptr(target), A.ptr(m), C.c_int64(x.numel()), C.c_int(int(apply_log)),
               C.c_float(threshold), A.ptr(maxbits), A.ptr(acc), A.ptr(loss), st)
        ctx.save_for_backward(x, target, m, maxbits)
        ctx.cfg = (int(apply_log), threshold)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, target, m, maxbits = ctx.saved_tensors
        apply_log, threshold = ctx.cfg
        dx = torch.empty_like(x)
        A.call("segsde_berhu_bwd", A.ptr(x), A.ptr(target), A.ptr(m), C.c_int64(x.numel()), C.c_int(apply_log),
               C.c_float(threshold), A.ptr(maxbits), A.ptr(g.contiguous().float()), A.ptr(dx), A.stream_ptr())
        return dx, None, None, None, None


def berhu(input, target, mask, apply_log=False):
    """Reference :5-15 (pseudo-depth loss, train.py:494): reverse Huber with the switch point at 0.2 * max of the
    masked absolute difference.  The maximum never leaves the device (the reference reads it with .item()); like
    there it is a constant for the gradient."""
    return _BerhuFn.apply(input, target, mask, bool(apply_log), 0.2)


def pixel_wise_entropy(logits, normalize=False):
    """Reference :40-47: entropy of the channel softmax per pixel in units of log2(C), optionally min-max
    normalised over the whole tensor.  Forward only (label_selection.py:449 calls it under no_grad)."""
    assert logits.dim() == 4
    A.require_cuda(logits)
    x = logits.detach().contiguous().float()          # NCHW planar, as the loader / model produce it
    n, c, h, w = x.shape
    out = torch.empty(n, h, w, device=x.device, dtype=torch.float32)
    mm = None
    if normalize:
        mm = torch.empty(2, device=x.device, dtype=torch.int32)
        mm[0].fill_(0x7F7FFFFF)
        mm[1].zero_()
    A.call("segsde_pixel_entropy", A.ptr(x), C.c_int(n), C.c_int(c), C.c_int(h), C.c_int(w), C.c_int(int(bool(normalize))),
           A.ptr(out), A.ptr(mm), A.stream_ptr())
    return out
