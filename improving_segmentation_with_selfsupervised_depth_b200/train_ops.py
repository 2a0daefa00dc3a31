"""Step-level ops either side of the model / loss call — SURVEY.md §8(a) rows T1-T4.  In the reference these are
`Trainer` methods / inline code of train.py (not part of `models/` or `loss/`), so a maintainer switches them by
replacing the cited lines with the calls below.  Everything runs in the sm_100a kernels of
csrc/train_ops.cu; scalars the reference reads back with `.item()` stay on the device.
"""
import ctypes as C

import torch

from . import _cabi as A
from . import ops


# ---- T1 --------------------------------------------------------------------------------------------------------
class _FeatureDistanceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        A.require_cuda(a, b)
        if a.shape != b.shape:
            raise ValueError("feature_distance: shapes differ: %s vs %s" % (tuple(a.shape), tuple(b.shape)))
        # same memory order for both (the sum does not care which)
        a = a.float()
        b = b.float()
        if a.stride() != b.stride() or not (a.is_contiguous() or a.is_contiguous(memory_format=torch.channels_last)):
            a, b = a.contiguous(), b.contiguous()
        dist = torch.empty((), device=a.device, dtype=torch.float32)
        A.call("segsde_feature_distance_fwd", A.ptr(a), A.ptr(b), C.c_int64(a.numel()), A.ptr(ops.zeros_f64(1, a.device)),
               A.ptr(dist), A.stream_ptr())
        ctx.save_for_backward(a, b, dist)
        return dist

    @staticmethod
    def backward(ctx, g):
        a, b, dist = ctx.saved_tensors
        need_a, need_b = ctx.needs_input_grad
        da = torch.empty_like(a) if need_a else None
        db = torch.empty_like(b) if need_b else None
        if da is not None or db is not None:
            A.call("segsde_feature_distance_bwd", A.ptr(a), A.ptr(b), C.c_int64(a.numel()), A.ptr(dist),
                   A.ptr(g.contiguous().float()), A.ptr(da), A.ptr(db), A.stream_ptr())
        return da, db


def feature_distance(a, b):
    """`torch.dist(a, b, p=2)` of train.py:482 (ImageNet feature-distance loss, dec6)."""
    return _FeatureDistanceFn.apply(a, b)


# ---- T2 --------------------------------------------------------------------------------------------------------
def normalize_depths(disp):
    """train.py:688-692: per-sample min-max normalisation of the (detached) student disparities, B x 1 x H x W."""
    A.require_cuda(disp)
    d = disp.detach().contiguous().float()
    b = d.shape[0]
    hw = d.numel() // b
    mm = torch.empty(b, 2, device=d.device, dtype=torch.int32)
    mm[:, 0].fill_(-1)          # 0xffffffff
    mm[:, 1].zero_()
    out = torch.empty_like(d)
    A.call("segsde_sample_minmax_normalize", A.ptr(d), C.c_int(b), C.c_int64(hw), A.ptr(mm), A.ptr(out), A.stream_ptr())
    return out


def depthcomp_mix_mask(depths, margin, foreground_threshold, generator=None):
    """`generate_mix_mask` in "depthcomp" mode (train.py:585-604): sample i keeps the pixels where it is in front of
    sample (i+1) % B (disparity >= other - margin) and above the foreground threshold.  Returns int64 B x H x W.
    (The reference asserts B == 2 and pairs (0,1), (1,0) — the same pairing.)  For a (lower, upper) threshold pair
    one threshold PER SAMPLE is drawn on the device generator, sample 0 first, exactly the reference's RNG
    consumption (`torch.rand(1, device=...)` per image, train.py:594-598)."""
    A.require_cuda(depths)
    d = depths.detach().contiguous().float()
    b = d.shape[0]
    hw = d.numel() // b
    thr_dev, thr = None, 0.0
    if isinstance(foreground_threshold, (tuple, list)):
        lo, hi = foreground_threshold
        assert hi > lo
        thr_dev = torch.cat([torch.rand(1, device=d.device, generator=generator) * (hi - lo) + lo for _ in range(b)])
    else:
        thr = float(foreground_threshold)
    mask = torch.empty((b,) + tuple(d.shape[-2:]), device=d.device, dtype=torch.int64)
    A.call("segsde_depthcomp_mask", A.ptr(d), C.c_int(b), C.c_int64(hw), C.c_float(margin), C.c_float(thr),
           A.ptr(thr_dev), C.c_int(1), A.ptr(mask), None, A.stream_ptr())
    return mask


def depth_mix_mask(depths, min_depth=0.1, max_depth=0.4, generator=None):
    """`generate_mix_mask` in "depth" mode (train.py:605-615, loader/transformmasks.py:33-42 with a one-element
    threshold): mask[i] = [disparity_i >= t_i], t_i ~ U(min_depth, max_depth) drawn per sample on the device.
    Returns fp32 B x H x W like the reference (`.float()`)."""
    A.require_cuda(depths)
    d = depths.detach().contiguous().float()
    b = d.shape[0]
    hw = d.numel() // b
    thr = torch.cat([torch.rand(1, device=d.device, generator=generator) * (max_depth - min_depth) + min_depth
                     for _ in range(b)])
    mask = torch.empty((b,) + tuple(d.shape[-2:]), device=d.device, dtype=torch.float32)
    A.call("segsde_depthcomp_mask", A.ptr(d), C.c_int(b), C.c_int64(hw), C.c_float(0.0), C.c_float(0.0), A.ptr(thr),
           C.c_int(0), None, A.ptr(mask), A.stream_ptr())
    return mask


def generate_mix_mask(mode, depths, margin=0.03, foreground_threshold=0.0, generator=None):
    """Dispatch over the depth-based modes of `Trainer.generate_mix_mask` (train.py:566-640).  "class" and
    "depthhist" need host-side numpy sampling in the reference and are not part of the GPU path."""
    if mode == "depthcomp":
        return depthcomp_mix_mask(depths, margin, foreground_threshold, generator)
    if mode == "depth":
        return depth_mix_mask(depths, generator=generator)
    raise NotImplementedError("mix-mask mode %r (device path covers 'depthcomp' and 'depth')" % (mode,))


def _mix_one(mask, x):
    A.require_cuda(mask, x)
    if x.dim() != 4 or mask.shape[0] != x.shape[0]:
        raise NotImplementedError("mix: B x C x H x W data with a B x H x W mask (the branch the reference's step uses)")
    x = x.float()
    b, c, h, w = x.shape
    if not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)):
        x = x.contiguous()
    mask = mask.reshape(b, h, w).contiguous()
    out = torch.empty_like(x)            # preserves the memory format
    sp = x.stride(3)
    assert x.stride(2) == w * sp and out.stride() == x.stride()
    mi = mask if mask.dtype == torch.int64 else None
    mf = None if mi is not None else mask.float()
    A.call("segsde_mix", A.ptr(x), A.ptr(out), A.ptr(mi), A.ptr(mf), C.c_int(b), C.c_int(c), C.c_int64(h * w),
           C.c_int64(x.stride(0)), C.c_int64(x.stride(1)), C.c_int64(sp), C.c_int64(out.stride(0)),
           C.c_int64(out.stride(1)), C.c_int64(out.stride(3)), A.stream_ptr())
    return out


def mix(mask, data=None, target=None):
    """`transformsgpu.mix` (loader/transformsgpu.py:33-47): out[i] = mask[i] * x[i] + (1 - mask[i]) * x[(i+1) % B]."""
    if data is not None:
        data = _mix_one(mask, data)
    if target is not None:
        target = _mix_one(mask, target)
    return data, target


# ---- T3 --------------------------------------------------------------------------------------------------------
def softmax_channels(logits):
    """`torch.softmax(logits.detach(), dim=1)` of train.py:667 (teacher probabilities; no gradient).  The result keeps
    the memory format of the input."""
    A.require_cuda(logits)
    x = logits.detach().float()
    if x.dim() != 4:
        raise ValueError("softmax_channels: B x C x H x W logits expected")
    if not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)):
        x = x.contiguous()
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    A.call("segsde_softmax_channels", A.ptr(x), A.ptr(out), C.c_int(b), C.c_int(c), C.c_int64(h * w), C.c_int64(x.stride(0)),
           C.c_int64(x.stride(1)), C.c_int64(x.stride(3)), C.c_int64(out.stride(0)), C.c_int64(out.stride(1)),
           C.c_int64(out.stride(3)), A.stream_ptr())
    return out


def pseudo_labels(teacher_softmax, threshold=0.968, ignore_index=250, weight_scale=1.0):
    """train.py:645-648: (pseudo_label int64 B x H x W, pixel weights fp32 B x H x W) — the label is the arg-max
    class (ignore_index where the maximum is exactly 0), every pixel weight is weight_scale * share of pixels whose
    maximum reaches the threshold."""
    A.require_cuda(teacher_softmax)
    p = teacher_softmax.detach().float()
    b, c, h, w = p.shape
    if not (p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)):
        p = p.contiguous()
    label = torch.empty(b, h, w, device=p.device, dtype=torch.int64)
    weight = torch.empty(b, h, w, device=p.device, dtype=torch.float32)
    count = torch.zeros(1, device=p.device, dtype=torch.int64)
    A.call("segsde_pseudo_label", A.ptr(p), C.c_int(b), C.c_int(c), C.c_int64(h * w), C.c_int64(p.stride(0)),
           C.c_int64(p.stride(1)), C.c_int64(p.stride(3)), C.c_float(threshold), C.c_int64(ignore_index), A.ptr(label),
           None, A.ptr(count), C.c_float(weight_scale), A.ptr(weight), A.stream_ptr())
    return label, weight


def calc_pseudo_label_loss(teacher_softmax, student_logits, consistency_weight=1.0, ignore_index=250):
    """`Trainer.calc_pseudo_label_loss` (train.py:644-651): consistency_weight * cross_entropy2d(student, pseudo label,
    pixel_weights = confidence share).  Returns (loss, pseudo_label)."""
    from .loss.loss import cross_entropy2d
    label, weight = pseudo_labels(teacher_softmax, ignore_index=ignore_index, weight_scale=consistency_weight)
    if ignore_index != 250:
        raise NotImplementedError("cross_entropy2d ignores label 250 (loss/loss.py:27)")
    return cross_entropy2d(input=student_logits, target=label, pixel_weights=weight), label


# ---- T4 --------------------------------------------------------------------------------------------------------
def multi_axpby(dst, src, alpha, beta):
    """dst[i] = alpha * dst[i] + beta * src[i] over lists of dense fp32 CUDA tensors, in place."""
    dst, src = list(dst), list(src)
    if len(dst) != len(src):
        raise ValueError("multi_axpby: %d destination and %d source tensors" % (len(dst), len(src)))
    if not dst:
        return
    keep = []
    for d, s in zip(dst, src):
        A.require_cuda(d, s)
        if d.dtype != torch.float32 or d.numel() != s.numel():
            raise ValueError("multi_axpby: fp32 tensors of equal size expected")
        dense = d.is_contiguous() or (d.dim() == 4 and d.is_contiguous(memory_format=torch.channels_last))
        if not dense:
            raise ValueError("multi_axpby: destination must be dense")
        if s.dtype != torch.float32 or s.stride() != d.stride():
            s = s.detach().float().contiguous() if d.is_contiguous() else s.detach().float().contiguous(memory_format=torch.channels_last)
        keep.append(s)
    n = len(dst)
    dp = (C.c_void_p * n)(*[d.data_ptr() for d in dst])
    sp = (C.c_void_p * n)(*[s.data_ptr() for s in keep])
    ne = (C.c_int64 * n)(*[d.numel() for d in dst])
    A.call("segsde_multi_axpby", C.c_int(n), dp, sp, ne, C.c_float(alpha), C.c_float(beta), A.stream_ptr())


def update_ema_variables(ema_params, model_params, alpha_teacher, iteration):
    """`Trainer.update_ema_variables` (train.py:346-358): alpha = min(1 - 1/(iteration+1), alpha_teacher);
    ema = alpha * ema + (1 - alpha) * param for every parameter pair — one multi-tensor launch per 48 tensors
    instead of one tiny kernel chain per parameter."""
    alpha = min(1 - 1 / (iteration + 1), alpha_teacher)
    with torch.no_grad():
        multi_axpby([e.data for e in ema_params], [p.data for p in model_params], alpha, 1 - alpha)
    return alpha
