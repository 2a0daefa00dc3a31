"""torch.autograd glue over the C-ABI kernels (include/segsde_b200.h).

Tensors cross this layer with the reference's NCHW *shape* but channels-last *memory* (channel
stride 1): the kernels see strided NHWC views, user code sees the shapes it expects.  No op here
computes anything with torch: torch only owns memory (caching allocator), streams and the autograd
graph.  Every op raises if it is handed a CPU tensor — there is no fallback.
"""
import ctypes as C
import os
import threading

import torch

from . import _cabi as A

# route eligible convolutions to the tcgen05 tensor-core kernels (set False to force the generic path)
USE_TC = True
# SEGSDE_CHECK_LABELS=1: cross_entropy raises on labels outside [0, C) that are not the ignore index (torch raises a
# device assert there); off by default because reading the counter synchronises the device
CHECK_LABELS = os.environ.get("SEGSDE_CHECK_LABELS", "0") == "1"


# when a list, every convolution launch is bracketed by CUDA events: (kind, algorithmic flops, ev0, ev1)
PROFILE = None


PROFILE_DESC = None   # optional parallel list of human-readable shapes (profiling scripts)
# when a list, every convolution launch appends (kind, route): route = "tc:<kernel>" (kernel = conv | conv256 | rowhalo |
# wgrad | wgrad3x3, the tcgen05 family), "cc:fewcout" (the heads' HBM-bound 1x1 GEMMs, deliberately on CUDA cores) or
# "generic" (CUDA-core fallback for shapes outside the family) — tests assert on it
ROUTES = None
_TC_KERNELS = {1: "conv", 2: "rowhalo", 3: "wgrad", 4: "wgrad3x3", 5: "conv256"}


def log_route(kind, tc, name=None):
    """name: a deliberate CUDA-core route ("cc:fewcout": the HBM-bound few-output-channel 1x1 GEMMs of the heads) as
    opposed to "generic", the fallback for shapes the tcgen05 family refuses."""
    if ROUTES is not None:
        ROUTES.append((kind, name or ("tc:" + _TC_KERNELS.get(A.lib().segsde_tc_last_kernel(), "?") if tc else "generic")))


def _timed(kind, flops, fn, desc=None):
    if PROFILE is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    PROFILE.append((kind, flops, e0, e1))
    if PROFILE_DESC is not None:
        PROFILE_DESC.append(desc)
    return r


# -------------------------------------------------------------------------------------------------
# layout helpers
# -------------------------------------------------------------------------------------------------
def cl_empty(n, c, h, w, device, zero=False):
    f = torch.zeros if zero else torch.empty
    return f((n, h, w, c), device=device, dtype=torch.float32).permute(0, 3, 1, 2)


class _ZeroPool:
    """Bump allocator over pre-zeroed chunks: accumulators that kernels fill with atomics (BatchNorm statistics,
    weight / bias gradients) are carved out of one memset chunk instead of costing a fill kernel each.  Every slice is
    handed out exactly once; a chunk is released by the caching allocator when its last slice dies.  Requests come
    from the forward thread and from autograd's backward thread, hence the lock."""

    def __init__(self, dtype, chunk_elems, align_elems):
        self.dtype, self.chunk, self.align = dtype, chunk_elems, align_elems
        self.state = {}                       # device -> [chunk tensor, elements used]
        self.lock = threading.Lock()

    def take(self, n, device):
        n_al = (n + self.align - 1) // self.align * self.align
        if n_al > self.chunk // 4:            # large requests get their own allocation
            return torch.zeros(n, device=device, dtype=self.dtype)
        with self.lock:
            st = self.state.get(device)
            if st is None or st[1] + n_al > self.chunk:
                st = [torch.zeros(self.chunk, device=device, dtype=self.dtype), 0]
                self.state[device] = st
            out = st[0][st[1]:st[1] + n]
            st[1] += n_al
        return out


_ZPOOL_CHUNK = 1 << 20        # float64 elements per chunk (8 MB)
_ZPOOL32_CHUNK = 1 << 25      # float32 elements per chunk (128 MB): about one training step's weight gradients
_POOL64 = _ZeroPool(torch.float64, _ZPOOL_CHUNK, 2)       # 16-byte aligned slices
_POOL32 = _ZeroPool(torch.float32, _ZPOOL32_CHUNK, 64)    # 256-byte aligned slices
_ZPOOL, _ZPOOL32 = _POOL64.state, _POOL32.state


def zeros_f64(n, device):
    """Zero-filled float64 accumulator (BatchNorm statistics / reductions) from the pool: one memset per ~1M elements
    instead of a fill kernel per layer."""
    return _POOL64.take(n, device)


def zeros_f32(n, device):
    """Zero-filled float32 accumulator (weight / bias / BatchNorm-parameter gradients) from the pool — the ~200
    per-layer fill kernels of a training step become one memset."""
    return _POOL32.take(n, device)


def zeros_like_w(w):
    """Zero gradient buffer for an OIHW-shaped weight held in channels_last (= OHWI) memory."""
    o, i, kh, kw = w.shape
    return zeros_f32(w.numel(), w.device).view(o, kh, kw, i).permute(0, 3, 1, 2)


def is_cl(x):
    return x.dim() == 4 and (x.shape[1] == 1 or x.stride(1) == 1) and x.dtype == torch.float32


def as_cl(x):
    """Returns x with channel stride 1 (converting with the layout kernel when necessary)."""
    A.require_cuda(x)
    if x.dtype != torch.float32:
        x = x.float()
    if is_cl(x) and all(s >= 0 for s in x.stride()):
        return x
    n, c, h, w = x.shape
    y = cl_empty(n, c, h, w, x.device)
    A.call("segsde_nchw_to_nhwc", A.ptr(x), C.c_int64(x.stride(0)), C.c_int64(x.stride(1)),
           C.c_int64(x.stride(2)), C.c_int64(x.stride(3)), C.byref(view(y)), A.stream_ptr())
    return y


def view(x, null=False):
    """NHWC view struct of an NCHW-shaped tensor with channel stride 1."""
    n, c, h, w = x.shape
    v = A.NHWC()
    v.ptr = None if null else x.data_ptr()
    v.n, v.h, v.w, v.c = n, h, w, c
    v.sn, v.sh, v.sw = x.stride(0), x.stride(2), x.stride(3)
    return v


def _ref(v):
    return C.byref(v) if v is not None else None


def ohwi(w):
    """Weight as a dense [O][kh][kw][I] block: an OIHW tensor in channels_last memory."""
    if w.permute(0, 2, 3, 1).is_contiguous():
        return w
    return w.contiguous(memory_format=torch.channels_last)


# -------------------------------------------------------------------------------------------------
# convolution
# -------------------------------------------------------------------------------------------------
def _desc(kh, kw, stride, pad, dil, pad_mode, up1, act, nchw, stride_w=0):
    d = A.ConvDesc()
    d.kh, d.kw, d.stride, d.pad, d.dil = kh, kw, stride, pad, dil
    d.pad_mode, d.up1, d.act, d.nchw_norm_in, d.stride_w = pad_mode, int(up1), act, int(nchw), stride_w
    return d


def _conv_out_hw(hc, wc, kh, kw, stride, pad, dil):
    return ((hc + 2 * pad - dil * (kh - 1) - 1) // stride + 1,
            (wc + 2 * pad - dil * (kw - 1) - 1) // stride + 1)


# (the convolution autograd op itself lives in conv_op.py; `conv2d` at the end of this module forwards to it)


# -------------------------------------------------------------------------------------------------
# batch norm (+ residual + ReLU)
# -------------------------------------------------------------------------------------------------
class _BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, training, momentum, eps, act, sums=None):
        x = as_cl(x)
        residual = as_cl(residual) if residual is not None else None
        n, c, h, w = x.shape
        dev, st = x.device, A.stream_ptr()
        mean = torch.empty(c, device=dev, dtype=torch.float32)
        invstd = torch.empty(c, device=dev, dtype=torch.float32)
        y = cl_empty(n, c, h, w, dev)
        gw = weight.detach() if weight is not None else None
        gb = bias.detach() if bias is not None else None
        vres = _ref(view(residual)) if residual is not None else None
        if training:
            if sums is None:       # not already produced by the convolution epilogue
                sums = zeros_f64(3 * c, dev)
                A.call("segsde_bn_stats", C.byref(view(x)), A.ptr(sums), st)
            # finalize (mean / invstd / running statistics) + normalise in one launch
            A.call("segsde_bn_apply_train", C.byref(view(x)), A.ptr(sums), C.c_int64(n * h * w), C.c_float(eps),
                   C.c_float(momentum), A.ptr(gw), A.ptr(gb), vres, C.byref(view(y)), C.c_int(act), A.ptr(mean),
                   A.ptr(invstd), A.ptr(running_mean), A.ptr(running_var), st)
        else:
            A.call("segsde_bn_eval_prepare", A.ptr(running_mean), A.ptr(running_var), C.c_int(c), C.c_float(eps),
                   A.ptr(mean), A.ptr(invstd), st)
            A.call("segsde_bn_apply", C.byref(view(x)), A.ptr(mean), A.ptr(invstd), A.ptr(gw), A.ptr(gb), vres,
                   C.byref(view(y)), C.c_int(act), st)
        # ReLU mask for the backward pass: with a residual it needs the saved output; without one the kernels recompute
        # it from x (one tensor read less in each of the two backward passes)
        ctx.save_for_backward(x, y if (act == A.ACT_RELU and residual is not None) else None, mean, invstd, gw, gb)
        ctx.cfg = (training, act, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, invstd, gw, gb = ctx.saved_tensors
        training, act, has_res = ctx.cfg
        dy = as_cl(dy)
        n, c, h, w = x.shape
        dev, st = x.device, A.stream_ptr()
        red = zeros_f64(2 * c, dev)
        vy = view(y) if y is not None else None
        A.call("segsde_bn_bwd_reduce", C.byref(view(x)), _ref(vy), C.byref(view(dy)), A.ptr(mean), A.ptr(invstd),
               A.ptr(gw), A.ptr(gb), C.c_int(act), A.ptr(red), st)
        need_x, need_w, need_b, need_r = ctx.needs_input_grad[:4]
        dx = cl_empty(n, c, h, w, dev) if need_x else None
        dres = cl_empty(n, c, h, w, dev) if (has_res and need_r) else None
        dgamma = zeros_f32(c, dev) if need_w else None
        dbeta = zeros_f32(c, dev) if need_b else None
        A.call("segsde_bn_bwd_apply", C.byref(view(x)), _ref(vy), C.byref(view(dy)), A.ptr(mean), A.ptr(invstd),
               A.ptr(gw), A.ptr(gb), C.c_int(act), C.c_int(1 if training else 0), A.ptr(red), C.c_int64(n * h * w),
               _ref(view(dx)) if dx is not None else None, _ref(view(dres)) if dres is not None else None,
               A.ptr(dgamma), A.ptr(dbeta), st)
        return dx, dgamma, dbeta, dres, None, None, None, None, None, None, None


def batch_norm(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5, residual=None,
               act=A.ACT_NONE, sums=None):
    """training=True: batch statistics (running buffers, when given, are updated in place);
    training=False: normalise with the running buffers."""
    return _BatchNormFn.apply(x, weight, bias, residual, running_mean, running_var, bool(training),
                              0.0 if momentum is None else momentum, eps, act, sums)


# -------------------------------------------------------------------------------------------------
# pooling / resampling / elementwise
# -------------------------------------------------------------------------------------------------
class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = as_cl(x)
        n, c, h, w = x.shape
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        y = cl_empty(n, c, ho, wo, x.device)
        idx = torch.empty(n * ho * wo * c, device=x.device, dtype=torch.uint8)
        A.call("segsde_maxpool3x3s2_fwd", C.byref(view(x)), C.byref(view(y)), A.ptr(idx), A.stream_ptr())
        ctx.save_for_backward(idx)
        ctx.in_shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        dy = as_cl(dy)
        dx = cl_empty(*ctx.in_shape, dy.device)
        A.call("segsde_maxpool3x3s2_bwd", C.byref(view(dy)), A.ptr(idx), C.byref(view(dx)), A.stream_ptr())
        return dx


def maxpool3x3s2(x):
    return _MaxPoolFn.apply(x)


class _SpatialMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        x = as_cl(x)
        n, c, h, w = x.shape
        y = cl_empty(n, c, 1, 1, x.device)
        A.call("segsde_spatial_mean_fwd", C.byref(view(x)), C.byref(view(y)), C.c_float(scale), A.stream_ptr())
        ctx.in_shape, ctx.scale = x.shape, scale
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_cl(dy)
        dx = cl_empty(*ctx.in_shape, dy.device)
        A.call("segsde_spatial_mean_bwd", C.byref(view(dy)), C.byref(view(dx)), C.c_float(ctx.scale), A.stream_ptr())
        return dx, None


def spatial_mean(x, scale=1.0):
    """[N,C,H,W] -> [N,C,1,1] = scale * mean over H,W."""
    return _SpatialMeanFn.apply(x, scale)


class _BroadcastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h, w):
        x = as_cl(x)
        n, c = x.shape[:2]
        y = cl_empty(n, c, h, w, x.device)
        A.call("segsde_broadcast_hw_fwd", C.byref(view(x)), C.byref(view(y)), A.stream_ptr())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_cl(dy)
        n, c = dy.shape[:2]
        dx = cl_empty(n, c, 1, 1, dy.device)
        A.call("segsde_broadcast_hw_bwd", C.byref(view(dy)), C.byref(view(dx)), A.stream_ptr())
        return dx, None, None


def broadcast_hw(x, h, w):
    """Bilinear resize from 1x1 (ASPPPooling) == broadcast."""
    return _BroadcastFn.apply(x, h, w)


class _CatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        xs = [as_cl(x) for x in xs]
        n, _, h, w = xs[0].shape
        cs = [x.shape[1] for x in xs]
        y = cl_empty(n, sum(cs), h, w, xs[0].device)
        st, c0 = A.stream_ptr(), 0
        for x, c in zip(xs, cs):
            A.call("segsde_copy_nhwc", C.byref(view(x)), C.byref(view(y[:, c0:c0 + c])), st)
            c0 += c
        ctx.cs = cs
        return y

    @staticmethod
    def backward(ctx, dy):
        out, c0 = [], 0
        for c in ctx.cs:
            out.append(dy[:, c0:c0 + c])
            c0 += c
        return tuple(out)


def cat_channels(xs):
    return _CatFn.apply(*xs)


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = as_cl(a), as_cl(b)
        y = cl_empty(*a.shape, a.device)
        A.call("segsde_add", C.byref(view(a)), C.byref(view(b)), C.byref(view(y)), A.stream_ptr())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return _AddFn.apply(a, b)


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, a):
        f, a = as_cl(f), as_cl(a)
        y = cl_empty(*f.shape, f.device)
        A.call("segsde_gate_fwd", C.byref(view(f)), C.byref(view(a)), C.byref(view(y)), A.stream_ptr())
        ctx.save_for_backward(f, a)
        return y

    @staticmethod
    def backward(ctx, dy):
        f, a = ctx.saved_tensors
        dy = as_cl(dy)
        df, da = cl_empty(*f.shape, f.device), cl_empty(*f.shape, f.device)
        A.call("segsde_gate_bwd", C.byref(view(f)), C.byref(view(a)), C.byref(view(dy)), C.byref(view(df)),
               C.byref(view(da)), A.stream_ptr())
        return df, da


def gate(features, attention):
    """features * sigmoid(attention) (SelfAttention, model_parts.py:43-45)."""
    return _GateFn.apply(features, attention)


class _BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h, w, align):
        x = as_cl(x)
        n, c = x.shape[:2]
        y = cl_empty(n, c, h, w, x.device)
        A.call("segsde_bilinear_fwd", C.byref(view(x)), C.byref(view(y)), C.c_int(int(align)), A.stream_ptr())
        ctx.in_shape, ctx.align = x.shape, align
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_cl(dy)
        dx = cl_empty(*ctx.in_shape, dy.device, zero=True)
        A.call("segsde_bilinear_bwd", C.byref(view(dy)), C.byref(view(dx)), C.c_int(int(ctx.align)), A.stream_ptr())
        return dx, None, None, None


def bilinear(x, size, align_corners=False):
    return _BilinearFn.apply(x, int(size[0]), int(size[1]), bool(align_corners))


class _ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        x = as_cl(x)
        y = cl_empty(*x.shape, x.device)
        A.call("segsde_act_fwd", C.byref(view(x)), C.byref(view(y)), C.c_int(act), A.stream_ptr())
        ctx.save_for_backward(y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = as_cl(dy)
        dz = cl_empty(*dy.shape, dy.device)
        A.call("segsde_act_bwd", C.byref(view(y)), C.byref(view(dy)), C.byref(view(dz)), C.c_int(ctx.act), A.stream_ptr())
        return dz, None


def activation(x, act):
    return _ActFn.apply(x, act)


_DROPOUT_STEP = [0]


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, channelwise, seed, replay_mask):
        x = as_cl(x)
        n, c, h, w = x.shape
        y = cl_empty(n, c, h, w, x.device)
        mask = torch.empty(n * c if channelwise else n * c * h * w, device=x.device, dtype=torch.uint8)
        _DROPOUT_STEP[0] += 1
        rm = replay_mask.to(x.device, torch.float32).contiguous() if replay_mask is not None else None
        A.call("segsde_dropout_fwd", C.byref(view(x)), C.byref(view(y)), C.c_float(p), C.c_uint64(seed),
               C.c_uint64(_DROPOUT_STEP[0]), A.ptr(rm), A.ptr(mask), C.c_int(int(channelwise)), A.stream_ptr())
        ctx.save_for_backward(mask)
        ctx.cfg = (p, channelwise)
        return y

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        p, channelwise = ctx.cfg
        dy = as_cl(dy)
        dx = cl_empty(*dy.shape, dy.device)
        A.call("segsde_dropout_bwd", C.byref(view(dy)), A.ptr(mask), C.c_float(p), C.c_int(int(channelwise)),
               C.byref(view(dx)), A.stream_ptr())
        return dx, None, None, None, None


def dropout(x, p, training, channelwise=False, seed=None, replay_mask=None):
    """seed=None: A.default_seed (torch.manual_seed + rank); the per-call counter separates layers and steps."""
    if not training or p <= 0.0:
        return x
    if seed is None:
        seed = A.default_seed(0xD20B007)
    return _DropoutFn.apply(x, float(p), channelwise, seed, replay_mask)


# -------------------------------------------------------------------------------------------------
# pose geometry
# -------------------------------------------------------------------------------------------------
class _PoseMatrixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, invert):
        A.require_cuda(vec)
        v = vec.detach().reshape(-1, 6).contiguous().float()
        b = v.shape[0]
        M = torch.empty(b, 4, 4, device=v.device, dtype=torch.float32)
        jac = torch.empty(b, 12, 6, device=v.device, dtype=torch.float32)
        A.call("segsde_pose_matrix_fwd", A.ptr(v), C.c_int(b), C.c_int(int(invert)), A.ptr(M), A.ptr(jac), A.stream_ptr())
        ctx.save_for_backward(jac)
        ctx.in_shape = vec.shape
        return M

    @staticmethod
    def backward(ctx, dM):
        (jac,) = ctx.saved_tensors
        b = jac.shape[0]
        dM = dM.contiguous().float()
        dv = torch.empty(b, 6, device=dM.device, dtype=torch.float32)
        A.call("segsde_pose_matrix_bwd", A.ptr(jac), A.ptr(dM), C.c_int(b), A.ptr(dv), A.stream_ptr())
        return dv.view(ctx.in_shape), None


def pose_matrix(vec6, invert):
    """vec6: [B,6] (axis-angle, translation) -> [B,4,4] (monodepth_layers.py:30-47)."""
    return _PoseMatrixFn.apply(vec6, invert)


# -------------------------------------------------------------------------------------------------
# cross entropy
# -------------------------------------------------------------------------------------------------
class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, pixel_weights, ignore_index):
        logits = as_cl(logits)
        A.require_cuda(target)
        if target.dtype != torch.int64:        # the kernel reads int64 labels (the loader's dtype)
            target = target.long()
        target = target.contiguous()
        if target.numel() != logits.shape[0] * logits.shape[2] * logits.shape[3]:
            raise ValueError("cross_entropy: target %s does not match logits %s" % (tuple(target.shape), tuple(logits.shape)))
        pw = pixel_weights.detach().contiguous().float() if pixel_weights is not None else None
        if pw is not None and pw.numel() != target.numel():
            raise ValueError("cross_entropy: pixel_weights %s do not match target %s" % (tuple(pw.shape), tuple(target.shape)))
        acc = torch.zeros(4, device=logits.device, dtype=torch.float32)
        A.call("segsde_ce_fwd", C.byref(view(logits)), A.ptr(target), A.ptr(pw), C.c_int(ignore_index),
               A.ptr(acc), A.stream_ptr())
        ctx.save_for_backward(logits, target, pw, acc)
        ctx.ignore_index = ignore_index
        return acc

    @staticmethod
    def backward(ctx, gacc):
        # gacc[0] = d loss / d (sum of NLL); the caller divides by the count through _ScalarDivFn
        logits, target, pw, acc = ctx.saved_tensors
        g = gacc.contiguous()
        dl = cl_empty(*logits.shape, logits.device)
        A.call("segsde_ce_bwd", C.byref(view(logits)), A.ptr(target), A.ptr(pw), C.c_int(ctx.ignore_index),
               A.ptr(g), A.ptr(acc), C.byref(view(dl)), A.stream_ptr())
        return dl, None, None, None


class _RatioFn(torch.autograd.Function):
    """loss = acc[0] / den where den is acc[1] (valid count, mean reduction) or a host constant."""

    @staticmethod
    def forward(ctx, acc, const_den):
        out = torch.empty((), device=acc.device, dtype=torch.float32)
        inv = torch.empty(2, device=acc.device, dtype=torch.float32)
        A.call("segsde_ratio", A.ptr(acc), C.c_float(const_den), A.ptr(out), A.ptr(inv), A.stream_ptr())
        ctx.save_for_backward(inv)
        return out

    @staticmethod
    def backward(ctx, g):
        (inv,) = ctx.saved_tensors
        out = torch.zeros(4, device=inv.device, dtype=torch.float32)
        A.call("segsde_scale_by_dev", A.ptr(inv), A.ptr(g.contiguous()), C.c_float(1.0), None, C.c_float(0.0),
               A.ptr(out), C.c_int(0), C.c_int64(1), A.stream_ptr())
        return out, None


def cross_entropy(logits, target, pixel_weights=None, ignore_index=250):
    """Mean NLL over valid pixels, or mean over all pixels of w*NLL when pixel_weights is given."""
    acc = _CrossEntropyFn.apply(logits, target, pixel_weights, ignore_index)
    if CHECK_LABELS and float(acc[3]) > 0:          # debug switch: costs a device synchronisation
        raise ValueError("cross_entropy: %d labels outside [0, %d) that are not ignore_index %d"
                         % (int(acc[3]), logits.shape[1], ignore_index))
    den = 0.0 if pixel_weights is None else float(target.numel())
    return _RatioFn.apply(acc, den)


# the convolution op proper lives in conv_op.py (tensor-core routing, halo preparation, dgrad-as-fprop), which itself
# imports this module: resolved at call time so that either module can be imported first
def conv2d(x1, weight, bias=None, x2=None, **kwargs):
    """y = act(conv(cat(up?(x1), x2)) + bias) — see conv_op.conv2d for the keywords."""
    from .conv_op import conv2d as impl
    return impl(x1, weight, bias, x2, **kwargs)
