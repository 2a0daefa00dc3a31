"""nn.Module shells around the sm_100a kernels.  They subclass the torch modules the reference uses
(so `isinstance(m, nn.BatchNorm2d)` checks such as train.py:433-440 keep working and the
state_dict keys / OIHW shapes are identical) but never call torch compute ops."""
import os

import torch
from torch import nn

from .. import _cabi as A
from .. import ops


class Conv2d(nn.Conv2d):
    """nn.Conv2d whose weight lives in channels_last memory (= the kernels' OHWI) and whose forward
    is `ops.conv2d`.  Extra forward arguments expose the fusions of the decoder."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.groups != 1 or self.padding_mode != "zeros":
            raise NotImplementedError("segsde_b200 Conv2d: groups=1 and zero padding_mode only")
        assert self.stride[0] == self.stride[1] and self.dilation[0] == self.dilation[1]
        assert self.padding[0] == self.padding[1]
        with torch.no_grad():
            self.weight.data = self.weight.data.contiguous(memory_format=torch.channels_last)

    def forward(self, x, x2=None, up1=False, act=A.ACT_NONE, pad=None, pad_mode=A.PAD_ZERO, nchw_norm_in=False,
                bn_stats=None):
        return ops.conv2d(x, self.weight, self.bias, x2=x2, stride=self.stride[0],
                          pad=self.padding[0] if pad is None else pad, dil=self.dilation[0], pad_mode=pad_mode,
                          up1=up1, act=act, nchw_norm_in=nchw_norm_in, bn_stats=bn_stats)


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d on the fused kernels: y = act(bn(x) [+ residual])."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._pending_batches = 0     # num_batches_tracked is flushed lazily (no per-step tiny kernel)

    def uses_batch_stats(self):
        return self.training or not self.track_running_stats

    def forward(self, x, residual=None, act=A.ACT_NONE, sums=None):
        use_batch_stats = self.training or not self.track_running_stats
        update = self.training and self.track_running_stats
        momentum = self.momentum
        if update:
            self._pending_batches += 1
            if momentum is None:      # cumulative moving average
                momentum = 1.0 / float(int(self.num_batches_tracked) + self._pending_batches)
        if use_batch_stats:           # running buffers are only passed when they must be updated
            rm, rv = (self.running_mean, self.running_var) if update else (None, None)
        else:
            rm, rv = self.running_mean, self.running_var
        return ops.batch_norm(x, self.weight, self.bias, rm, rv, use_batch_stats, momentum or 0.0, self.eps,
                              residual=residual, act=act, sums=sums if use_batch_stats else None)

    def _flush(self):
        if self._pending_batches and self.num_batches_tracked is not None:
            self.num_batches_tracked += self._pending_batches
        self._pending_batches = 0

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        self._flush()
        super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, *args, **kwargs):
        self._pending_batches = 0
        super()._load_from_state_dict(*args, **kwargs)


# eval-mode inference (no autograd): fold BatchNorm's running statistics into the convolution weights so that
# conv -> BN -> ReLU is ONE convolution with a bias + ReLU epilogue (the reference's validate(),
# train.py:817-923, runs the same layers as conv + BN + ReLU kernels)
FOLD_EVAL_BN = os.environ.get("SEGSDE_FOLD_EVAL_BN", "1") != "0"


def fold_bn(conv, bn):
    """(w', b') with BatchNorm's eval-mode affine map folded into the convolution (segsde_bn_fold)."""
    import ctypes as C
    w = ops.ohwi(conv.weight.detach())
    cout = w.shape[0]
    wf = torch.empty_like(w)
    bf = torch.empty(cout, device=w.device, dtype=torch.float32)
    A.call("segsde_bn_fold", A.ptr(w), A.ptr(conv.bias.detach() if conv.bias is not None else None),
           A.ptr(bn.weight.detach() if bn.weight is not None else None),
           A.ptr(bn.bias.detach() if bn.bias is not None else None), A.ptr(bn.running_mean), A.ptr(bn.running_var),
           C.c_float(bn.eps), C.c_int(cout), C.c_int(w.numel() // cout), A.ptr(wf), A.ptr(bf), A.stream_ptr())
    return wf, bf


def conv_bn(conv, bn, x, residual=None, act=A.ACT_NONE, **conv_kw):
    """bn(conv(x)) [+ residual] [ReLU] with the BatchNorm batch statistics accumulated in the convolution's
    tensor-core epilogue (no separate pass over the conv output) when the BN layer normalises with batch statistics;
    in eval-mode inference the BatchNorm is folded into the convolution (no BatchNorm pass at all)."""
    if (FOLD_EVAL_BN and residual is None and not torch.is_grad_enabled() and isinstance(bn, BatchNorm2d)
            and isinstance(conv, Conv2d) and not bn.uses_batch_stats() and bn.running_mean is not None
            and not conv_kw.get("nchw_norm_in", False) and conv.weight.is_cuda):
        wf, bf = fold_bn(conv, bn)
        return ops.conv2d(x, wf, bf, stride=conv.stride[0], pad=conv.padding[0], dil=conv.dilation[0], act=act, **conv_kw)
    if (isinstance(bn, BatchNorm2d) and bn.uses_batch_stats() and isinstance(conv, Conv2d) and conv.bias is None
            and os.environ.get("SEGSDE_NO_BNFUSE", "0") != "1"):
        sums = ops.zeros_f64(3 * conv.out_channels, conv.weight.device)
        return bn(conv(x, bn_stats=sums, **conv_kw), residual=residual, act=act, sums=sums)
    return bn(conv(x, **conv_kw), residual=residual, act=act)


def _next_mask(m):
    """`replay_mask`: None, one NCHW 0/1 tensor (every call), or a list consumed one mask per training-mode call."""
    if isinstance(m.replay_mask, list):
        return m.replay_mask.pop(0) if (m.training and m.p > 0 and m.replay_mask) else None
    return m.replay_mask


class Dropout(nn.Dropout):
    """nn.Dropout on the Philox dropout kernel; `replay_mask` pins the mask(s) in parity tests."""
    replay_mask = None

    def forward(self, x):
        return ops.dropout(x, self.p, self.training, channelwise=False, replay_mask=_next_mask(self))


class Dropout2d(nn.Dropout2d):
    replay_mask = None

    def forward(self, x):
        return ops.dropout(x, self.p, self.training, channelwise=True, replay_mask=_next_mask(self))
