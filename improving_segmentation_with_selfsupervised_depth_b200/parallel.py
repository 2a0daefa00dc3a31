"""Batch-axis data parallelism: one process per GPU, model replicated, one NCCL all-reduce of the flat
fp32 gradient buffer per optimizer step over NVLink/NVSwitch.  BatchNorm statistics stay
per rank (the reference is single-process; this matches running it per GPU at the per-GPU batch)."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _cabi as A


class GradSync:
    """Owns one flat fp32 buffer; every trainable parameter's .grad is a view into it (same strides as the
    parameter, so channels_last weights keep their layout).  `reduce()` = all_reduce(SUM) then 1/world."""

    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            k = p.numel()
            seg = self.flat[off:off + k]
            if p.dim() == 4 and p.permute(0, 2, 3, 1).is_contiguous():
                o, i, kh, kw = p.shape
                g = seg.view(o, kh, kw, i).permute(0, 3, 1, 2)
            else:
                g = seg.view(p.shape)
            p.grad = g
            off += k

    def zero(self):
        self.flat.zero_()

    def reduce(self):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.flat.is_cuda:
                A.call("segsde_axpby", A.ptr(self.flat), C.c_float(1.0 / world), A.ptr(self.flat), C.c_int(0),
                       C.c_int64(self.flat.numel()), A.stream_ptr())
            else:       # gloo/CPU path exists for the host-logic tests only
                self.flat.mul_(1.0 / world)
        return self.flat
