"""Batch-axis data parallelism: one process per GPU, model replicated, one NCCL all-reduce of the flat
fp32 gradient buffer per optimizer step over NVLink/NVSwitch.  BatchNorm statistics stay
per rank (the reference is single-process; this matches running it per GPU at the per-GPU batch).

The reference's step calls `.backward()` twice (train.py:486,510) and clips after (train.py:516-524): the
exchange therefore happens ONCE, in `reduce()`, after the last backward and before clipping / the optimizer
step.  `reduce()` issues one collective per bucket on a side stream so that bucket k+1's all-reduce overlaps
bucket k's 1/world scaling, and verifies first that every `.grad` still lives in the flat buffer —
`optimizer.zero_grad()` (set_to_none=True is torch's default) drops the views, after which autograd would
write into fresh tensors and the ranks would silently diverge."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _cabi as A


class GradSync:
    """Owns one flat fp32 buffer; every trainable parameter's .grad is a view into it (same strides as the
    parameter, so channels_last weights keep their layout).  `reduce()` = all_reduce(SUM) then 1/world.

    Use `sync.zero()` (or `optimizer.zero_grad(set_to_none=False)`) between steps.  If the views were dropped
    anyway, `reduce()` re-attaches them: a gradient that autograd meanwhile wrote into a private tensor is copied
    into its slot, a parameter without a gradient contributes zeros."""

    def __init__(self, params, process_group=None, bucket_bytes=64 << 20):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradSync: no trainable parameters")
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views, self.offsets = [], []
        off = 0
        for p in self.params:
            k = p.numel()
            self.views.append(self._view(p, off))
            self.offsets.append(off)
            off += k
        self.attach()
        # bucket boundaries on parameter boundaries, ~bucket_bytes each
        self.buckets, start = [], 0
        limit = max(1, bucket_bytes // 4)
        for o, p in zip(self.offsets, self.params):
            if o + p.numel() - start >= limit:
                self.buckets.append((start, o + p.numel()))
                start = o + p.numel()
        if start < n:
            self.buckets.append((start, n))
        self._side = None

    def _view(self, p, off):
        seg = self.flat[off:off + p.numel()]
        if p.dim() == 4 and p.permute(0, 2, 3, 1).is_contiguous():
            o, i, kh, kw = p.shape
            return seg.view(o, kh, kw, i).permute(0, 3, 1, 2)
        return seg.view(p.shape)

    def attach(self):
        """(Re-)bind every parameter's .grad to its slot of the flat buffer.  Returns the number re-bound."""
        fixed = 0
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is not None and g.data_ptr() == v.data_ptr() and g.stride() == v.stride():
                continue
            if g is None:
                v.zero_()
            else:
                v.copy_(g)
            p.grad = v
            fixed += 1
        return fixed

    def zero(self):
        self.attach()
        self.flat.zero_()

    def reduce(self):
        self.attach()
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world <= 1:
            return self.flat
        if not self.flat.is_cuda:       # gloo/CPU path exists for the host-logic tests only
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / world)
            return self.flat
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.flat.device)
        side = self._side
        side.wait_stream(main)
        done = []
        with torch.cuda.stream(side):
            for a, b in self.buckets:
                dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group)
                ev = torch.cuda.Event()
                ev.record(side)
                done.append(ev)
        inv = C.c_float(1.0 / world)
        for (a, b), ev in zip(self.buckets, done):       # scale bucket k while bucket k+1 is still on the wire
            main.wait_event(ev)
            seg = self.flat[a:b]
            A.call("segsde_axpby", A.ptr(seg), inv, A.ptr(seg), C.c_int(0), C.c_int64(b - a), A.stream_ptr())
        return self.flat
