"""Optimizer step and gradient clipping on multi-tensor kernels.

The reference picks `torch.optim.Adam` / `SGD` by name (utils/optimizers.py:7-30, train.py:291-295) and clips with
`torch.nn.utils.clip_grad_norm_` (train.py:516-524).  `Adam` and `SGD` below take the same constructor arguments and
keep the same `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq` / `momentum_buffer`), so a maintainer switches
by registering them in `key2opt`; `clip_grad_norm_` has torch's signature and return value.  One launch per 36
tensors instead of several small kernels per parameter; no host synchronisation (the total norm stays on the device).
"""
import ctypes as C

import torch

from . import _cabi as A
from . import ops


def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _like(t, ref):
    """t with ref's memory order (gradients normally already share the parameter's strides)."""
    if t.stride() == ref.stride():
        return t
    out = torch.empty_like(ref)
    out.copy_(t)
    return out


def _collect(group_params):
    """Parameters with gradients, their gradients in the parameter's memory order."""
    ps, gs = [], []
    for p in group_params:
        if p.grad is None:
            continue
        A.require_cuda(p, p.grad)
        if p.dtype != torch.float32 or p.grad.is_sparse or not _dense(p):
            raise NotImplementedError("segsde optimizers handle dense fp32 CUDA parameters")
        ps.append(p)
        gs.append(_like(p.grad.detach().float(), p))
    return ps, gs


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam (amsgrad=False, maximize=False) on `segsde_multi_adam`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by the reference configs")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs = _collect(group["params"])
            if not ps:
                continue
            by_step = {}
            for p, g in zip(ps, gs):
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append((p, g, st))
            beta1, beta2 = group["betas"]
            for step, items in by_step.items():       # normally a single bucket: all parameters share the step count
                n = len(items)
                numel = (C.c_int64 * n)(*[p.numel() for p, _, _ in items])
                A.call("segsde_multi_adam", C.c_int(n), _ptr_array([p for p, _, _ in items]),
                       _ptr_array([g for _, g, _ in items]), _ptr_array([s["exp_avg"] for _, _, s in items]),
                       _ptr_array([s["exp_avg_sq"] for _, _, s in items]), numel, C.c_float(group["lr"]),
                       C.c_float(beta1), C.c_float(beta2), C.c_float(group["eps"]), C.c_float(group["weight_decay"]),
                       C.c_int64(step), A.stream_ptr())
        return loss


class SGD(torch.optim.Optimizer):
    """torch.optim.SGD (maximize=False) on `segsde_multi_sgd`."""

    def __init__(self, params, lr=1e-3, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("invalid SGD hyper-parameters")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs = _collect(group["params"])
            if not ps:
                continue
            mom = group["momentum"]
            buckets = {True: [], False: []}            # first use of a parameter: the buffer starts as the gradient
            for p, g in zip(ps, gs):
                st = self.state[p]
                first = mom != 0 and st.get("momentum_buffer") is None
                if first:
                    st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.preserve_format)
                buckets[first].append((p, g, st.get("momentum_buffer")))
            for first, items in buckets.items():
                if not items:
                    continue
                n = len(items)
                numel = (C.c_int64 * n)(*[p.numel() for p, _, _ in items])
                bufs = _ptr_array([b for _, _, b in items]) if mom != 0 else None
                A.call("segsde_multi_sgd", C.c_int(n), _ptr_array([p for p, _, _ in items]),
                       _ptr_array([g for _, g, _ in items]), bufs, numel, C.c_float(group["lr"]), C.c_float(mom),
                       C.c_float(group["dampening"]), C.c_float(group["weight_decay"]), C.c_int(int(group["nesterov"])),
                       C.c_int(int(first)), A.stream_ptr())
        return loss


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ (L2): scales every gradient by min(1, max_norm / (total_norm + 1e-6)) in place and
    returns the total norm as a 0-dim device tensor (no host synchronisation)."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("only the L2 norm (the reference's default, train.py:521-524)")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    for g in grads:
        A.require_cuda(g)
        if g.dtype != torch.float32 or not _dense(g):
            raise NotImplementedError("clip_grad_norm_: dense fp32 CUDA gradients")
    dev = grads[0].device
    n = len(grads)
    total, coef = torch.empty((), device=dev), torch.empty((), device=dev)
    numel = (C.c_int64 * n)(*[g.numel() for g in grads])
    A.call("segsde_multi_clip_grad_norm", C.c_int(n), _ptr_array(grads), numel, C.c_float(max_norm),
           A.ptr(ops.zeros_f64(1, dev)), A.ptr(total), A.ptr(coef), A.stream_ptr())
    return total


class GradScaler:
    """`torch.cuda.amp.GradScaler` for the calls the reference makes (train.py:336, 486-530): `scale(loss)`,
    `unscale_(optimizer)`, `step(optimizer)`, `update()`, plus `get_scale` / `state_dict` / `load_state_dict`.  The
    scale, the growth tracker and the found-inf flag live on the device; unscaling + the inf check of ALL gradients is
    one multi-tensor launch per 36 tensors.  `step` reads the flag back once (as torch's does) to decide whether to
    skip the optimizer.  The convolutions of this library compute in TF32 with fp32 storage under any autocast state
    (the custom Functions take fp32 inputs), so with `enabled=True` the scaler guards only against overflow of the
    scaled loss — it exists so that a `train.py` written for AMP runs unchanged."""

    def __init__(self, init_scale=2.0 ** 16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, enabled=True):
        self._enabled = bool(enabled)
        self._init_scale, self._growth_factor = float(init_scale), float(growth_factor)
        self._backoff_factor, self._growth_interval = float(backoff_factor), int(growth_interval)
        self._scale = self._tracker = self._found = None
        self._unscaled = set()

    def is_enabled(self):
        return self._enabled

    def _lazy(self, dev):
        if self._scale is None:
            self._scale = torch.full((1,), self._init_scale, device=dev, dtype=torch.float32)
            self._tracker = torch.zeros(1, device=dev, dtype=torch.int32)
            self._found = torch.zeros(1, device=dev, dtype=torch.float32)

    def scale(self, outputs):
        if not self._enabled:
            return outputs
        A.require_cuda(outputs)
        self._lazy(outputs.device)
        return outputs * self._scale.squeeze(0)

    def unscale_(self, optimizer):
        if not self._enabled:
            return
        if id(optimizer) in self._unscaled:
            raise RuntimeError("unscale_() has already been called on this optimizer since the last update().")
        grads = [p.grad for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
        if grads:
            for g in grads:
                A.require_cuda(g)
                if g.dtype != torch.float32 or not _dense(g):
                    raise NotImplementedError("GradScaler.unscale_: dense fp32 CUDA gradients")
            self._lazy(grads[0].device)
            n = len(grads)
            numel = (C.c_int64 * n)(*[g.numel() for g in grads])
            A.call("segsde_multi_unscale", C.c_int(n), _ptr_array(grads), numel, A.ptr(self._scale), A.ptr(self._found),
                   A.stream_ptr())
        self._unscaled.add(id(optimizer))

    def step(self, optimizer, *args, **kwargs):
        if not self._enabled:
            return optimizer.step(*args, **kwargs)
        if id(optimizer) not in self._unscaled:
            self.unscale_(optimizer)
        if self._found is not None and float(self._found) != 0.0:      # one read-back per step, like torch's GradScaler
            return None
        return optimizer.step(*args, **kwargs)

    def update(self, new_scale=None):
        if not self._enabled or self._scale is None:
            return
        if new_scale is not None:
            self._scale.fill_(float(new_scale))
        else:
            A.call("segsde_amp_update_scale", A.ptr(self._scale), A.ptr(self._tracker), A.ptr(self._found),
                   C.c_float(self._growth_factor), C.c_float(self._backoff_factor), C.c_int(self._growth_interval),
                   A.stream_ptr())
        self._found.zero_()
        self._unscaled.clear()

    def get_scale(self):
        if not self._enabled:
            return 1.0
        return self._init_scale if self._scale is None else float(self._scale)

    def state_dict(self):
        if not self._enabled:
            return {}
        return {"scale": self.get_scale(), "growth_factor": self._growth_factor, "backoff_factor": self._backoff_factor,
                "growth_interval": self._growth_interval,
                "_growth_tracker": 0 if self._tracker is None else int(self._tracker)}

    def load_state_dict(self, state):
        if not self._enabled:
            return
        self._init_scale = float(state["scale"])
        self._growth_factor, self._backoff_factor = float(state["growth_factor"]), float(state["backoff_factor"])
        self._growth_interval = int(state["growth_interval"])
        if self._scale is not None:
            self._scale.fill_(self._init_scale)
            self._tracker.fill_(int(state["_growth_tracker"]))
