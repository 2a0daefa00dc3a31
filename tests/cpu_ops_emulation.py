"""TEST INFRASTRUCTURE ONLY — a torch-functional stand-in for the kernel-backed `ops` layer, so that the Python host
logic of the drop-in (`models/*.py`: encoder / decoder wiring, skip connections, exec_layer, PAD cross gating, pose
routing, module <-> state_dict mapping) can be exercised on a machine without a GPU.  Nothing in the product package
imports this file; the product path has no CPU route (ops.* raise SegsdeError on CPU tensors).

`install(monkeypatch)` swaps the entry points of `ops` that the model files call for the functions below and disables
the CUDA-only guard.  Each function restates the documented contract of the op it replaces (ops.py docstrings / the
C header), with torch.nn.functional.
"""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_ELU, ACT_SIGMOID = 0, 1, 2, 3
PAD_ZERO, PAD_REFLECT = 0, 1


def _act(x, act):
    return {ACT_NONE: lambda t: t, ACT_RELU: F.relu, ACT_ELU: F.elu, ACT_SIGMOID: torch.sigmoid}[int(act)](x)


def conv2d(x1, weight, bias=None, x2=None, stride=1, pad=0, dil=1, pad_mode=PAD_ZERO, up1=False, act=ACT_NONE,
           nchw_norm_in=False, bn_stats=None):
    x = F.interpolate(x1, scale_factor=2, mode="nearest") if up1 else x1
    if x2 is not None:
        x = torch.cat([x, x2], 1)
    if nchw_norm_in:
        x = (x - 0.45) / 0.225
    if pad_mode == PAD_REFLECT and pad > 0:
        y = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), weight, bias, stride, 0, dil)
    else:
        y = F.conv2d(x, weight, bias, stride, pad, dil)
    if bn_stats is not None:          # (sum, sum of squares, shift = 0) per channel of the pre-activation output
        c = y.shape[1]
        with torch.no_grad():
            yd = y.detach().double()
            bn_stats[:c] = yd.sum((0, 2, 3))
            bn_stats[c:2 * c] = (yd * yd).sum((0, 2, 3))
            bn_stats[2 * c:3 * c] = 0
    return _act(y, act)


def batch_norm(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5, residual=None,
               act=ACT_NONE, sums=None):
    y = F.batch_norm(x, running_mean, running_var, weight, bias, bool(training), 0.0 if momentum is None else momentum, eps)
    if residual is not None:
        y = y + residual
    return _act(y, act)


def maxpool3x3s2(x):
    return F.max_pool2d(x, 3, 2, 1)


def spatial_mean(x, scale=1.0):
    return scale * x.mean((2, 3), keepdim=True)


def broadcast_hw(x, h, w):
    return x.expand(-1, -1, h, w)


def cat_channels(xs):
    return torch.cat(list(xs), 1)


def add(a, b):
    return a + b


def gate(features, attention):
    return features * torch.sigmoid(attention)


def bilinear(x, size, align_corners=False):
    return F.interpolate(x, size=(int(size[0]), int(size[1])), mode="bilinear", align_corners=bool(align_corners))


def activation(x, act):
    return _act(x, act)


def dropout(x, p, training, channelwise=False, seed=0, replay_mask=None):
    if not training or p <= 0.0:
        return x
    if replay_mask is not None:
        return x * replay_mask.to(x.dtype) / (1.0 - p)
    return F.dropout2d(x, p, True) if channelwise else F.dropout(x, p, True)


def pose_matrix(vec6, invert):
    """[B,6] (axis-angle, translation) -> [B,4,4]: T @ R, or R^T @ (-T) when invert (monodepth_layers.py:30-47)."""
    aa, t = vec6[:, :3], vec6[:, 3:]
    B = vec6.shape[0]
    angle = aa.norm(dim=1, keepdim=True)
    axis = aa / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[:, 0:1], axis[:, 1:2], axis[:, 2:3]
    rows = [torch.cat([x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa], 1),
            torch.cat([x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa], 1),
            torch.cat([z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca], 1)]
    R = torch.zeros(B, 4, 4, dtype=vec6.dtype)
    R[:, :3, :3] = torch.stack(rows, 1)
    R[:, 3, 3] = 1
    if invert:
        R, t = R.transpose(1, 2), -t
    T = torch.eye(4, dtype=vec6.dtype).repeat(B, 1, 1)
    T[:, :3, 3] = t
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def cross_entropy(logits, target, pixel_weights=None, ignore_index=250):
    c = logits.shape[1]
    flat = logits.permute(0, 2, 3, 1).reshape(-1, c)
    loss = F.cross_entropy(flat, target.reshape(-1), reduction="mean" if pixel_weights is None else "none",
                           ignore_index=ignore_index)
    if pixel_weights is not None:
        loss = (pixel_weights.reshape(-1).detach() * loss).mean()
    return loss


NAMES = ("conv2d", "batch_norm", "maxpool3x3s2", "spatial_mean", "broadcast_hw", "cat_channels", "add", "gate", "bilinear",
         "activation", "dropout", "pose_matrix", "cross_entropy")


def install(monkeypatch):
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi, ops
    for n in NAMES:
        monkeypatch.setattr(ops, n, globals()[n])
    monkeypatch.setattr(_cabi, "require_cuda", lambda *ts: None)
