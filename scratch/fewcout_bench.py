"""Timing of the 19-class 1x1 segmentation heads (conv_fewcout.cu) at the joint configuration's shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
import ctypes as C
dev = torch.device('cuda')
def ev():
    return torch.cuda.Event(enable_timing=True)
for (B, cin, co, H, W) in [(8, 64, 19, 512, 1024), (8, 128, 19, 128, 256)]:
    x = torch.randn(B, cin, H, W, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    w = (torch.randn(co, cin, 1, 1, device=dev) * 0.1).contiguous(memory_format=torch.channels_last).requires_grad_()
    b = torch.zeros(co, device=dev, requires_grad=True)
    y = ops.conv2d(x, w, b); dy = torch.randn_like(y)
    y.backward(dy)
    torch.cuda.synchronize()
    A.PROFILE = []
    for _ in range(5):
        x.grad = None; w.grad = None
        y = ops.conv2d(x, w, b)
        y.backward(dy)
    torch.cuda.synchronize()
    agg = {}
    for nm, e0, e1 in A.PROFILE:
        agg.setdefault(nm, []).append(e0.elapsed_time(e1))
    A.PROFILE = None
    px = B * H * W
    print('%d->%d @%dx%d B=%d: in %.2f GB out %.2f GB' % (cin, co, H, W, B, px * cin * 4 / 1e9, px * co * 4 / 1e9))
    for nm, v in agg.items():
        v = sorted(v)
        print('   %-28s median %.3f ms' % (nm, v[len(v) // 2]))
# disparity heads (C -> 1, 3x3 reflect, sigmoid) in both forms
from improving_segmentation_with_selfsupervised_depth_b200 import conv_op
for few in (True, False):
    conv_op.HEAD_FEWCOUT = few
    for (B, cin, H, W) in [(12, 64, 512, 1024), (12, 128, 256, 512)]:
        x = torch.randn(B, cin, H, W, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
        w = (torch.randn(1, cin, 3, 3, device=dev) * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_()
        b = torch.zeros(1, device=dev, requires_grad=True)
        y = ops.conv2d(x, w, b, pad=1, pad_mode=A.PAD_REFLECT, act=A.ACT_SIGMOID); dy = torch.randn_like(y)
        y.backward(dy)
        torch.cuda.synchronize()
        A.PROFILE = []
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(5):
            x.grad = None; w.grad = None
            y = ops.conv2d(x, w, b, pad=1, pad_mode=A.PAD_REFLECT, act=A.ACT_SIGMOID)
            y.backward(dy)
        e1.record()
        torch.cuda.synchronize()
        agg = {}
        for nm, a0, a1 in A.PROFILE:
            agg.setdefault(nm, []).append(a0.elapsed_time(a1))
        A.PROFILE = None
        print('head %d->1 @%dx%d B=%d fewcout=%s: fwd+bwd %.3f ms (with event overhead)' % (cin, H, W, B, few, e0.elapsed_time(e1) / 5))
        for nm, v in agg.items():
            v = sorted(v)
            if v[len(v) // 2] > 0.01:
                print('   %-28s median %.3f ms x%d' % (nm, v[len(v) // 2], len(v) // 5))
