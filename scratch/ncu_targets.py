"""Launches each hot kernel a couple of times at the bench shapes (B=12, 512x1024) for an ncu --set full capture."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
from improving_segmentation_with_selfsupervised_depth_b200.synthetic import MONO_LOSS_KW, synthetic_inputs
models, loss = P.install_dropin()
B, H, W = 12, 512, 1024
dev = torch.device('cuda')
def cl(*s): return torch.randn(*s, device=dev).contiguous(memory_format=torch.channels_last)
# 1. disparity head 64->1 @512x1024 (c1 kernels) with sigmoid
x = cl(B, 64, H, W).requires_grad_()
w = (torch.randn(1, 64, 3, 3, device=dev) * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_()
b = torch.zeros(1, device=dev, requires_grad=True)
for _ in range(2):
    y = ops.conv2d(x, w, b, pad=1, pad_mode=A.PAD_REFLECT, act=A.ACT_SIGMOID)
    y.backward(torch.ones_like(y))
del x, y
# 2. decoder convs on tensor cores
def conv(c1, c2, cout, h, w_, up):
    x1 = cl(B, c1, h, w_).requires_grad_()
    x2 = cl(B, c2, h * (2 if up else 1), w_ * (2 if up else 1)).requires_grad_() if c2 else None
    wt = (torch.randn(cout, c1 + c2, 3, 3, device=dev) * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_()
    bb = torch.zeros(cout, device=dev, requires_grad=True)
    for _ in range(2):
        y = ops.conv2d(x1, wt, bb, x2=x2, pad=1, pad_mode=A.PAD_REFLECT, up1=up, act=A.ACT_ELU)
        y.backward(torch.ones_like(y))
conv(64, 0, 64, 256, 512, True)
conv(128, 64, 128, 128, 256, True)
# 2b. generic kernel with 256-pixel tiles: encoder 1x1 convs (ResNet-50 layer3 / layer4 shapes) and a narrow 3x3
def conv1(cin, cout, h, w_, k=1, dil=1):
    x1 = cl(B, cin, h, w_).requires_grad_()
    wt = (torch.randn(cout, cin, k, k, device=dev) * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_()
    for _ in range(2):
        y = ops.conv2d(x1, wt, None, pad=dil * (k // 2), dil=dil)
        y.backward(torch.ones_like(y))
conv1(1024, 2048, 32, 64)
conv1(256, 1024, 32, 64)
conv1(512, 512, 32, 64, k=3, dil=2)
conv1(2048, 256, 32, 64, k=3, dil=12)
# 3. BN on an encoder-sized tensor, stem
x = cl(B, 256, 128, 256).requires_grad_()
g, bt = torch.ones(256, device=dev, requires_grad=True), torch.zeros(256, device=dev, requires_grad=True)
rm, rv = torch.zeros(256, device=dev), torch.ones(256, device=dev)
for _ in range(2):
    y = ops.batch_norm(x, g, bt, rm, rv, True, 0.1, 1e-5, act=A.ACT_RELU)
    y.backward(torch.ones_like(y))
img = torch.rand(B, 3, H, W, device=dev)
w7 = (torch.randn(64, 3, 7, 7, device=dev) * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_()
for _ in range(2):
    y = ops.conv2d(img, w7, stride=2, pad=3, nchw_norm_in=True)
    y.backward(torch.ones_like(y))
# 4. loss
inputs = {k: v.to(dev) for k, v in synthetic_inputs(B, H, W).items()}
ml = loss.MonodepthLoss(height=H, width=W, batch_size=B, **MONO_LOSS_KW)
disps = [torch.rand(B, 1, H >> s, W >> s, device=dev) * 0.6 + 0.2 for s in range(4)]
T = {f: torch.eye(4, device=dev).repeat(B, 1, 1) for f in (-1, 1)}
T[-1][:, 0, 3] = 0.05; T[1][:, 0, 3] = -0.05
out = {("disp", s): disps[s].clone().requires_grad_() for s in range(4)}
out.update({("cam_T_cam", 0, f): T[f] for f in (-1, 1)})
ml.generate_images_pred(inputs, out)
ml.compute_losses(inputs, out)["loss"].backward()
torch.cuda.synchronize()
print("done")
# 5. round-2 additions: the 19-class 1x1 segmentation head (few-output-channel kernels), max-pool
xs = cl(8, 64, H, W).requires_grad_()
wsg = (torch.randn(19, 64, 1, 1, device=dev) * 0.1).contiguous(memory_format=torch.channels_last).requires_grad_()
for _ in range(2):
    y = ops.conv2d(xs, wsg, torch.zeros(19, device=dev, requires_grad=True))
    y.backward(torch.ones_like(y))
xm = cl(B, 64, 256, 512).requires_grad_()
for _ in range(2):
    y = ops.maxpool3x3s2(xm)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done 5")
