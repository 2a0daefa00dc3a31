import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import segsde_oracle as O
from helpers import *
golden = np.load('/root/repo/tests/golden/reference_golden.npz')
from improving_segmentation_with_selfsupervised_depth_b200.loss import MonodepthLoss
B, H, W, inputs, disps, Ts = loss_case(golden)
for variant in ["avg_reprojection"]:
    kw = dict(LOSS_KW); kw[variant] = True
    ml = MonodepthLoss(num_scales=4, frame_ids=[0,-1,1], height=H, width=W, batch_size=B, **kw)
    ml.replay_noise = loss_noise(B, H, W, True)
    gin = {k: v.cuda() for k, v in inputs.items()}
    gd = [d.cuda().requires_grad_() for d in disps]
    gT = {f: t.cuda().requires_grad_() for f, t in Ts.items()}
    outputs = {("disp", s): gd[s] for s in range(4)}
    outputs.update({("cam_T_cam", 0, f): gT[f] for f in (-1, 1)})
    ml.generate_images_pred(gin, outputs)
    losses = ml.compute_losses(gin, outputs)
    grads = torch.autograd.grad(losses["loss"], gd + [gT[-1], gT[1]])
    for i, g in enumerate(grads):
        ref = torch.from_numpy(golden["loss_%s_grad%d" % (variant, i)])
        d = (g.cpu() - ref).abs()
        sc = ref.abs().max()
        print(variant, i, 'max rel', (d.max()/sc).item(), 'frac bad', (d > 2e-4*sc).float().mean().item(), 'l2', ((g.cpu()-ref).norm()/ref.norm()).item())
# shapes case 192x640
H, W = 192, 640
B = 2
inputs = O.synthetic_inputs(B, H, W, seed=9)
g = torch.Generator().manual_seed(4)
disps = [torch.rand(B, 1, max(H >> s, 1), max(W >> s, 1), generator=g).mul(0.6).add(0.2) for s in range(4)]
Ts = {}
for f in (-1, 1):
    Ts[f] = O.transformation_from_parameters(torch.randn(B, 1, 3, generator=g) * 0.01, torch.randn(B, 1, 3, generator=g) * 0.05, invert=f < 0)
noise = [torch.randn(B, 2, H, W, generator=g) * 1e-5 for _ in range(4)]
cd = [d.clone().requires_grad_() for d in disps]
cT = {f: t.clone().requires_grad_() for f, t in Ts.items()}
ol = O.monodepth_loss(inputs, cd, cT, [0, -1, 1], H, W, noise=noise)
ol["loss"].backward()
ml = MonodepthLoss(num_scales=4, frame_ids=[0,-1,1], height=H, width=W, batch_size=B, **LOSS_KW)
ml.replay_noise = noise
gd = [d.cuda().requires_grad_() for d in disps]
gT = {f: t.cuda().requires_grad_() for f, t in Ts.items()}
outputs = {("disp", s): gd[s] for s in range(4)}
outputs.update({("cam_T_cam", 0, f): gT[f] for f in (-1, 1)})
gin = {k: v.cuda() for k, v in inputs.items()}
ml.generate_images_pred(gin, outputs)
gl = ml.compute_losses(gin, outputs)
for k in ol: print(k, ol[k].item(), gl[k].item())
gl["loss"].backward()
for s in range(4):
    a, b = gd[s].grad.cpu(), cd[s].grad
    d = (a-b).abs(); sc = b.abs().max()
    print('disp', s, 'max rel', (d.max()/sc).item(), 'frac bad', (d > 2e-4*sc).float().mean().item(), 'l2', ((a-b).norm()/b.norm()).item())
for f in (-1,1):
    a, b = gT[f].grad.cpu(), cT[f].grad
    print('T', f, ((a-b).abs().max()/b.abs().max()).item()); print(a[0]); print(b[0])
