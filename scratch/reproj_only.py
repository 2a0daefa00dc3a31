"""Times (CUDA events) the fused reprojection launch alone at the bench workload (B=12, 512x1024, 4 scales, gradients on).
   python scratch/reproj_only.py [--iters 20] [--B 12] [--H 512] [--W 1024] [--fwd]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import improving_segmentation_with_selfsupervised_depth_b200 as P      # noqa: E402
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A      # noqa: E402
from improving_segmentation_with_selfsupervised_depth_b200.synthetic import MONO_LOSS_KW, synthetic_inputs      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--B", type=int, default=12)
ap.add_argument("--H", type=int, default=512)
ap.add_argument("--W", type=int, default=1024)
ap.add_argument("--fwd", action="store_true")
a = ap.parse_args()
_, loss = P.install_dropin()
B, H, W = a.B, a.H, a.W
inp = {k: v.cuda() for k, v in synthetic_inputs(B, H, W, seed=1).items()}
g = torch.Generator().manual_seed(0)
disps = [(0.3 + 0.4 * torch.rand(B, 1, H >> s, W >> s, generator=g)).cuda().requires_grad_(not a.fwd) for s in range(4)]
T = {}
for f in (-1, 1):
    M = torch.eye(4).repeat(B, 1, 1)
    M[:, 0, 3] = 0.05 * f
    M[:, 2, 3] = 0.02 * f
    T[f] = M.cuda().requires_grad_(not a.fwd)
out = {("disp", s): disps[s] for s in range(4)}
out.update({("cam_T_cam", 0, f): T[f] for f in (-1, 1)})
ml = loss.MonodepthLoss(height=H, width=W, batch_size=B, **MONO_LOSS_KW)
A.PROFILE, A.PROFILE_NAMES = [], {"segsde_reproj_fused"}
for i in range(a.iters + 3):
    if i == 3:
        A.PROFILE.clear()
    l = ml.compute_losses(inp, out)["loss"]
torch.cuda.synchronize()
ms = sorted(e0.elapsed_time(e1) for _, e0, e1 in A.PROFILE)
alg = B * (4 * 9 * H * W + (1 if a.fwd else 2) * sum(4 * (H >> s) * (W >> s) for s in range(4)))
print("reproj fused launch: median %.3f ms  min %.3f ms  (%d iters) loss %.6f  algorithmic %.1f MB -> %.0f GB/s"
      % (ms[len(ms) // 2], ms[0], len(ms), float(l.detach()), alg / 1e6, alg / (ms[len(ms) // 2] * 1e-3) / 1e9))
