import sys, json, contextlib, io
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import segsde_oracle as O
from helpers import unpack_mask
from test_gpu_model import build
golden = np.load('/root/repo/tests/golden/reference_golden.npz')
contracts = json.load(open('/root/repo/tests/golden/state_dict_contracts.json'))
from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
name, (H, W) = "mono_r18", (64, 128)
B = 2
p = "model_%s_" % name
model, sd = build(contracts, name, H, W)
mask = unpack_mask(golden, p)
for mod in model.modules():
    if isinstance(mod, Dropout): mod.replay_mask = mask
inputs = O.synthetic_inputs(B, H, W, seed=5)
g = torch.Generator().manual_seed(77)
wd = [torch.randn(B, 1, H >> s, W >> s, generator=g) for s in range(4)]
wT = {f: torch.randn(B, 4, 4, generator=g) for f in (-1, 1)}
osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
cfg = {"num_layers": 18, "rswd": [False]*3, "frame_ids": [0, -1, 1]}
ref = O.model_forward(osd, inputs, cfg, O.BNMode(True), dropout_mask=mask)
for t in ref["features"]: t.retain_grad()
for i in range(5): ref[("upconv", i)].retain_grad()
rl = sum((ref[("disp", s)] * wd[s]).sum() for s in range(4)) + 100 * sum((ref[("cam_T_cam", 0, f)] * wT[f]).sum() for f in (-1, 1))
rl.backward()
gin = {k: v.cuda() for k, v in inputs.items()}
with contextlib.redirect_stdout(io.StringIO()):
    out = model(gin)
feats = model.models["encoder"].features
for t in feats: t.retain_grad()
for i in range(5): out[("upconv", i)].retain_grad()
gl = sum((out[("disp", s)] * wd[s].cuda()).sum() for s in range(4)) + 100 * sum((out[("cam_T_cam", 0, f)] * wT[f].cuda()).sum() for f in (-1, 1))
gl.backward()
def e(a, b): return ((a.cpu() - b).norm() / (b.norm() + 1e-12)).item()
for i in range(5):
    print('feat', i, 'fwd', e(feats[i], ref["features"][i]), 'grad', e(feats[i].grad, ref["features"][i].grad))
for i in range(4, -1, -1):
    print('upconv', i, 'fwd', e(out[("upconv", i)], ref[("upconv", i)]), 'grad', e(out[("upconv", i)].grad, ref[("upconv", i)].grad))
for n, q in model.named_parameters():
    if osd[n].grad is None: continue
    er = e(q.grad, osd[n].grad)
    if er > 5e-4: print(n, er)
