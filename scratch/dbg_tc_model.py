import sys, os, contextlib, io, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import torch
import segsde_oracle as O
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200 import ops, conv_op
from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
contracts = json.load(open('/root/repo/tests/golden/state_dict_contracts.json'))
H, W, B = 128, 256, 2
models, _ = P.install_dropin()
cfg = dict(contracts["mono_r50"]["cfg"]); cfg.update({"height": H, "width": W, "crop_h": H, "crop_w": W})
cfg["depth_args"] = dict(cfg["depth_args"], max_scale_size=[H, W], aspp_pooling=False)
def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
with contextlib.redirect_stdout(io.StringIO()):
    model = models.get_model(cfg, 19)
sd = O.synthetic_state_dict(model.state_dict(), seed=3)
g = torch.Generator().manual_seed(78)
mask = (torch.rand(B, 256, H // 16, W // 16, generator=g) >= 0.5).float()
inputs = O.synthetic_inputs(B, H, W, seed=5)
osd = {k: v.clone() for k, v in sd.items()}
ocfg = {"num_layers": 50, "rswd": [False, False, True], "frame_ids": [0, -1, 1], "depth_args": {"aspp_pooling": False}}
with torch.no_grad():
    ref = O.model_forward(osd, inputs, ocfg, O.BNMode(True), dropout_mask=mask)
gin = {k: v.cuda() for k, v in inputs.items()}
for tc, fuse in ((True, True),):
    ops.USE_TC = tc
    os.environ["SEGSDE_NO_BNFUSE"] = "0" if fuse else "1"
    model.load_state_dict(sd); m = model.cuda().train()
    for mod in m.modules():
        if isinstance(mod, Dropout): mod.replay_mask = mask
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = m(gin)
    feats = m.models["encoder"].features
    print("tc=%s fuse=%s feats %s upconv %s disp %s" % (tc, fuse, ["%.1e" % l2(feats[i], ref["features"][i]) for i in range(5)],
          ["%.1e" % l2(out[("upconv", i)], ref[("upconv", i)]) for i in range(5)], ["%.1e" % l2(out[("disp", s)], ref[("disp", s)]) for s in range(4)]))
