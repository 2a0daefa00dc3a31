import sys, os, contextlib, io, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import torch
import segsde_oracle as O
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200 import ops
from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
contracts = json.load(open('/root/repo/tests/golden/state_dict_contracts.json'))
models, _ = P.install_dropin()
def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
for name, nl, rswd, (H, W) in (("mono_r18", 18, [False]*3, (64, 256)), ("mono_r50", 50, [False, False, True], (64, 128))):
    B = 2
    cfg = dict(contracts[name]["cfg"]); cfg.update({"height": H, "width": W, "crop_h": H, "crop_w": W})
    cfg["depth_args"] = dict(cfg["depth_args"], max_scale_size=[H, W], aspp_pooling=False)
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.get_model(cfg, 19)
    sd = O.synthetic_state_dict(model.state_dict(), seed=3)
    g = torch.Generator().manual_seed(78)
    fh = H // (32 if nl == 18 else 16); fw = W // (32 if nl == 18 else 16)
    mask = (torch.rand(B, 256, fh, fw, generator=g) >= 0.5).float()
    inputs = O.synthetic_inputs(B, H, W, seed=5)
    osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    ocfg = {"num_layers": nl, "rswd": rswd, "frame_ids": [0, -1, 1], "depth_args": {"aspp_pooling": False}}
    ref = O.model_forward(osd, inputs, ocfg, O.BNMode(True), dropout_mask=mask)
    wu = [torch.randn(ref[("upconv", i)].shape, generator=g) for i in range(5)]
    rl = sum((ref[("upconv", i)] * wu[i]).mean() for i in range(5)) + 100 * ref[("cam_T_cam", 0, 1)].sum()
    rl.backward()
    gin = {k: v.cuda() for k, v in inputs.items()}
    # oracle on the GPU with cuDNN TF32 convolutions (what the reference does on this GPU by default)
    torch.backends.cudnn.allow_tf32 = True
    csd = {k: v.detach().cuda().clone().requires_grad_(v.requires_grad) for k, v in osd.items()}
    cref = O.model_forward(csd, gin, ocfg, O.BNMode(True), dropout_mask=mask.cuda())
    cl = sum((cref[("upconv", i)] * wu[i].cuda()).mean() for i in range(5)) + 100 * cref[("cam_T_cam", 0, 1)].sum()
    cl.backward()
    grp = {}
    for n, q in csd.items():
        r = osd[n].grad
        if r is None or q.grad is None or r.norm().item() == 0: continue
        key = n.split('.')[1] + ('.' + n.split('.')[3] if n.split('.')[1] == 'encoder' and 'layer' in n else '')
        grp[key] = max(grp.get(key, 0.0), l2(q.grad, r))
    print(name, "cudnn-tf32 oracle", "feats", ["%.1e" % l2(cref["features"][i], ref["features"][i]) for i in range(5)], "upconv", ["%.1e" % l2(cref[("upconv", i)], ref[("upconv", i)]) for i in range(5)])
    print("   grad err by group:", {k: "%.1e" % v for k, v in grp.items()})
    for tc in (False, True):
        ops.USE_TC = tc
        model.load_state_dict(sd); m = model.cuda().train()
        for p_ in m.parameters(): p_.grad = None
        for mod in m.modules():
            if isinstance(mod, Dropout): mod.replay_mask = mask
        with contextlib.redirect_stdout(io.StringIO()):
            out = m(gin)
        gl = sum((out[("upconv", i)] * wu[i].cuda()).mean() for i in range(5)) + 100 * out[("cam_T_cam", 0, 1)].sum()
        gl.backward()
        feats = m.models["encoder"].features
        grp = {}
        for n, q in m.named_parameters():
            r = osd[n].grad
            if r is None or r.norm().item() == 0: continue
            key = n.split('.')[1] + ('.' + n.split('.')[3] if n.split('.')[1] == 'encoder' and 'layer' in n else '')
            grp[key] = max(grp.get(key, 0.0), l2(q.grad, r))
        print(name, "tc=%s" % tc, "feats", ["%.1e" % l2(feats[i], ref["features"][i]) for i in range(5)], "upconv", ["%.1e" % l2(out[("upconv", i)], ref[("upconv", i)]) for i in range(5)])
        print("   grad err by group:", {k: "%.1e" % v for k, v in grp.items()})
