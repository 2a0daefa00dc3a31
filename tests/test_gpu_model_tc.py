"""GPU: the BENCHMARKED route — the whole network through the tcgen05 kernels, ASPP pooling ON — at BASELINE sizes.

TF32 operand rounding (2^-11 relative) is amplified by a random-weight, train-mode-BatchNorm ResNet for ANY
implementation, so every check has two references:
  * the fp32 CPU oracle (the reference's arithmetic), and
  * the same oracle on this GPU with cuDNN's TF32 convolutions (torch's default = what the reference itself runs with
    on this hardware): its distance to the CPU oracle is the NOISE FLOOR, and the tcgen05 path must stay within a
    stated multiple of it.
Eval-mode BatchNorm (running statistics) does not amplify: there the bounds are absolute.
"""
import contextlib
import io

import numpy as np
import pytest
import torch

import segsde_oracle as O
from helpers import LOSS_KW, noise_floor_retry

pytestmark = pytest.mark.gpu


def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def make(backbone, H, W, seed=3, freeze=False):
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    from improving_segmentation_with_selfsupervised_depth_b200.synthetic import mono_config
    models, _ = P.install_dropin()
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.get_model(mono_config(backbone, H, W, freeze_backbone=freeze), 19)
    sd = O.synthetic_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    return model.cuda(), sd


def oracle_cfg(backbone):
    return {"num_layers": int(backbone[6:]), "rswd": [False, False, True], "frame_ids": [0, -1, 1]}


def tc_routes(ops):
    kinds = {}
    for kind, r in ops.ROUTES:
        kinds.setdefault(kind, []).append(r)
    return kinds


@pytest.mark.parametrize("backbone,H,W", [("resnet50", 192, 640), ("resnet101", 192, 640)])
@noise_floor_retry
def test_train_step_tc_route_config1(backbone, H, W):
    """BASELINE configs[0] geometry (192x640, B=2), train mode, ASPP pooling on, ResNet-50 and the ResNet-101 every
    shipped YAML uses: forward activations, the photometric loss and every parameter gradient of one step."""
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    from improving_segmentation_with_selfsupervised_depth_b200.loss import MonodepthLoss
    from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
    B = 2
    model, sd = make(backbone, H, W)
    model.train()
    g = torch.Generator().manual_seed(5)
    mask = (torch.rand(B, 256, H // 16, W // 16, generator=g) >= 0.5).float()
    for mod in model.modules():
        if isinstance(mod, Dropout):
            mod.replay_mask = mask
    inputs = O.synthetic_inputs(B, H, W, seed=9)
    noise = [torch.randn(B, 2, H, W, generator=g) * 1e-5 for _ in range(4)]
    cfg = oracle_cfg(backbone)

    def oracle_step(dev, tf32):
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = tf32
        if dev == "cuda":
            torch.cuda.empty_cache()      # cuDNN's algorithm choice depends on the workspace it can get: same start for every run
        try:
            osd = {k: v.to(dev).clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
            inp = {k: v.to(dev) for k, v in inputs.items()}
            ref = O.model_forward(osd, inp, cfg, O.BNMode(True), dropout_mask=mask.to(dev))
            loss = O.monodepth_loss(inp, [ref[("disp", s)] for s in range(4)],
                                    {f: ref[("cam_T_cam", 0, f)] for f in (-1, 1)}, [0, -1, 1], H, W,
                                    noise=[n.to(dev) for n in noise])["loss"]
            loss.backward()
            return ref, loss.detach(), {k: v.grad for k, v in osd.items() if v.grad is not None}
        finally:
            torch.backends.cudnn.allow_tf32 = prev
    ref, rloss, rgrad = oracle_step("cpu", False)             # the reference's arithmetic
    cref, closs, cgrad = oracle_step("cuda", True)            # the reference as it runs on this GPU
    ops.USE_TC, ops.ROUTES = True, []
    try:
        gin = {k: v.cuda() for k, v in inputs.items()}
        with contextlib.redirect_stdout(io.StringIO()):
            out = model(gin)
        ml = MonodepthLoss(height=H, width=W, batch_size=B, num_scales=4, frame_ids=[0, -1, 1], **LOSS_KW)
        ml.replay_noise = noise
        ml.generate_images_pred(gin, out)
        loss = ml.compute_losses(gin, out)["loss"]
        loss.backward()
        routes = tc_routes(ops)
    finally:
        ops.USE_TC, ops.ROUTES = False, None
    # the route under test really is the tensor-core one: every convolution except the 12-channel pose outputs and the
    # 1x1-pixel ASPP pooling branch (shapes outside the family by construction)
    assert routes["fprop"].count("generic") <= 6, [r for r in routes["fprop"] if r == "generic"]
    # weight gradients: 32-pixel GEMM-K boxes need Wo % 32 == 0 — at 192x640 that holds down to 1/4 resolution (160 px);
    # the 80- and 40-pixel-wide layers take the generic kernel here (at the bench's 512x1024 every width qualifies)
    assert sum(r.startswith("tc:") for r in routes["wgrad"]) >= 20, routes["wgrad"]
    assert "tc:rowhalo" in routes["fprop"] and "tc:wgrad3x3" in routes["wgrad"]
    feats = model.models["encoder"].features
    for i in range(5):
        floor = l2(cref["features"][i], ref["features"][i])
        assert l2(feats[i], ref["features"][i]) < 2.5 * floor + 1e-3, ("feature", i, floor)
    for s in range(4):
        floor = l2(cref[("disp", s)], ref[("disp", s)])
        assert l2(out[("disp", s)], ref[("disp", s)]) < 2.5 * floor + 1e-3, ("disp", s, floor)
    lfloor = abs(closs.item() - rloss.item()) / abs(rloss.item())
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) < 3 * lfloor + 2e-3, (loss.item(), rloss.item(), closs.item())
    bad, ratios = [], []
    for n, q in model.named_parameters():
        r = rgrad.get(n)
        if r is None or r.norm().item() == 0:
            continue
        e, floor = l2(q.grad, r), l2(cgrad[n], r)
        ratios.append(e / (floor + 1e-3))
        if floor > 0.3:          # the reference's own TF32 run is already noise on this parameter (tiny bias gradients)
            continue
        if e > 3.0 * floor + 0.05:
            bad.append((n, e, floor))
    assert not bad, bad[:8]
    assert float(np.median(ratios)) < 1.6, float(np.median(ratios))


@noise_floor_retry
def test_forward_loss_512x1024_tc_route():
    """The bench geometry itself (512x1024, batch 2 so that the CPU oracle finishes in seconds): forward + photometric
    loss on the tcgen05 route vs the fp32 CPU oracle, in eval mode (running statistics) and in train mode (what bench.py
    runs), each held to 2.5x (disparities) / 3x (losses) the distance the cuDNN-TF32 oracle itself has to the fp32 one."""
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    from improving_segmentation_with_selfsupervised_depth_b200.loss import MonodepthLoss
    B, H, W = 2, 512, 1024
    model, sd = make("resnet50", H, W, freeze=True)
    inputs = O.synthetic_inputs(B, H, W, seed=4)
    g = torch.Generator().manual_seed(6)
    noise = [torch.randn(B, 2, H, W, generator=g) * 1e-5 for _ in range(4)]
    cfg = oracle_cfg("resnet50")
    gin = {k: v.cuda() for k, v in inputs.items()}
    for training in (False, True):
        model.train(training)
        for mod in model.modules():              # deterministic: an all-ones mask on both sides (y = 2x in train mode)
            if isinstance(mod, torch.nn.Dropout):
                mod.replay_mask = torch.ones(B, 256, H // 16, W // 16)

        def oracle(dev, tf32):
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = tf32
            if dev == "cuda":
                torch.cuda.empty_cache()
            try:
                with torch.no_grad():
                    osd = {k: v.to(dev) for k, v in sd.items()}
                    inp = {k: v.to(dev) for k, v in inputs.items()}
                    ones = torch.ones(B, 256, H // 16, W // 16, device=dev)
                    ref = O.model_forward(osd, inp, cfg, O.BNMode(training), dropout_mask=ones if training else None)
                    loss = O.monodepth_loss(inp, [ref[("disp", s)] for s in range(4)],
                                            {f: ref[("cam_T_cam", 0, f)] for f in (-1, 1)}, [0, -1, 1], H, W,
                                            noise=[n.to(dev) for n in noise])
                return ref, loss
            finally:
                torch.backends.cudnn.allow_tf32 = prev
        ref, rl = oracle("cpu", False)
        cref, cl = oracle("cuda", True)
        ops.USE_TC, ops.ROUTES = True, []
        try:
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                out = model(gin)
                ml = MonodepthLoss(height=H, width=W, batch_size=B, num_scales=4, frame_ids=[0, -1, 1], **LOSS_KW)
                ml.replay_noise = noise
                ml.generate_images_pred(gin, out)
                losses = ml.compute_losses(gin, out)
            routes = tc_routes(ops)["fprop"]
        finally:
            ops.USE_TC, ops.ROUTES = False, None
        assert routes.count("generic") <= 6 and "tc:rowhalo" in routes and "tc:conv256" in routes, routes
        for s in range(4):
            e, floor = l2(out[("disp", s)], ref[("disp", s)]), l2(cref[("disp", s)], ref[("disp", s)])
            print("512x1024 %s disp%d: tcgen05 vs fp32 oracle %.2e, cuDNN-TF32 oracle vs fp32 oracle %.2e" % (
                "train" if training else "eval", s, e, floor))
            assert e < 2.5 * floor + 1e-3, ("train" if training else "eval", "disp", s, e, floor)
        for f in (-1, 1):
            assert l2(out[("cam_T_cam", 0, f)], ref[("cam_T_cam", 0, f)]) < 2e-3
        for key in ["loss"] + ["loss/%d" % s for s in range(4)]:
            e = abs(losses[key].item() - rl[key].item()) / abs(rl[key].item())
            floor = abs(cl[key].item() - rl[key].item()) / abs(rl[key].item())
            assert e < 3 * floor + (2e-3 if training else 1e-3), (training, key, e, floor)
