"""GPU: the whole drop-in network (encoder + ASPP depth decoder + pose net) and one training-step gradient
against the committed reference outputs and the CPU oracle.  fp32 CUDA-core path: activations 2e-4, loss 2e-5,
gradients 5e-3 relative (per-tensor norm) — batch-norm over B*H*W ~ 100 elements at the bottleneck amplifies
summation-order noise in the gradients."""
import contextlib
import io

import numpy as np
import pytest
import torch

import segsde_oracle as O
from helpers import LOSS_KW, noise_floor_retry, rel_err, unpack_mask

pytestmark = pytest.mark.gpu


def build(contracts, name, H, W, use_tc=False):
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    ops.USE_TC = use_tc
    models, _ = P.install_dropin()
    c = contracts[name]
    cfg = dict(c["cfg"])
    cfg.update({"height": H, "width": W, "crop_h": H, "crop_w": W})
    cfg["depth_args"] = dict(cfg["depth_args"], max_scale_size=[H, W])
    with contextlib.redirect_stdout(io.StringIO()):
        m = models.get_model(cfg, 19)
    sd = O.synthetic_state_dict({k: torch.empty(s) for k, s in c["state_dict"].items()}, seed=1)
    m.load_state_dict(sd)
    return m.cuda().train(), sd


@pytest.mark.parametrize("name,hw", [("mono_r18", (64, 128)), ("mono_r50", (64, 96))])
def test_model_step_vs_reference_golden(golden, contracts, name, hw):
    from improving_segmentation_with_selfsupervised_depth_b200.loss import MonodepthLoss
    from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
    H, W = hw
    B = 2
    p = "model_%s_" % name
    model, _ = build(contracts, name, H, W)
    for mod in model.modules():
        if isinstance(mod, Dropout):
            mod.replay_mask = unpack_mask(golden, p)
    inputs = {k: v.cuda() for k, v in O.synthetic_inputs(B, H, W, seed=5).items()}
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    for s in range(4):
        assert out[("disp", s)].shape == golden[p + "disp%d" % s].shape
        assert rel_err(out[("disp", s)], golden[p + "disp%d" % s]) < 2e-4, s
    for f in (-1, 1):
        assert rel_err(out[("cam_T_cam", 0, f)], golden[p + "T%d" % f]) < 1e-4
    ml = MonodepthLoss(num_scales=4, frame_ids=[0, -1, 1], height=H, width=W, batch_size=B, **LOSS_KW)
    torch.manual_seed(31)
    ml.replay_noise = [torch.randn(B, 2, H, W) * 0.00001 for _ in range(4)]
    ml.generate_images_pred(inputs, out)
    losses = ml.compute_losses(inputs, out)
    vals = np.array([losses["loss/%d" % s].item() for s in range(4)] + [losses["loss"].item()])
    np.testing.assert_allclose(vals, golden[p + "losses"], rtol=5e-5)
    # end-to-end gradient of the photometric loss: loose bound (see helpers.grad_close for why the loss
    # gradient is only piecewise continuous); the tight per-tensor check is test_network_gradients_vs_oracle
    losses["loss"].backward()
    params = dict(model.named_parameters())
    names = [str(n) for n in golden[p + "grad_names"]]
    norms = np.array([params[n].grad.norm().item() for n in names])
    ref = golden[p + "grad_norms"]
    bad = [(n, a, b) for n, a, b in zip(names, norms, ref) if abs(a - b) > 3e-2 * b + 1e-9]
    assert not bad, bad[:5]
    assert rel_err(params["models.pose.net.3.weight"].grad, golden[p + "grad_pose3"]) < 3e-2
    sd = model.state_dict()
    assert rel_err(sd["models.encoder.encoder.bn1.running_mean"], golden[p + "bn1_running_mean"]) < 1e-5
    assert rel_err(sd["models.encoder.encoder.bn1.running_var"], golden[p + "bn1_running_var"]) < 1e-5
    assert int(sd["models.encoder.encoder.bn1.num_batches_tracked"]) == 1


@pytest.mark.parametrize("name,hw", [("mono_r18", (64, 128)), ("mono_r50", (64, 96))])
def test_network_gradients_vs_oracle(golden, contracts, name, hw):
    """Backward of every network kernel in context: a smooth surrogate loss (fixed random weights on the
    disparities and pose matrices) so that the only discontinuities are ReLU / max-pool ties; every parameter
    gradient must match the CPU oracle's autograd to 2e-3 (relative L2 per tensor)."""
    from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
    H, W = hw
    B = 2
    p = "model_%s_" % name
    model, sd = build(contracts, name, H, W)
    mask = unpack_mask(golden, p)
    for mod in model.modules():
        if isinstance(mod, Dropout):
            mod.replay_mask = mask
    inputs = O.synthetic_inputs(B, H, W, seed=5)
    g = torch.Generator().manual_seed(77)
    wd = [torch.randn(B, 1, H >> s, W >> s, generator=g) for s in range(4)]
    wT = {f: torch.randn(B, 4, 4, generator=g) for f in (-1, 1)}
    c = contracts[name]
    cfg = {"num_layers": int(c["cfg"]["backbone_name"][6:]), "rswd": c["cfg"]["replace_stride_with_dilation"],
           "frame_ids": [0, -1, 1]}
    # The oracle is evaluated in fp32 (the reference's arithmetic) AND in fp64: ASPPPooling's BatchNorm sees
    # B x 256 x 1 x 1 = two values per channel, which makes the encoder gradient ill-conditioned (the fp32 CPU
    # result itself is ~1e-2 away from fp64 for ResNet-50 here), so the bound is 2e-3 + 2x the oracle's own
    # fp32 error per tensor, measured against the fp64 result.
    grads = {}
    for dt in (torch.float32, torch.float64):
        osd = {k: (v.to(dt) if v.dtype.is_floating_point else v).clone().requires_grad_(
            v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
        inp = {k: v.to(dt) for k, v in inputs.items()}
        old = torch.get_default_dtype()
        torch.set_default_dtype(dt)
        try:
            ref = O.model_forward(osd, inp, cfg, O.BNMode(True), dropout_mask=mask.to(dt))
            rl = sum((ref[("disp", s)] * wd[s].to(dt)).sum() for s in range(4)) + \
                100 * sum((ref[("cam_T_cam", 0, f)] * wT[f].to(dt)).sum() for f in (-1, 1))
            rl.backward()
        finally:
            torch.set_default_dtype(old)
        grads[dt] = {k: v.grad.double() for k, v in osd.items() if v.grad is not None}
    gin = {k: v.cuda() for k, v in inputs.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(gin)
    # weighted sums via the library's own scale/accumulate entry point would hide nothing: use torch for the
    # test-side reduction only
    gl = sum((out[("disp", s)] * wd[s].cuda()).sum() for s in range(4)) + \
        100 * sum((out[("cam_T_cam", 0, f)] * wT[f].cuda()).sum() for f in (-1, 1))
    assert abs(gl.item() - rl.item()) < 1e-3 * abs(rl.item()) + 1e-3
    gl.backward()
    bad = []
    for n, q in model.named_parameters():
        if n not in grads[torch.float64]:
            assert q.grad is None or q.grad.abs().max().item() == 0, n
            continue
        r64, r32 = grads[torch.float64][n], grads[torch.float32][n]
        own = ((r32 - r64).norm() / (r64.norm() + 1e-12)).item()
        e = ((q.grad.double().cpu() - r64).norm() / (r64.norm() + 1e-12)).item()
        if e > 2e-3 + 4 * own:
            bad.append((n, e, own))
    assert not bad, bad[:8]


def test_frozen_encoder_and_eval_mode(contracts):
    """Config-2 style freezing (freeze_backbone: only decoder+pose get gradients) and eval-mode BN."""
    H, W, B = 64, 96, 2
    model, sd = build(contracts, "mono_r50", H, W)
    for q in model.models["encoder"].parameters():
        q.requires_grad = False
    inputs = {k: v.cuda() for k, v in O.synthetic_inputs(B, H, W, seed=6).items()}
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    (out[("disp", 0)].mean() + out[("cam_T_cam", 0, 1)].sum()).backward()
    assert all(q.grad is None for q in model.models["encoder"].parameters())
    assert model.models["depth"].convs[("upconv", 0, 1)].block[0].conv.weight.grad is not None
    assert model.models["pose"].net[3].weight.grad is not None
    model.eval()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        ev = model(inputs)
    model2, _ = build(contracts, "mono_r50", H, W)
    osd = {k: v.clone() for k, v in model.state_dict().items()}
    cpu_sd = {k: v.cpu() for k, v in osd.items()}
    cfg = {"num_layers": 50, "rswd": [False, False, True], "frame_ids": [0, -1, 1]}
    cin = {k: v.cpu() for k, v in inputs.items()}
    with torch.no_grad():
        ref = O.model_forward(cpu_sd, cin, cfg, O.BNMode(False))
    for s in range(4):
        assert rel_err(ev[("disp", s)], ref[("disp", s)]) < 2e-4


@noise_floor_retry
def test_model_tensor_core_path_vs_oracle(contracts):
    """The network through the tcgen05 route (TF32 operands, fp32 accumulate) in TRAIN mode (batch statistics from the
    convolution epilogues).  To keep the comparison with the fp32 oracle well conditioned the ASPP pooling branch is
    switched off (its BatchNorm over B x 256 x 1 x 1 amplifies rounding by 1/sqrt(var+eps), see
    test_network_gradients_vs_oracle) — with batch-normalised activations the ELUs do not saturate, so gradients are
    meaningful.  Activations are held to depth-aware absolute bounds; parameter gradients to the error the
    oracle itself shows on this GPU with cuDNN TF32 convolutions (the reference's default numerics)."""
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
    H, W, B = 64, 128, 2
    models, _ = P.install_dropin()
    cfg = dict(contracts["mono_r50"]["cfg"])
    cfg.update({"height": H, "width": W, "crop_h": H, "crop_w": W})
    cfg["depth_args"] = dict(cfg["depth_args"], max_scale_size=[H, W], aspp_pooling=False)
    ops.USE_TC = True

    def l2(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return ((a - b).norm() / (b.norm() + 1e-30)).item()
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            model = models.get_model(cfg, 19)
        sd = O.synthetic_state_dict(model.state_dict(), seed=3)
        model.load_state_dict(sd)
        model = model.cuda().train()
        g = torch.Generator().manual_seed(78)
        mask = (torch.rand(B, 256, H // 16, W // 16, generator=g) >= 0.5).float()
        for mod in model.modules():
            if isinstance(mod, Dropout):
                mod.replay_mask = mask
        inputs = O.synthetic_inputs(B, H, W, seed=5)
        osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
        ocfg = {"num_layers": 50, "rswd": [False, False, True], "frame_ids": [0, -1, 1],
                "depth_args": {"aspp_pooling": False}}
        ref = O.model_forward(osd, inputs, ocfg, O.BNMode(True), dropout_mask=mask)
        wu = [torch.randn(ref[("upconv", i)].shape, generator=g) for i in range(5)]
        wd = [torch.randn(ref[("disp", s)].shape, generator=g) for s in range(4)]
        rl = sum((ref[("upconv", i)] * wu[i]).mean() for i in range(5)) + sum((ref[("disp", s)] * wd[s]).mean() for s in range(4)) \
            + 100 * ref[("cam_T_cam", 0, 1)].sum()
        rl.backward()
        gin = {k: v.cuda() for k, v in inputs.items()}
        ops.PROFILE = []
        with contextlib.redirect_stdout(io.StringIO()):
            out = model(gin)
        assert "fprop" in {k for k, *_ in ops.PROFILE}
        feats = model.models["encoder"].features
        for i in range(5):
            # a random-weight ResNet-50 amplifies any perturbation ~4x per stage (the fp32 CUDA path itself goes
            # 6e-8 -> 5e-5 from stem to bottleneck); TF32 enters at 4e-4, so the bound grows with depth
            assert l2(feats[i], ref["features"][i]) < (2e-3, 6e-3, 2e-2, 8e-2, 2e-1)[i], ("feature", i)
            assert l2(out[("upconv", i)], ref[("upconv", i)]) < 8e-2, ("upconv", i)
        for s in range(4):
            assert l2(out[("disp", s)], ref[("disp", s)]) < 4e-2, ("disp", s)
        assert rel_err(out[("cam_T_cam", 0, 1)], ref[("cam_T_cam", 0, 1)]) < 1e-3
        gl = sum((out[("upconv", i)] * wu[i].cuda()).mean() for i in range(5)) + \
            sum((out[("disp", s)] * wd[s].cuda()).mean() for s in range(4)) + 100 * out[("cam_T_cam", 0, 1)].sum()
        gl.backward()
        # Noise floor: the same oracle on this GPU with cuDNN TF32 convolutions — the numerics the reference itself
        # runs with by default (torch.backends.cudnn.allow_tf32 is True).  A random-weight, train-mode-BN ResNet-50
        # amplifies TF32 rounding to O(0.5) relative error in the encoder gradients for BOTH implementations, so the
        # tcgen05 path is held to that floor rather than to an absolute bound (per-kernel TF32 parity with tight
        # bounds is tests/test_gpu_tc.py; the fp32 route is checked tightly in test_network_gradients_vs_oracle).
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        torch.cuda.empty_cache()      # cuDNN's algorithm choice depends on the workspace it can get
        try:
            csd = {k: v.detach().cuda().clone().requires_grad_(v.requires_grad) for k, v in osd.items()}
            cref = O.model_forward(csd, gin, ocfg, O.BNMode(True), dropout_mask=mask.cuda())
            cl = sum((cref[("upconv", i)] * wu[i].cuda()).mean() for i in range(5)) + \
                sum((cref[("disp", s)] * wd[s].cuda()).mean() for s in range(4)) + 100 * cref[("cam_T_cam", 0, 1)].sum()
            cl.backward()
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        for i in range(5):
            assert l2(feats[i], ref["features"][i]) < 2.5 * l2(cref["features"][i], ref["features"][i]) + 1e-3, ("feature/floor", i)
            assert l2(out[("upconv", i)], ref[("upconv", i)]) < 2.5 * l2(cref[("upconv", i)], ref[("upconv", i)]) + 1e-3
        bad, ratios = [], []
        for n, q in model.named_parameters():
            r = osd[n].grad
            if r is None or r.norm().item() == 0:
                continue
            e, floor = l2(q.grad, r), l2(csd[n].grad, r)
            ratios.append(e / (floor + 1e-3))
            if e > 3.0 * floor + 0.05:
                bad.append((n, e, floor))
        assert not bad, bad[:8]
        assert float(np.median(ratios)) < 1.6, float(np.median(ratios))
        rm = model.state_dict()["models.encoder.encoder.layer3.0.bn2.running_var"]
        assert l2(rm, osd["models.encoder.encoder.layer3.0.bn2.running_var"]) < 1e-2
    finally:
        ops.PROFILE = None
        ops.USE_TC = False


@pytest.mark.parametrize("use_tc", [False, True])
@pytest.mark.parametrize("name", ["segdec_r50", "pad_r50"])
def test_seg_decoders_vs_reference_golden(golden, contracts, name, use_tc):
    """JointSegDepthDecoder / PAD (+ SelfAttention gate, bilinear resize, seg heads) and cross_entropy2d through the
    drop-in API against the reference outputs.  fp32 CUDA-core route: logits 2e-4, loss 2e-5, per-parameter gradient
    norms 3e-2.  tcgen05 route (what configs 4 / 5 run): TF32 operands through a train-mode-BatchNorm ResNet-50 whose
    bottleneck is 4 x 6 pixels here (BatchNorm over 48 values amplifies operand rounding ~100x) — logits / disparities
    0.15, loss 3e-2 (gradients not compared at this geometry): a smoke bound that proves the tensor-core kernels ran and produced the right
    network; the quantitative TF32 bounds (against the cuDNN-TF32 noise floor of the reference itself) are
    tests/test_gpu_model_tc.py and the PAD step of tests/test_gpu_dropin_trainer.py."""
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    from improving_segmentation_with_selfsupervised_depth_b200.loss.loss import cross_entropy2d
    from helpers import unpack_named_mask
    ops.USE_TC = use_tc
    t_act, t_loss, t_grad = (0.15, 3e-2, 0.5) if use_tc else (2e-4, 2e-5, 3e-2)
    ops.ROUTES = [] if use_tc else None
    models, _ = P.install_dropin()
    H, W, B = 64, 96, 2
    c = contracts[name]
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.get_model(dict(c["cfg"]), 19)
    sd = O.synthetic_state_dict({k: torch.empty(s) for k, s in c["state_dict"].items()}, seed=2)
    model.load_state_dict(sd)
    model = model.cuda().train()
    p = "seg_%s_" % name
    if name == "segdec_r50":
        dec = model.models["segmentation"]
        dec.unet_dec.convs[("upconv", 4, 0)].project[3].replay_mask = unpack_named_mask(golden, p + "mask_aspp")
        dec.head[4].replay_mask = unpack_named_mask(golden, p + "mask_head")
    else:
        dec = model.models["mtl_decoder"]
        dec.depth_dec.convs[("upconv", 4, 0)].project[3].replay_mask = unpack_named_mask(golden, p + "mask_depth")
        dec.seg_dec.convs[("upconv", 4, 0)].project[3].replay_mask = unpack_named_mask(golden, p + "mask_seg")
    inputs = {k: v.cuda() for k, v in O.synthetic_inputs(B, H, W, seed=6, labels=True).items()}
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    print(name, "tc" if use_tc else "fp32", "semantics err", rel_err(out["semantics"][:, :, ::4, ::4], golden[p + "semantics"]))
    assert rel_err(out["semantics"][:, :, ::4, ::4], golden[p + "semantics"]) < t_act
    loss = cross_entropy2d(input=out["semantics"], target=inputs["lbl"])
    if "intermediate_semantics" in out:
        assert rel_err(out["intermediate_semantics"], golden[p + "intermediate"]) < t_act
        for s in range(4):
            assert rel_err(out[("disp", s)], golden[p + "disp%d" % s]) < t_act
        loss = ops.add(loss.reshape(1, 1, 1, 1), cross_entropy2d(input=out["intermediate_semantics"],
                                                                 target=inputs["lbl"]).reshape(1, 1, 1, 1)).reshape(()) / 2
    print(name, "loss err", abs(loss.item() - float(golden[p + "loss"])) / abs(float(golden[p + "loss"])))
    assert abs(loss.item() - float(golden[p + "loss"])) < t_loss * abs(float(golden[p + "loss"]))
    loss.backward()
    routes, ops.ROUTES = ops.ROUTES, None
    ops.USE_TC = False
    if use_tc:
        tc = [r for _, r in routes if r.startswith("tc:")]
        # at 64x96 the deep layers are 4-12 pixels wide (outside the family: boxes of >= 8 pixels, wgrad 32-pixel boxes)
        assert len(tc) > 0.3 * len(routes) and any(r.endswith("wgrad3x3") for r in tc), (len(tc), len(routes))
    if use_tc:      # TF32 through BatchNorm over 48 values: per-parameter gradient norms carry O(1) noise for ANY TF32
        return      # implementation at this geometry; gradients of the tcgen05 route are bounded in test_gpu_model_tc.py
    params = dict(model.named_parameters())
    names = [str(n) for n in golden[p + "grad_names"]]
    bad, worst = [], 0.0
    for n, ref in zip(names, golden[p + "grad_norms"]):
        g = params[n].grad
        v = 0.0 if g is None else g.norm().item()
        worst = max(worst, abs(v - ref) / (ref + 1e-7))
        if abs(v - ref) > t_grad * ref + 1e-7:
            bad.append((n, v, float(ref)))
    print(name, "worst grad-norm err", worst)
    assert not bad, bad[:6]
