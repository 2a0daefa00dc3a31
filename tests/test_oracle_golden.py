"""CPU: the oracle restatement reproduces the committed reference outputs (tests/golden)."""
import numpy as np
import pytest
import torch

import segsde_oracle as O
from helpers import LOSS_KW, loss_case, loss_noise, rel_err, unpack_mask


@pytest.mark.parametrize("variant", ["default", "no_ssim", "avg_reprojection", "disable_automasking"])
def test_loss_matches_reference_golden(golden, variant):
    B, H, W, inputs, disps, Ts = loss_case(golden)
    for d in disps:
        d.requires_grad_()
    for t in Ts.values():
        t.requires_grad_()
    kw = dict(LOSS_KW)
    if variant != "default":
        kw[variant] = True
    noise = loss_noise(B, H, W, kw["avg_reprojection"])
    extras = {}
    out = O.monodepth_loss(inputs, disps, Ts, [0, -1, 1], H, W, kw["min_depth"], kw["max_depth"],
                           kw["disparity_smoothness"], kw["no_ssim"], kw["avg_reprojection"],
                           kw["disable_automasking"], noise, extras)
    vals = np.array([out["loss/%d" % s].item() for s in range(4)] + [out["loss"].item()])
    np.testing.assert_allclose(vals, golden["loss_%s_values" % variant], rtol=1e-6)
    grads = torch.autograd.grad(out["loss"], disps + [Ts[-1], Ts[1]])
    for i, g in enumerate(grads):
        assert rel_err(g, golden["loss_%s_grad%d" % (variant, i)]) < 1e-5
    if variant == "default":
        assert rel_err(extras[("color", -1, 0)], golden["loss_default_color_m1_s0"]) < 1e-6
        assert rel_err(extras[("sample", 1, 2)], golden["loss_default_sample_p1_s2"]) < 1e-6
        assert (extras["identity_selection/0"].numpy() == golden["loss_default_identsel_s0"]).all()


def test_cross_entropy_matches_reference_golden(golden):
    lg = torch.from_numpy(golden["ce_logits"]).requires_grad_()
    tgt, pw = torch.from_numpy(golden["ce_target"]), torch.from_numpy(golden["ce_pw"])
    l0 = O.cross_entropy2d(lg, tgt)
    assert abs(l0.item() - float(golden["ce_loss"])) < 1e-6
    assert rel_err(torch.autograd.grad(l0, lg)[0], golden["ce_grad"]) < 1e-6
    l1 = O.cross_entropy2d(lg, tgt, pixel_weights=pw)
    assert abs(l1.item() - float(golden["ce_pw_loss"])) < 1e-6
    assert rel_err(torch.autograd.grad(l1, lg)[0], golden["ce_pw_grad"]) < 1e-6
    assert abs(O.cross_entropy2d(torch.from_numpy(golden["ce_small_logits"]), tgt).item()
               - float(golden["ce_small_loss"])) < 1e-6


@pytest.mark.parametrize("name,hw", [("mono_r18", (64, 128)), ("mono_r50", (64, 96))])
def test_model_matches_reference_golden(golden, contracts, name, hw):
    H, W = hw
    B = 2
    c = contracts[name]
    template = {k: torch.empty(s) for k, s in c["state_dict"].items()}
    sd = O.synthetic_state_dict(template, seed=1)
    osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    inputs = O.synthetic_inputs(B, H, W, seed=5)
    p = "model_%s_" % name
    cfg = {"num_layers": int(c["cfg"]["backbone_name"][6:]), "rswd": c["cfg"]["replace_stride_with_dilation"],
           "frame_ids": [0, -1, 1]}
    out = O.model_forward(osd, inputs, cfg, O.BNMode(True), dropout_mask=unpack_mask(golden, p))
    for s in range(4):
        assert rel_err(out[("disp", s)], golden[p + "disp%d" % s]) < 1e-5
    for f in (-1, 1):
        assert rel_err(out[("cam_T_cam", 0, f)], golden[p + "T%d" % f]) < 1e-5
    torch.manual_seed(31)
    noise = [torch.randn(B, 2, H, W) * 0.00001 for _ in range(4)]
    losses = O.monodepth_loss(inputs, [out[("disp", s)] for s in range(4)],
                              {f: out[("cam_T_cam", 0, f)] for f in (-1, 1)}, [0, -1, 1], H, W, noise=noise)
    vals = np.array([losses["loss/%d" % s].item() for s in range(4)] + [losses["loss"].item()])
    np.testing.assert_allclose(vals, golden[p + "losses"], rtol=1e-5)
    losses["loss"].backward()
    names = [str(n) for n in golden[p + "grad_names"]]
    norms = np.array([osd[n].grad.norm().item() for n in names])
    np.testing.assert_allclose(norms, golden[p + "grad_norms"], rtol=2e-3, atol=1e-9)
    assert rel_err(osd["models.encoder.encoder.conv1.weight"].grad, golden[p + "grad_enc_conv1"]) < 2e-3
    assert rel_err(osd["models.encoder.encoder.bn1.running_mean"], golden[p + "bn1_running_mean"]) < 1e-5


@pytest.mark.parametrize("name", ["segdec_r50", "pad_r50"])
def test_seg_decoders_match_reference_golden(golden, contracts, name):
    """JointSegDepthDecoder / PAD restatements (models/joint_segmentation_depth_decoder.py) + cross_entropy2d."""
    from helpers import unpack_named_mask
    H, W, B = 64, 96, 2
    c = contracts[name]
    sd = O.synthetic_state_dict({k: torch.empty(s) for k, s in c["state_dict"].items()}, seed=2)
    osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    inputs = O.synthetic_inputs(B, H, W, seed=6, labels=True)
    p = "seg_%s_" % name
    mode = O.BNMode(True)
    feats = O.resnet_features(osd, "models.encoder.encoder.", inputs[("color_aug", 0, 0)], 50, [False, False, True], mode)
    da = c["cfg"]["depth_args"]
    if name == "segdec_r50":
        masks = {"aspp": unpack_named_mask(golden, p + "mask_aspp"), "head": unpack_named_mask(golden, p + "mask_head")}
        sem = O.joint_seg_depth_decoder(osd, "models.segmentation.", feats, mode=mode, dropout_masks=masks, depth_args=da)
        loss = O.cross_entropy2d(sem, inputs["lbl"])
    else:
        masks = {"depth": unpack_named_mask(golden, p + "mask_depth"), "seg": unpack_named_mask(golden, p + "mask_seg")}
        out = O.pad_decoder(osd, "models.mtl_decoder.", feats, mode=mode, dropout_masks=masks, depth_args=da)
        sem = out["semantics"]
        loss = (O.cross_entropy2d(sem, inputs["lbl"]) + O.cross_entropy2d(out["intermediate_semantics"], inputs["lbl"])) / 2
        assert rel_err(out["intermediate_semantics"], golden[p + "intermediate"]) < 1e-5
        for s in range(4):
            assert rel_err(out[("disp", s)], golden[p + "disp%d" % s]) < 1e-5
    assert rel_err(sem[:, :, ::4, ::4], golden[p + "semantics"]) < 1e-5
    assert abs(loss.item() - float(golden[p + "loss"])) < 1e-5 * abs(float(golden[p + "loss"]))
    loss.backward()
    names = [str(n) for n in golden[p + "grad_names"]]
    have = [n for n in names if osd[n].grad is not None]
    norms = np.array([osd[n].grad.norm().item() for n in have])
    ref = np.array([golden[p + "grad_norms"][names.index(n)] for n in have])
    np.testing.assert_allclose(norms, ref, rtol=5e-3, atol=1e-9)
    # parameters the reference gave a (zero) gradient that the functional oracle never touched must be zero there too
    rest = [golden[p + "grad_norms"][names.index(n)] for n in names if n not in have]
    assert all(v == 0 for v in rest), rest[:3]
