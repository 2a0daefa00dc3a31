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
