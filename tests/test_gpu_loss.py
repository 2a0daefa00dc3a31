"""GPU: the fused loss kernels (through the drop-in Python API -> ctypes -> C-ABI) against the CPU oracle and
the committed reference outputs.  Tolerances (fp32): loss scalars 2e-5 rel, gradients 2e-4 rel of the max."""
import numpy as np
import pytest
import torch

import segsde_oracle as O
from helpers import LOSS_KW, grad_close, loss_case, loss_noise, rel_err

pytestmark = pytest.mark.gpu


def _cuda_inputs(inputs):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inputs.items()}


def _mine(kw, B, H, W, **extra):
    from improving_segmentation_with_selfsupervised_depth_b200.loss import MonodepthLoss
    return MonodepthLoss(num_scales=4, frame_ids=[0, -1, 1], height=H, width=W, batch_size=B, **kw, **extra)


@pytest.mark.parametrize("variant", ["default", "no_ssim", "avg_reprojection", "disable_automasking"])
def test_monodepth_loss_vs_reference_golden(golden, variant):
    B, H, W, inputs, disps, Ts = loss_case(golden)
    kw = dict(LOSS_KW)
    if variant != "default":
        kw[variant] = True
    ml = _mine(kw, B, H, W, materialize_outputs=True)
    ml.replay_noise = loss_noise(B, H, W, kw["avg_reprojection"])
    gin = _cuda_inputs(inputs)
    gd = [d.cuda().requires_grad_() for d in disps]
    gT = {f: t.cuda().requires_grad_() for f, t in Ts.items()}
    outputs = {("disp", s): gd[s] for s in range(4)}
    outputs.update({("cam_T_cam", 0, f): gT[f] for f in (-1, 1)})
    ml.generate_images_pred(gin, outputs)
    losses = ml.compute_losses(gin, outputs)
    vals = np.array([losses["loss/%d" % s].item() for s in range(4)] + [losses["loss"].item()])
    np.testing.assert_allclose(vals, golden["loss_%s_values" % variant], rtol=2e-5)
    grads = torch.autograd.grad(losses["loss"], gd + [gT[-1], gT[1]])
    for i, g in enumerate(grads):
        if i < 4:
            assert grad_close(g, golden["loss_%s_grad%d" % (variant, i)], 2e-4), i
        else:
            assert rel_err(g, golden["loss_%s_grad%d" % (variant, i)]) < 1e-2, i
    if variant == "default":
        assert rel_err(outputs[("color", -1, 0)], golden["loss_default_color_m1_s0"]) < 1e-5
        assert rel_err(outputs[("sample", 1, 2)], golden["loss_default_sample_p1_s2"]) < 1e-5
        sel = outputs["identity_selection/0"].cpu().numpy()
        assert (sel != golden["loss_default_identsel_s0"]).mean() < 1e-3


@pytest.mark.parametrize("hw", [(33, 70), (192, 640)])
def test_monodepth_loss_vs_oracle_shapes(hw):
    """Ragged (non tile-multiple) and the config-1 size; per-scale losses weighted individually."""
    H, W = hw
    B = 2
    inputs = O.synthetic_inputs(B, H, W, seed=9)
    g = torch.Generator().manual_seed(4)
    disps = [torch.rand(B, 1, max(H >> s, 1), max(W >> s, 1), generator=g).mul(0.6).add(0.2) for s in range(4)]
    Ts = {}
    for f in (-1, 1):
        Ts[f] = O.transformation_from_parameters(torch.randn(B, 1, 3, generator=g) * 0.01,
                                                 torch.randn(B, 1, 3, generator=g) * 0.05, invert=f < 0)
    noise = [torch.randn(B, 2, H, W, generator=g) * 1e-5 for _ in range(4)]
    wts = torch.tensor([0.3, 1.7, 0.5, 2.0, 1.0])
    cd = [d.clone().requires_grad_() for d in disps]
    cT = {f: t.clone().requires_grad_() for f, t in Ts.items()}
    ol = O.monodepth_loss(inputs, cd, cT, [0, -1, 1], H, W, noise=noise)
    ovec = torch.stack([ol["loss/%d" % s] for s in range(4)] + [ol["loss"]])
    (ovec * wts).sum().backward()
    ml = _mine(dict(LOSS_KW), B, H, W)
    ml.replay_noise = noise
    gd = [d.cuda().requires_grad_() for d in disps]
    gT = {f: t.cuda().requires_grad_() for f, t in Ts.items()}
    outputs = {("disp", s): gd[s] for s in range(4)}
    outputs.update({("cam_T_cam", 0, f): gT[f] for f in (-1, 1)})
    gin = _cuda_inputs(inputs)
    ml.generate_images_pred(gin, outputs)
    gl = ml.compute_losses(gin, outputs)
    gvec = torch.stack([gl["loss/%d" % s] for s in range(4)] + [gl["loss"]])
    assert rel_err(gvec, ovec) < 2e-5
    (gvec * wts.cuda()).sum().backward()
    for s in range(4):
        assert grad_close(gd[s].grad, cd[s].grad, 2e-4), s
    for f in (-1, 1):
        assert rel_err(gT[f].grad, cT[f].grad) < 3e-2, f     # sums over all pixels incl. the events above


def test_loss_without_grad_and_philox_noise():
    """Forward-only variant (1-px halo kernel) equals the gradient variant; in-kernel Philox noise changes the
    loss by less than the noise scale."""
    B, H, W = 2, 64, 96
    inputs = _cuda_inputs(O.synthetic_inputs(B, H, W, seed=2))
    g = torch.Generator().manual_seed(8)
    disps = [torch.rand(B, 1, H >> s, W >> s, generator=g).mul(0.6).add(0.2).cuda() for s in range(4)]
    T = {f: torch.eye(4).repeat(B, 1, 1).cuda() for f in (-1, 1)}
    T[-1][:, 0, 3] = 0.05
    T[1][:, 0, 3] = -0.05
    noise = [torch.randn(B, 2, H, W, generator=g) * 1e-5 for _ in range(4)]

    def run(req, replay):
        ml = _mine(dict(LOSS_KW), B, H, W)
        ml.replay_noise = noise if replay else None
        out = {("disp", s): disps[s].clone().requires_grad_(req) for s in range(4)}
        out.update({("cam_T_cam", 0, f): T[f] for f in (-1, 1)})
        ml.generate_images_pred(inputs, out)
        return ml.compute_losses(inputs, out)["loss"].item()

    a, b, c = run(False, True), run(True, True), run(True, False)
    assert abs(a - b) < 1e-6 * abs(b)
    assert abs(c - b) < 1e-4


def test_cross_entropy_vs_reference_golden(golden):
    from improving_segmentation_with_selfsupervised_depth_b200.loss.loss import cross_entropy2d
    lg = torch.from_numpy(golden["ce_logits"]).cuda().requires_grad_()
    tgt, pw = torch.from_numpy(golden["ce_target"]).cuda(), torch.from_numpy(golden["ce_pw"]).cuda()
    l0 = cross_entropy2d(input=lg, target=tgt)
    assert abs(l0.item() - float(golden["ce_loss"])) < 2e-6
    assert rel_err(torch.autograd.grad(l0, lg)[0], golden["ce_grad"]) < 1e-5
    l1 = cross_entropy2d(lg, tgt, pixel_weights=pw)
    assert abs(l1.item() - float(golden["ce_pw_loss"])) < 2e-6
    assert rel_err(torch.autograd.grad(l1, lg)[0], golden["ce_pw_grad"]) < 1e-5
    ls = cross_entropy2d(torch.from_numpy(golden["ce_small_logits"]).cuda(), tgt)
    assert abs(ls.item() - float(golden["ce_small_loss"])) < 1e-5
    # all pixels ignored -> NaN like F.cross_entropy
    allign = torch.full_like(tgt, 250)
    assert torch.isnan(cross_entropy2d(lg, allign))


def test_geometry_layers_vs_oracle():
    from improving_segmentation_with_selfsupervised_depth_b200.models import monodepth_layers as L
    B, H, W = 2, 24, 40
    g = torch.Generator().manual_seed(1)
    inputs = O.synthetic_inputs(B, H, W, seed=3)
    depth = torch.rand(B, 1, H, W, generator=g) * 10 + 1
    T = O.transformation_from_parameters(torch.randn(B, 1, 3, generator=g) * 0.02, torch.randn(B, 1, 3, generator=g) * 0.1)
    pts = O.backproject(depth, inputs[("inv_K", 0)])
    grid = O.project(pts, inputs[("K", 0)], T, H, W)
    mp = L.BackprojectDepth(B, H, W)(depth.cuda(), inputs[("inv_K", 0)].cuda())
    assert rel_err(mp, pts) < 1e-6
    mg = L.Project3D(B, H, W)(mp, inputs[("K", 0)].cuda(), T.cuda())
    assert rel_err(mg, grid) < 1e-5
    x, y = inputs[("color", 0, 0)], inputs[("color", 1, 0)]
    assert rel_err(L.SSIM()(x.cuda(), y.cuda()), O.ssim(x, y)) < 1e-4
    aa, tr = torch.randn(3, 1, 3, generator=g) * 0.1, torch.randn(3, 1, 3, generator=g)
    for inv in (False, True):
        a, t = aa.clone().requires_grad_(), tr.clone().requires_grad_()
        ref = O.transformation_from_parameters(a, t, inv)
        wgt = torch.randn(3, 4, 4, generator=g)
        (ref * wgt).sum().backward()
        ga, gt = aa.cuda().requires_grad_(), tr.cuda().requires_grad_()
        mine = L.transformation_from_parameters(ga, gt, inv)
        assert rel_err(mine, ref) < 1e-6
        (mine * wgt.cuda()).sum().backward()
        assert rel_err(ga.grad, a.grad) < 1e-4 and rel_err(gt.grad, t.grad) < 1e-5
