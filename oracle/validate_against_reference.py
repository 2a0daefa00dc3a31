"""Pins oracle/segsde_oracle.py against the UNMODIFIED reference (importable only in the build
container: /root/reference).  Run:  python oracle/validate_against_reference.py
Prints max abs/rel differences per quantity and exits non-zero when any exceeds its tolerance.
The result of the last run is recorded in oracle/VALIDATION.md.
"""
import contextlib
import io
import os
import sys

import torch

REF = os.environ.get("SEGSDE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present at %s (this script only runs in the build container)" % REF)
    sys.path.insert(0, REF)
    from configs.machine_config import MachineConfig
    MachineConfig("ws")
    import models as ref_models      # noqa
    import loss as ref_loss          # noqa
    return ref_models, ref_loss


def ref_model_cfg(backbone, H, W, rswd=(False, False, True)):
    return {
        "arch": "joint_segmentation_depth", "backbone_name": backbone,
        "replace_stride_with_dilation": list(rswd),
        "segmentation_name": None, "segmentation_args": None,
        "depth_args": {"intermediate_aspp": True, "aspp_rates": [6, 12, 18], "n_upconv": 4,
                       "num_ch_dec": [64, 128, 128, 256, 256], "max_scale_size": [H, W]},
        "pose_model_input": "pairs", "backbone_pretraining": "none", "depth_pretraining": "none",
        "pose_pretraining": "none", "freeze_backbone": False, "freeze_depth": False, "freeze_pose": False,
        "freeze_segmentation": True, "disable_monodepth": False, "disable_pose": False,
        "enable_imnet_encoder": False, "provide_uncropped_for_pose": False,
        "frame_ids": [0, -1, 1], "num_scales": 4, "height": H, "width": W, "crop_h": H, "crop_w": W,
    }


LOSS_KW = dict(min_depth=0.1, max_depth=100, test_min_depth=1e-3, test_max_depth=80, disparity_smoothness=1e-3,
               no_ssim=False, avg_reprojection=False, disable_automasking=False)


def maxdiff(a, b):
    d = (a - b).abs().max().item()
    r = d / (b.abs().max().item() + 1e-12)
    return d, r


def main():
    import segsde_oracle as O
    ref_models, ref_loss = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ok = True
    report = []

    def chk(name, a, b, tol):
        nonlocal ok
        d, r = maxdiff(a.detach().float(), b.detach().float())
        good = r <= tol
        ok &= good
        report.append("%-44s max|d|=%.3e rel=%.3e tol=%.1e %s" % (name, d, r, tol, "ok" if good else "FAIL"))
        print(report[-1])

    # ---- loss only: random disparities + poses ------------------------------------------------
    B, H, W = 2, 64, 96
    inputs = O.synthetic_inputs(B, H, W, seed=7)
    g = torch.Generator().manual_seed(3)
    disps = [torch.rand(B, 1, H >> s, W >> s, generator=g).mul(0.6).add(0.2).requires_grad_() for s in range(4)]
    Ts = {}
    for f in (-1, 1):
        aa = (torch.randn(B, 1, 3, generator=g) * 0.01)
        tr = (torch.randn(B, 1, 3, generator=g) * 0.05)
        Ts[f] = O.transformation_from_parameters(aa, tr, invert=f < 0).requires_grad_()
    for variant in ({}, {"no_ssim": True}, {"avg_reprojection": True}, {"disable_automasking": True}):
        kw = dict(LOSS_KW)
        kw.update(variant)
        rl = ref_loss.MonodepthLoss(num_scales=4, frame_ids=[0, -1, 1], height=H, width=W, batch_size=B, **kw)
        outputs = {("disp", s): disps[s] for s in range(4)}
        outputs.update({("cam_T_cam", 0, f): Ts[f] for f in (-1, 1)})
        torch.manual_seed(11)
        rl.generate_images_pred(inputs, outputs)
        rlosses = rl.compute_losses(inputs, outputs)
        rg = torch.autograd.grad(rlosses["loss"], disps + [Ts[-1], Ts[1]])
        torch.manual_seed(11)
        nf = 1 if kw["avg_reprojection"] else 2
        noise = [torch.randn(B, nf, H, W) * 0.00001 for _ in range(4)]
        extras = {}
        ol = O.monodepth_loss(inputs, disps, Ts, [0, -1, 1], H, W, kw["min_depth"], kw["max_depth"],
                              kw["disparity_smoothness"], kw["no_ssim"], kw["avg_reprojection"],
                              kw["disable_automasking"], noise, extras)
        og = torch.autograd.grad(ol["loss"], disps + [Ts[-1], Ts[1]])
        tag = "loss[%s] " % (",".join(variant) or "default")
        for k in rlosses:
            chk(tag + k, ol[k], rlosses[k], 1e-6)
        for i, (a, b) in enumerate(zip(og, rg)):
            chk(tag + "grad%d" % i, a, b, 1e-5)
        chk(tag + "color(-1,0)", extras[("color", -1, 0)], outputs[("color", -1, 0)], 1e-6)
        chk(tag + "sample(1,2)", extras[("sample", 1, 2)], outputs[("sample", 1, 2)], 1e-6)

    # ---- CE -------------------------------------------------------------------------------------
    lg = torch.randn(2, 19, 32, 48, generator=g)
    tgt = torch.randint(0, 19, (2, 32, 48), generator=g)
    tgt[0, :4] = 250
    pw = torch.rand(2, 32, 48, generator=g)
    from loss.loss import cross_entropy2d as ref_ce
    chk("cross_entropy2d", O.cross_entropy2d(lg, tgt), ref_ce(lg, tgt), 1e-6)
    chk("cross_entropy2d pixel_weights", O.cross_entropy2d(lg, tgt, pixel_weights=pw), ref_ce(lg, tgt, pixel_weights=pw), 1e-6)
    lg_small = torch.randn(2, 19, 16, 24, generator=g)
    chk("cross_entropy2d resized", O.cross_entropy2d(lg_small, tgt), ref_ce(lg_small, tgt), 1e-6)

    # ---- pose geometry ---------------------------------------------------------------------------
    from models.monodepth_layers import transformation_from_parameters as ref_tfp
    aa, tr = torch.randn(3, 1, 3, generator=g) * 0.1, torch.randn(3, 1, 3, generator=g)
    for inv in (False, True):
        chk("transformation_from_parameters inv=%s" % inv, O.transformation_from_parameters(aa, tr, inv),
            ref_tfp(aa, tr, inv), 1e-6)

    # ---- full model forward/backward (train-mode BN, dropout replayed) ---------------------------
    for backbone, (H, W) in (("resnet18", (64, 128)), ("resnet50", (64, 96))):
        nl = int(backbone[6:])
        rswd = [False, False, nl >= 50]      # torchvision BasicBlock has no dilation support
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref_models.get_model(ref_model_cfg(backbone, H, W, rswd), 19)
        sd = O.synthetic_state_dict(model.state_dict(), seed=1)
        model.load_state_dict(sd)
        model.train()
        inputs = O.synthetic_inputs(B, H, W, seed=5)
        torch.manual_seed(21)
        with contextlib.redirect_stdout(io.StringIO()):
            rout = model(inputs)
        torch.manual_seed(21)   # the only RNG consumer in this forward is the ASPP dropout (model_parts.py:25)
        osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
        cfg = {"num_layers": nl, "rswd": rswd, "frame_ids": [0, -1, 1]}
        # recover dropout mask from the reference activations: upconv(4,0) output == 0 where dropped
        aspp_ref = {}
        h = model.models["depth"].convs[("upconv", 4, 0)].register_forward_hook(lambda m, i, o: aspp_ref.setdefault("y", o))
        torch.manual_seed(21)
        with contextlib.redirect_stdout(io.StringIO()):
            model.load_state_dict(sd)     # reset running stats mutated by the first forward
            rout = model(inputs)
        h.remove()
        mask = (aspp_ref["y"] != 0).float()
        oout = O.model_forward(osd, inputs, cfg, O.BNMode(True), dropout_mask=mask)
        tag = backbone + " "
        for s in range(4):
            chk(tag + "disp%d" % s, oout[("disp", s)], rout[("disp", s)], 2e-5)
        chk(tag + "bottleneck", oout["bottleneck"], rout["bottleneck"], 2e-5)
        for f in (-1, 1):
            chk(tag + "cam_T_cam %d" % f, oout[("cam_T_cam", 0, f)], rout[("cam_T_cam", 0, f)], 1e-5)
        chk(tag + "running_mean(bn1)", osd["models.encoder.encoder.bn1.running_mean"],
            model.state_dict()["models.encoder.encoder.bn1.running_mean"], 1e-5)
        # loss + gradients through the whole net
        rl = ref_loss.MonodepthLoss(num_scales=4, frame_ids=[0, -1, 1], height=H, width=W, batch_size=B, **LOSS_KW)
        torch.manual_seed(31)
        rl.generate_images_pred(inputs, rout)
        rloss = rl.compute_losses(inputs, rout)["loss"]
        rloss.backward()
        torch.manual_seed(31)
        noise = [torch.randn(B, 2, H, W) * 0.00001 for _ in range(4)]
        oloss = O.monodepth_loss(inputs, [oout[("disp", s)] for s in range(4)],
                                 {f: oout[("cam_T_cam", 0, f)] for f in (-1, 1)}, [0, -1, 1], H, W, noise=noise)["loss"]
        oloss.backward()
        chk(tag + "loss", oloss, rloss, 1e-5)
        worst = 0.0
        for name, p in model.named_parameters():
            if p.grad is None:
                continue
            og = osd[name].grad
            d = (og - p.grad).abs().max().item() / (p.grad.abs().max().item() + 1e-12)
            worst = max(worst, d)
        good = worst < 2e-3
        ok &= good
        report.append("%-44s worst rel=%.3e tol=2.0e-03 %s" % (tag + "param grads", worst, "ok" if good else "FAIL"))
        print(report[-1])

    with open(os.path.join(HERE, "VALIDATION.md"), "w") as f:
        f.write("# Oracle vs unmodified reference (%s)\n\n" % REF)
        f.write("torch %s, CPU fp32. Generated by oracle/validate_against_reference.py.\n\n```\n" % torch.__version__)
        f.write("\n".join(report) + "\n```\n\nRESULT: %s\n" % ("ALL OK" if ok else "FAILURES"))
        f.write("\nThe joint segmentation decoders (JointSegDepthDecoder, PAD) and cross_entropy2d variants are pinned through "
                "the committed reference outputs instead: tests/golden/make_golden.py runs the unmodified reference on "
                "seeded inputs, tests/test_oracle_golden.py replays them against this oracle (CPU, any machine).\n")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
