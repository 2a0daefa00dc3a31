"""Generates the committed golden fixtures from the UNMODIFIED reference (only runnable where
/root/reference exists).  The reference has no golden vectors of its own, so these are
outputs of the reference itself on seeded synthetic inputs:

    python tests/golden/make_golden.py

Weights are not stored: both sides rebuild them with oracle.segsde_oracle.synthetic_state_dict
(a recipe over sorted state_dict keys + shapes + seed).  Inputs come from synthetic_inputs(seed).
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.dont_write_bytecode = True
import segsde_oracle as O                                    # noqa: E402
from validate_against_reference import LOSS_KW, import_reference, ref_model_cfg   # noqa: E402


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def main():
    ref_models, ref_loss = import_reference()
    torch.set_num_threads(8)
    out = {}

    # ---- 1. state_dict key/shape contracts ------------------------------------------------------
    H, W = 64, 96
    contracts = {}
    cfgs = {"mono_r50": ref_model_cfg("resnet50", H, W), "mono_r18": ref_model_cfg("resnet18", H, W, (False,) * 3)}
    pad = ref_model_cfg("resnet50", H, W)
    pad.update({"segmentation_name": "mtl_pad", "freeze_segmentation": False,
                "segmentation_args": {"weights": "none", "output_stride": 1, "distillation_layer": 7,
                                      "side_output": True, "final_layer": 9}})
    cfgs["pad_r50"] = pad
    jsd = ref_model_cfg("resnet50", H, W)
    jsd.update({"segmentation_name": "joint_seg_depth_dec", "freeze_segmentation": False,
                "segmentation_args": {"weights": "none", "layers": [9], "output_stride": 1}})
    cfgs["segdec_r50"] = jsd
    for name, cfg in cfgs.items():
        with quiet():
            m = ref_models.get_model(cfg, 19)
        contracts[name] = {"cfg": cfg, "state_dict": {k: list(v.shape) for k, v in m.state_dict().items()},
                           "trainable": [k for k, p in m.named_parameters() if p.requires_grad]}
    with open(os.path.join(HERE, "state_dict_contracts.json"), "w") as f:
        json.dump(contracts, f, indent=0, sort_keys=True)

    # ---- 2. loss-level golden ----------------------------------------------------------------------
    B, H, W = 2, 64, 96
    inputs = O.synthetic_inputs(B, H, W, seed=7)
    g = torch.Generator().manual_seed(3)
    disps = [torch.rand(B, 1, H >> s, W >> s, generator=g).mul(0.6).add(0.2).requires_grad_() for s in range(4)]
    Ts = {}
    for f in (-1, 1):
        aa, tr = torch.randn(B, 1, 3, generator=g) * 0.01, torch.randn(B, 1, 3, generator=g) * 0.05
        Ts[f] = O.transformation_from_parameters(aa, tr, invert=f < 0).requires_grad_()
    for i, s in enumerate(range(4)):
        out["loss_disp%d" % s] = disps[s].detach().numpy()
    out["loss_T-1"], out["loss_T1"] = Ts[-1].detach().numpy(), Ts[1].detach().numpy()
    for variant in ("default", "no_ssim", "avg_reprojection", "disable_automasking"):
        kw = dict(LOSS_KW)
        if variant != "default":
            kw[variant] = True
        rl = ref_loss.MonodepthLoss(num_scales=4, frame_ids=[0, -1, 1], height=H, width=W, batch_size=B, **kw)
        outputs = {("disp", s): disps[s] for s in range(4)}
        outputs.update({("cam_T_cam", 0, f): Ts[f] for f in (-1, 1)})
        torch.manual_seed(11)
        rl.generate_images_pred(inputs, outputs)
        rlosses = rl.compute_losses(inputs, outputs)
        grads = torch.autograd.grad(rlosses["loss"], disps + [Ts[-1], Ts[1]])
        out["loss_%s_values" % variant] = np.array([rlosses["loss/%d" % s].item() for s in range(4)] +
                                                   [rlosses["loss"].item()], dtype=np.float64)
        for i, gr in enumerate(grads):
            out["loss_%s_grad%d" % (variant, i)] = gr.numpy()
        if variant == "default":
            out["loss_default_color_m1_s0"] = outputs[("color", -1, 0)].detach().numpy()
            out["loss_default_sample_p1_s2"] = outputs[("sample", 1, 2)].detach().numpy()
            out["loss_default_identsel_s0"] = outputs["identity_selection/0"].numpy()

    # ---- 3. cross entropy --------------------------------------------------------------------------
    from loss.loss import cross_entropy2d as ref_ce
    g = torch.Generator().manual_seed(5)
    lg = torch.randn(2, 19, 32, 48, generator=g).requires_grad_()
    tgt = torch.randint(0, 19, (2, 32, 48), generator=g)
    tgt[0, :4] = 250
    pw = torch.rand(2, 32, 48, generator=g)
    out["ce_logits"], out["ce_target"], out["ce_pw"] = lg.detach().numpy(), tgt.numpy(), pw.numpy()
    l0 = ref_ce(lg, tgt)
    out["ce_loss"], out["ce_grad"] = l0.item(), torch.autograd.grad(l0, lg)[0].numpy()
    l1 = ref_ce(lg, tgt, pixel_weights=pw)
    out["ce_pw_loss"], out["ce_pw_grad"] = l1.item(), torch.autograd.grad(l1, lg)[0].numpy()
    lgs = torch.randn(2, 19, 16, 24, generator=g)
    out["ce_small_logits"], out["ce_small_loss"] = lgs.numpy(), ref_ce(lgs, tgt).item()

    # ---- 4. model-level golden: fwd + loss + grads, train-mode BN ----------------------------------
    for name, (H, W) in (("mono_r18", (64, 128)), ("mono_r50", (64, 96))):
        cfg = dict(cfgs[name])
        cfg.update({"height": H, "width": W, "crop_h": H, "crop_w": W})
        cfg["depth_args"] = dict(cfg["depth_args"], max_scale_size=[H, W])
        with quiet():
            model = ref_models.get_model(cfg, 19)
        sd = O.synthetic_state_dict(model.state_dict(), seed=1)
        model.load_state_dict(sd)
        model.train()
        inputs = O.synthetic_inputs(B, H, W, seed=5)
        rec = {}
        hook = model.models["depth"].convs[("upconv", 4, 0)].register_forward_hook(
            lambda m, i, o: rec.setdefault("y", o))
        torch.manual_seed(21)
        with quiet():
            rout = model(inputs)
        hook.remove()
        mask = (rec["y"] != 0)
        rl = ref_loss.MonodepthLoss(num_scales=4, frame_ids=[0, -1, 1], height=H, width=W, batch_size=B, **LOSS_KW)
        torch.manual_seed(31)
        rl.generate_images_pred(inputs, rout)
        losses = rl.compute_losses(inputs, rout)
        losses["loss"].backward()
        p = "model_%s_" % name
        out[p + "dropout_mask"] = np.packbits(mask.numpy().reshape(-1))
        out[p + "dropout_shape"] = np.array(mask.shape)
        for s in range(4):
            out[p + "disp%d" % s] = rout[("disp", s)].detach().numpy()
        for f in (-1, 1):
            out[p + "T%d" % f] = rout[("cam_T_cam", 0, f)].detach().numpy()
        out[p + "losses"] = np.array([losses["loss/%d" % s].item() for s in range(4)] + [losses["loss"].item()])
        names = [n for n, q in model.named_parameters() if q.grad is not None]
        out[p + "grad_names"] = np.array(names)
        out[p + "grad_norms"] = np.array([dict(model.named_parameters())[n].grad.norm().item() for n in names])
        out[p + "grad_enc_conv1"] = dict(model.named_parameters())["models.encoder.encoder.conv1.weight"].grad.numpy()
        out[p + "grad_pose3"] = dict(model.named_parameters())["models.pose.net.3.weight"].grad.numpy()
        out[p + "bn1_running_mean"] = model.state_dict()["models.encoder.encoder.bn1.running_mean"].numpy()
        out[p + "bn1_running_var"] = model.state_dict()["models.encoder.encoder.bn1.running_var"].numpy()
    # ---- 5. segmentation decoders (JointSegDepthDecoder, PAD): fwd + CE loss + grads -----------------
    from loss.loss import cross_entropy2d as ref_ce2
    H, W = 64, 96
    for name in ("segdec_r50", "pad_r50"):
        cfg = dict(cfgs[name])
        with quiet():
            model = ref_models.get_model(cfg, 19)
        sd = O.synthetic_state_dict(model.state_dict(), seed=2)
        model.load_state_dict(sd)
        model.train()
        inputs = O.synthetic_inputs(B, H, W, seed=6, labels=True)
        rec = {}

        def grab(key):
            def hook(m, i, o):
                rec[key] = (i[0].detach().clone(), o.detach().clone())
            return hook
        hooks = []
        if name == "segdec_r50":
            dec = model.models["segmentation"]
            hooks.append(dec.unet_dec.convs[("upconv", 4, 0)].project[3].register_forward_hook(grab("aspp")))
            hooks.append(dec.head[4].register_forward_hook(grab("head")))
        else:
            dec = model.models["mtl_decoder"]
            hooks.append(dec.depth_dec.convs[("upconv", 4, 0)].project[3].register_forward_hook(grab("depth")))
            hooks.append(dec.seg_dec.convs[("upconv", 4, 0)].project[3].register_forward_hook(grab("seg")))
        torch.manual_seed(41)
        with quiet():
            rout = model(inputs)
        for h in hooks:
            h.remove()
        p = "seg_%s_" % name
        for key, (i_, o_) in rec.items():
            mask = torch.where(i_ != 0, (o_ != 0), torch.ones_like(o_, dtype=torch.bool))
            out[p + "mask_" + key] = np.packbits(mask.numpy().reshape(-1))
            out[p + "mask_" + key + "_shape"] = np.array(mask.shape)
        loss = ref_ce2(input=rout["semantics"], target=inputs["lbl"])
        if "intermediate_semantics" in rout:
            loss = (loss + ref_ce2(input=rout["intermediate_semantics"], target=inputs["lbl"])) / 2
            out[p + "intermediate"] = rout["intermediate_semantics"].detach().numpy()
        loss.backward()
        out[p + "semantics"] = rout["semantics"].detach().numpy()[:, :, ::4, ::4].copy()
        out[p + "loss"] = loss.item()
        names = [n for n, q in model.named_parameters() if q.grad is not None]
        out[p + "grad_names"] = np.array(names)
        out[p + "grad_norms"] = np.array([dict(model.named_parameters())[n].grad.norm().item() for n in names])
        if name == "pad_r50":
            for s_ in range(4):
                out[p + "disp%d" % s_] = rout[("disp", s_)].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"),
          os.path.getsize(os.path.join(HERE, "reference_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
