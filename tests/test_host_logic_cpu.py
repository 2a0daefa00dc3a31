"""CPU: the Python host logic of the drop-in (`models/*.py` — encoder / decoder wiring, skip connections, exec_layer,
PAD cross gating, pose routing, BatchNorm bookkeeping, state_dict mapping) against the committed outputs of the
unmodified reference, with the kernel-backed `ops` entry points swapped for the torch-functional stand-ins of
tests/cpu_ops_emulation.py (test infrastructure; the product has no CPU route).  Tolerance 2e-4 (the GPU tests' own
activation bound): both sides are fp32, a random-weight ResNet-50 amplifies summation-order noise to ~5e-5."""
import contextlib
import io

import pytest
import torch

import cpu_ops_emulation as E
import segsde_oracle as O
from helpers import rel_err, unpack_mask, unpack_named_mask


def _model(contracts, name, cfg_update, seed):
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    models, _ = P.install_dropin()
    c = contracts[name]
    cfg = dict(c["cfg"])
    cfg.update(cfg_update)
    if "depth_args" in cfg and "height" in cfg_update:
        cfg["depth_args"] = dict(cfg["depth_args"], max_scale_size=[cfg_update["height"], cfg_update["width"]])
    with contextlib.redirect_stdout(io.StringIO()):
        m = models.get_model(cfg, 19)
    m.load_state_dict(O.synthetic_state_dict({k: torch.empty(s) for k, s in c["state_dict"].items()}, seed=seed))
    return m.train()


@pytest.mark.parametrize("name,hw", [("mono_r18", (64, 128)), ("mono_r50", (64, 96))])
def test_monodepth_model_wiring(golden, contracts, name, hw, monkeypatch):
    from improving_segmentation_with_selfsupervised_depth_b200.models.layers import Dropout
    E.install(monkeypatch)
    H, W = hw
    p = "model_%s_" % name
    model = _model(contracts, name, {"height": H, "width": W, "crop_h": H, "crop_w": W}, seed=1)
    for mod in model.modules():
        if isinstance(mod, Dropout):
            mod.replay_mask = unpack_mask(golden, p)
    inputs = O.synthetic_inputs(2, H, W, seed=5)
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    for s in range(4):
        assert out[("disp", s)].shape == golden[p + "disp%d" % s].shape
        assert rel_err(out[("disp", s)], golden[p + "disp%d" % s]) < 2e-4, s
    for f in (-1, 1):
        assert rel_err(out[("cam_T_cam", 0, f)], golden[p + "T%d" % f]) < 1e-4
    sd = model.state_dict()
    assert rel_err(sd["models.encoder.encoder.bn1.running_mean"], golden[p + "bn1_running_mean"]) < 1e-5
    assert rel_err(sd["models.encoder.encoder.bn1.running_var"], golden[p + "bn1_running_var"]) < 1e-5
    assert int(sd["models.encoder.encoder.bn1.num_batches_tracked"]) == 1
    # the gradient reaches every trainable parameter through the wiring
    sum(out[("disp", s)].mean() for s in range(4)).backward()
    assert all(q.grad is not None for n, q in model.named_parameters() if q.requires_grad and ".depth." in n)


@pytest.mark.parametrize("name", ["segdec_r50", "pad_r50"])
def test_segmentation_decoder_wiring(golden, contracts, name, monkeypatch):
    E.install(monkeypatch)
    from improving_segmentation_with_selfsupervised_depth_b200.loss.loss import cross_entropy2d
    H, W = 64, 96
    model = _model(contracts, name, {}, seed=2)
    p = "seg_%s_" % name
    if name == "segdec_r50":
        dec = model.models["segmentation"]
        dec.unet_dec.convs[("upconv", 4, 0)].project[3].replay_mask = unpack_named_mask(golden, p + "mask_aspp")
        dec.head[4].replay_mask = unpack_named_mask(golden, p + "mask_head")
    else:
        dec = model.models["mtl_decoder"]
        dec.depth_dec.convs[("upconv", 4, 0)].project[3].replay_mask = unpack_named_mask(golden, p + "mask_depth")
        dec.seg_dec.convs[("upconv", 4, 0)].project[3].replay_mask = unpack_named_mask(golden, p + "mask_seg")
    inputs = O.synthetic_inputs(2, H, W, seed=6, labels=True)
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    assert rel_err(out["semantics"][:, :, ::4, ::4], golden[p + "semantics"]) < 2e-4
    loss = cross_entropy2d(input=out["semantics"], target=inputs["lbl"])
    if "intermediate_semantics" in out:
        assert rel_err(out["intermediate_semantics"], golden[p + "intermediate"]) < 2e-4
        for s in range(4):
            assert rel_err(out[("disp", s)], golden[p + "disp%d" % s]) < 2e-4
        loss = (loss + cross_entropy2d(input=out["intermediate_semantics"], target=inputs["lbl"])) / 2
    assert abs(float(loss.detach()) - float(golden[p + "loss"])) < 2e-5 * abs(float(golden[p + "loss"]))


def test_eval_mode_and_frozen_backbone(contracts, monkeypatch):
    """Eval mode normalises with the running statistics (oracle BNMode(False)), `freeze_backbone` only clears
    requires_grad (joint_segmentation_depth.py:158-179) and the frozen encoder receives no gradient."""
    E.install(monkeypatch)
    H, W = 64, 96
    model = _model(contracts, "mono_r18", {"height": H, "width": W, "crop_h": H, "crop_w": W, "freeze_backbone": True}, seed=4)
    enc_params = [q for n, q in model.named_parameters() if n.startswith("models.encoder.")]
    assert enc_params and not any(q.requires_grad for q in enc_params)
    assert all(q.requires_grad for n, q in model.named_parameters() if n.startswith("models.depth."))
    inputs = O.synthetic_inputs(2, H, W, seed=9)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.eval()
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    ref = O.model_forward(sd, inputs, {"num_layers": 18, "rswd": [False] * 3, "frame_ids": [0, -1, 1]}, O.BNMode(False))
    for s in range(4):
        assert rel_err(out[("disp", s)], ref[("disp", s)]) < 2e-4
    assert rel_err(out[("cam_T_cam", 0, -1)], ref[("cam_T_cam", 0, -1)]) < 1e-4
    # eval mode must not touch the running statistics
    after = model.state_dict()
    assert all(torch.equal(after[k], v) for k, v in sd.items() if "running" in k or "num_batches" in k)
    model.train()
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    sum(out[("disp", s)].mean() for s in range(4)).backward()
    assert all(q.grad is None for q in enc_params)


def test_depth_decoder_partial_execution(contracts, monkeypatch):
    """`exec_layer` + injected `x` (how PAD drives its decoders in two halves) compose to the full pass."""
    E.install(monkeypatch)
    from improving_segmentation_with_selfsupervised_depth_b200.models.depth_decoder import DepthDecoder
    torch.manual_seed(0)
    enc_ch = [16, 32, 64, 128, 128]
    dec = DepthDecoder(enc_ch, range(4), [64, 96], num_ch_dec=[16, 32, 32, 64, 64], intermediate_aspp=True,
                       aspp_pooling=True, batch_norm=True).eval()
    feats = [torch.randn(2, c, 32 >> min(i, 3), 48 >> min(i, 3)) for i, c in enumerate(enc_ch)]   # dilated: f3, f4 share a size
    full = {k: v.clone() for k, v in dec(feats).items()}
    top = {k: v.clone() for k, v in dec(feats, exec_layer=[4, 3, 2]).items()}
    assert set(top) == {("upconv", 4), ("upconv", 3), ("upconv", 2), ("disp", 3), ("disp", 2)}
    rest = dec(feats, x=top[("upconv", 2)], exec_layer=[1, 0])
    assert set(rest) == {("upconv", 1), ("upconv", 0), ("disp", 1), ("disp", 0)}
    for k, v in {**top, **rest}.items():
        assert torch.allclose(v, full[k], atol=1e-6), k
    assert full[("upconv", 4)].shape[-2:] == feats[3].shape[-2:]          # no upsampling between equally sized stages
    assert full[("disp", 0)].shape == (2, 1, 64, 96)
    dec.enable_disparity = False
    assert not any(k[0] == "disp" for k in dec(feats))
