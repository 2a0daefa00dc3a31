"""Golden fixtures for the step-level ops of train.py, produced by calling the UNMODIFIED
reference's own functions (only runnable where /root/reference exists):

    python tests/golden/make_golden_train_ops.py     ->  tests/golden/train_ops_golden.npz

`Trainer` methods are called unbound with a SimpleNamespace standing in for `self` (they only read a few
attributes); matplotlib / kornia / ray are stubbed so that `import train` works.  Inputs are regenerated from seeds
by `train_ops_inputs` (imported by the tests); only outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
MARGIN, FT, CONSISTENCY = 0.03, 0.25, 1.5
EMA_CASES = [(0, 0.99), (7, 0.99), (500, 0.99)]         # (iteration, alpha_teacher)


def train_ops_inputs():
    g = torch.Generator().manual_seed(77)
    feats = torch.randn(2, 64, 6, 10, generator=g), torch.randn(2, 64, 6, 10, generator=g)
    depths = torch.rand(2, 1, 24, 40, generator=g)                      # the reference's depthcomp needs B = 2
    imgs = torch.rand(2, 3, 24, 40, generator=g)
    logits_t = torch.randn(2, 19, 24, 40, generator=g) * 4.0
    softmax_t = torch.softmax(logits_t, 1)
    softmax_t[0, :, 2, 3] = 0.0                                          # an all-zero pixel -> ignore label
    softmax_t[1, :, 5, 7] = 0.0
    softmax_t[0, :, 4, 4] = 0.0
    softmax_t[0, 3, 4, 4] = 0.5
    softmax_t[0, 9, 4, 4] = 0.5                                          # a tie: the first maximum wins
    logits_s = torch.randn(2, 19, 24, 40, generator=g)
    params = [torch.randn(s, generator=g) for s in ((64, 3, 7, 7), (64,), (19, 64, 1, 1), (70001,), (1,))]
    ema = [torch.randn(p.shape, generator=g) for p in params]
    return dict(feats=feats, depths=depths, imgs=imgs, softmax_t=softmax_t, logits_s=logits_s, params=params, ema=ema)


def import_reference_train():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    for name in ["matplotlib", "matplotlib.pyplot", "kornia", "ray", "ray.tune"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["ray"].tune = sys.modules["ray.tune"]
    from configs.machine_config import MachineConfig
    MachineConfig("ws")
    import train
    from loader import transformsgpu
    return train, transformsgpu


def main():
    train, transformsgpu = import_reference_train()
    x = train_ops_inputs()
    out = {}
    # T1: train.py:482
    a = x["feats"][0].clone().requires_grad_()
    d = torch.dist(a, x["feats"][1], p=2)
    d.backward()
    out["t1/dist"], out["t1/grad"] = d.detach().numpy(), a.grad.numpy()
    # T2: train.py:688-692 (inline code, restated verbatim on a copy), generate_mix_mask, transformsgpu.mix
    depths = x["depths"].clone()
    for j in range(depths.shape[0]):
        dmin, dmax = torch.min(depths[j]), torch.max(depths[j])
        depths[j] = torch.clamp(depths[j], dmin, dmax)
        depths[j] = (depths[j] - dmin) / (dmax - dmin)
    out["t2/depths_norm"] = depths.numpy()
    ns = types.SimpleNamespace(mix_mask="depthcomp", cfg={"training": {"batch_size": 2}}, depthcomp_margin=MARGIN,
                               depthcomp_foreground_threshold=FT, device="cpu")
    mask = train.Trainer.generate_mix_mask(ns, "depthcomp", None, x["imgs"], depths)
    out["t2/mask"] = mask.numpy()
    out["t2/mix_img"] = transformsgpu.mix(mask=mask, data=x["imgs"])[0].numpy()
    out["t2/mix_softmax"] = transformsgpu.mix(mask=mask, data=x["softmax_t"])[0].numpy()
    # T3: Trainer.calc_pseudo_label_loss (train.py:644-651)
    ns3 = types.SimpleNamespace(unlabeled_loader=types.SimpleNamespace(ignore_index=250), consistency_weight=CONSISTENCY,
                                device="cpu")
    s = x["logits_s"].clone().requires_grad_()
    loss, label = train.Trainer.calc_pseudo_label_loss(ns3, teacher_softmax=x["softmax_t"].clone(), student_logits=s)
    loss.backward()
    out["t3/loss"], out["t3/label"], out["t3/grad"] = loss.detach().numpy(), label.numpy(), s.grad.numpy()
    # T4: Trainer.update_ema_variables (train.py:346-358)
    for it, alpha in EMA_CASES:
        ema_model = types.SimpleNamespace(parameters=lambda e=[t.clone() for t in x["ema"]]: e)
        model = types.SimpleNamespace(parameters=lambda: x["params"])
        ns4 = types.SimpleNamespace(cfg={"training": {"save_monodepth_ema": False}, "model": {"segmentation_name": "none"}})
        train.Trainer.update_ema_variables(ns4, ema_model, model, alpha, it)
        for k, t in enumerate(ema_model.parameters()):
            out["t4/it%d/%d" % (it, k)] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "train_ops_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
