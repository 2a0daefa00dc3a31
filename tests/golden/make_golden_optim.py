"""Golden fixtures for the optimizer step / gradient clipping: the classes the UNMODIFIED
reference instantiates (utils/optimizers.py:get_optimizer -> torch.optim.Adam / SGD, train.py:291-295) and
torch.nn.utils.clip_grad_norm_ (train.py:521-524), run on the CPU for three steps on seeded parameters / gradients.

    python tests/golden/make_golden_optim.py     ->  tests/golden/optim_golden.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(64, 3, 7, 7), (64,), (19, 64, 1, 1), (70001,), (1,), (32, 16, 3, 3)]
CASES = {
    "adam": dict(name="adam", lr=1e-4),
    "adam_wd": dict(name="adam", lr=3e-3, weight_decay=1e-2, betas=(0.8, 0.95), eps=1e-6),
    "sgd": dict(name="sgd", lr=1e-2, momentum=0.9, weight_decay=5e-4),            # cityscapes_joint.yml
    "sgd_plain": dict(name="sgd", lr=5e-2),
    "sgd_nesterov": dict(name="sgd", lr=1e-2, momentum=0.8, nesterov=True, weight_decay=1e-3),
    "sgd_damp": dict(name="sgd", lr=1e-2, momentum=0.7, dampening=0.3),
}
STEPS = 3
CLIP = [("small", 10.0, 0.01), ("large", 10.0, 30.0)]        # (name, max_norm, gradient scale)
BIG = 3                                                      # index of the > 64K-element tensor (two chunks)
BIG_CASES = ("adam", "sgd", "clip_large")                    # only these keep it (fixture size)


def select(case, tensors):
    return tensors if case in BIG_CASES else [t for i, t in enumerate(tensors) if i != BIG]


def optim_inputs():
    g = torch.Generator().manual_seed(4242)
    params = [torch.randn(s, generator=g) for s in SHAPES]
    grads = [[torch.randn(s, generator=g) * (0.1 + 0.3 * k) for s in SHAPES] for k in range(STEPS)]
    return params, grads


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    from utils.optimizers import get_optimizer            # the reference's own factory
    params0, grads = optim_inputs()
    out = {}
    for case, opt_cfg in CASES.items():
        cls = get_optimizer({"training": {"optimizer": dict(opt_cfg)}})
        kw = {k: v for k, v in opt_cfg.items() if k != "name"}     # train.py:292-293
        ps = [torch.nn.Parameter(p.clone()) for p in select(case, params0)]
        opt = cls(ps, **kw)
        for k in range(STEPS):
            for p, g in zip(ps, select(case, grads[k])):
                p.grad = g.clone()
            opt.step()
        for i, p in enumerate(ps):
            out["%s/p%d" % (case, i)] = p.detach().numpy()
    for name, max_norm, scale in CLIP:
        ps = [torch.nn.Parameter(p.clone()) for p in select("clip_" + name, params0)]
        for p, g in zip(ps, select("clip_" + name, grads[0])):
            p.grad = g.clone() * scale
        total = torch.nn.utils.clip_grad_norm_(ps, max_norm)
        out["clip_%s/total" % name] = total.numpy()
        for i, p in enumerate(ps):
            out["clip_%s/g%d" % (name, i)] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "optim_golden.npz"), **out)
    print(len(out), "arrays")


if __name__ == "__main__":
    main()
