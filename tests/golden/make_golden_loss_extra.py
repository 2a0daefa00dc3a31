"""Golden fixtures for loss/loss.py's `berhu` and `pixel_wise_entropy`, produced by the UNMODIFIED reference (only
runnable where /root/reference exists):

    python tests/golden/make_golden_loss_extra.py     ->  tests/golden/loss_extra_golden.npz

Inputs are regenerated from seeds by `loss_extra_inputs` below (imported by the tests), only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES = {"berhu": [("plain", False), ("log", True)], "entropy": [("raw", False), ("norm", True)]}


def loss_extra_inputs():
    """Seeded inputs: a sigmoid-like disparity map with gradient, a pseudo-depth target, the bottom-10 % mask of
    train.py:491-493, and 19-class logits."""
    g = torch.Generator().manual_seed(2024)
    disp = torch.rand(2, 1, 40, 64, generator=g) * 0.8 + 0.1
    target = (disp + 0.3 * torch.randn(2, 1, 40, 64, generator=g)).clamp(0.01, 2.0)
    target[0, 0, 3, 5] = 3.0            # one outlier sets the maximum, most pixels fall below the switch point
    mask = torch.ones(2, 1, 40, 64)
    mask[:, :, 36:, :] = 0
    logits = torch.randn(2, 19, 24, 40, generator=g) * 3.0
    return disp, target, mask, logits


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    from loss.loss import berhu, pixel_wise_entropy          # the reference's own functions
    disp, target, mask, logits = loss_extra_inputs()
    out = {}
    for name, apply_log in CASES["berhu"]:
        x = disp.clone().requires_grad_()
        loss = berhu(x, target, mask, apply_log=apply_log)
        loss.backward()
        out["berhu/%s/loss" % name] = loss.detach().numpy()
        out["berhu/%s/grad" % name] = x.grad.numpy()
    for name, norm in CASES["entropy"]:
        out["entropy/%s" % name] = pixel_wise_entropy(logits, normalize=norm).numpy()
    np.savez_compressed(os.path.join(HERE, "loss_extra_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
