"""berhu (loss/loss.py:5-15) and pixel_wise_entropy (loss/loss.py:40-47): the oracle against outputs of the
unmodified reference (tests/golden/loss_extra_golden.npz, CPU) and the CUDA kernels against both (GPU).
Tolerances: fp32 reductions over 5k-20k elements -> 2e-6 relative on the loss, 1e-5 on gradients / entropy maps."""
import os
import sys

import numpy as np
import pytest
import torch

import segsde_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_loss_extra import CASES, loss_extra_inputs      # noqa: E402


@pytest.fixture(scope="module")
def extra():
    return np.load(os.path.join(ROOT, "tests", "golden", "loss_extra_golden.npz"), allow_pickle=False)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("name,apply_log", CASES["berhu"])
def test_oracle_berhu_matches_reference(extra, name, apply_log):
    disp, target, mask, _ = loss_extra_inputs()
    x = disp.clone().requires_grad_()
    loss = O.berhu(x, target, mask, apply_log=apply_log)
    loss.backward()
    assert rel(loss.detach(), extra["berhu/%s/loss" % name]) < 1e-6
    assert rel(x.grad, extra["berhu/%s/grad" % name]) < 1e-6


@pytest.mark.parametrize("name,norm", CASES["entropy"])
def test_oracle_entropy_matches_reference(extra, name, norm):
    logits = loss_extra_inputs()[3]
    assert rel(O.pixel_wise_entropy(logits, normalize=norm), extra["entropy/%s" % name]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name,apply_log", CASES["berhu"])
def test_gpu_berhu(extra, name, apply_log):
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    _, loss_pkg = P.install_dropin()
    from loss.loss import berhu
    disp, target, mask, _ = loss_extra_inputs()
    x = disp.cuda().requires_grad_()
    loss = berhu(x, target.cuda(), mask.cuda(), apply_log=apply_log)
    (3.0 * loss).backward()
    assert rel(loss.detach(), extra["berhu/%s/loss" % name]) < 2e-6
    assert rel(x.grad / 3.0, extra["berhu/%s/grad" % name]) < 1e-5
    # a larger, unmasked case against the oracle (mask = None is "all ones")
    g = torch.Generator().manual_seed(5)
    big, tgt = torch.rand(3, 1, 96, 160, generator=g), torch.rand(3, 1, 96, 160, generator=g) * 2
    xo = big.clone().requires_grad_()
    lo = O.berhu(xo, tgt, torch.ones_like(big), apply_log=apply_log)
    lo.backward()
    xg = big.cuda().requires_grad_()
    lg = berhu(xg, tgt.cuda(), None, apply_log=apply_log)
    lg.backward()
    assert rel(lg.detach(), lo.detach()) < 2e-6
    assert rel(xg.grad, xo.grad) < 1e-5
    # all-equal input: zero loss and a zero (not NaN) gradient
    z = torch.full((1, 1, 8, 8), 0.5, device="cuda", requires_grad=True)
    lz = berhu(z, torch.full((1, 1, 8, 8), 0.5, device="cuda"), None)
    lz.backward()
    assert float(lz.detach()) == 0.0 and float(z.grad.abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name,norm", CASES["entropy"])
def test_gpu_pixel_wise_entropy(extra, name, norm):
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    P.install_dropin()
    from loss.loss import pixel_wise_entropy
    logits = loss_extra_inputs()[3]
    out = pixel_wise_entropy(logits.cuda(), normalize=norm)
    assert tuple(out.shape) == (2, 24, 40)
    assert rel(out, extra["entropy/%s" % name]) < 1e-5
    # a channels-last / strided input is handled as well (made contiguous)
    out2 = pixel_wise_entropy(logits.cuda().contiguous(memory_format=torch.channels_last), normalize=norm)
    assert rel(out2, extra["entropy/%s" % name]) < 1e-5
