"""Optimizer step / gradient clipping: the multi-tensor kernels against three steps of the classes
the reference instantiates (torch.optim.Adam / SGD via utils/optimizers.py) and torch.nn.utils.clip_grad_norm_, as
stored by tests/golden/make_golden_optim.py.  fp32, same arithmetic: 2e-6 relative on parameters / gradients.
The CPU test re-runs torch on this machine against the stored arrays (the fixture itself is reproducible)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_optim import CASES, CLIP, STEPS, optim_inputs, select      # noqa: E402

KEY2OPT = {"adam": torch.optim.Adam, "sgd": torch.optim.SGD}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "optim_golden.npz"), allow_pickle=False)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def run(case, cls, device, channels_last=False):
    params0, grads = optim_inputs()
    kw = {k: v for k, v in CASES[case].items() if k != "name"}

    def put(t):
        t = t.clone().to(device)
        return t.contiguous(memory_format=torch.channels_last) if (channels_last and t.dim() == 4) else t
    ps = [torch.nn.Parameter(put(p)) for p in select(case, params0)]
    opt = cls(ps, **kw)
    for k in range(STEPS):
        for p, g in zip(ps, select(case, grads[k])):
            p.grad = put(g)
        opt.step()
    return ps, opt


@pytest.mark.parametrize("case", list(CASES))
def test_fixture_reproducible_with_torch_cpu(gold, case):
    ps, _ = run(case, KEY2OPT[CASES[case]["name"]], "cpu")
    for i, p in enumerate(ps):
        assert rel(p.detach(), gold["%s/p%d" % (case, i)]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("case", list(CASES))
def test_gpu_optimizers(gold, case, channels_last):
    from improving_segmentation_with_selfsupervised_depth_b200 import optim
    cls = {"adam": optim.Adam, "sgd": optim.SGD}[CASES[case]["name"]]
    ps, opt = run(case, cls, "cuda", channels_last)
    for i, p in enumerate(ps):
        assert rel(p.detach(), gold["%s/p%d" % (case, i)]) < 2e-6, (case, i)
    sd = opt.state_dict()                     # torch's layout: state per parameter index, same keys
    st0 = sd["state"][0]
    if CASES[case]["name"] == "adam":
        assert set(st0) == {"step", "exp_avg", "exp_avg_sq"} and int(st0["step"]) == STEPS
    elif CASES[case].get("momentum", 0):
        assert set(st0) == {"momentum_buffer"}


@pytest.mark.gpu
@pytest.mark.parametrize("name,max_norm,scale", CLIP)
def test_gpu_clip_grad_norm(gold, name, max_norm, scale):
    from improving_segmentation_with_selfsupervised_depth_b200 import optim
    params0, grads = optim_inputs()
    case = "clip_" + name
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in select(case, params0)]
    for p, g in zip(ps, select(case, grads[0])):
        p.grad = (g.clone() * scale).cuda()
    total = optim.clip_grad_norm_(ps, max_norm)
    assert total.is_cuda and rel(total, gold[case + "/total"]) < 2e-6
    for i, p in enumerate(ps):
        assert rel(p.grad, gold["%s/g%d" % (case, i)]) < 2e-6
    # a parameter without a gradient is skipped, like in torch
    ps.append(torch.nn.Parameter(torch.zeros(3, device="cuda")))
    optim.clip_grad_norm_(ps, max_norm)
