"""Step-level ops of train.py: the oracle against outputs of the unmodified reference's own
functions (tests/golden/train_ops_golden.npz, CPU) and the CUDA kernels against both (GPU).  Integer outputs (mix
mask, pseudo labels) must match exactly; fp32 results to 1e-6 (elementwise) / 3e-6 (reductions) relative."""
import os
import sys

import numpy as np
import pytest
import torch

import segsde_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_train_ops import CONSISTENCY, EMA_CASES, FT, MARGIN, train_ops_inputs      # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "train_ops_golden.npz"), allow_pickle=False)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ------------------------------------------------------------------------------------------ CPU: oracle vs reference
def test_oracle_feature_distance(gold):
    x = train_ops_inputs()
    a = x["feats"][0].clone().requires_grad_()
    d = O.feature_distance(a, x["feats"][1])
    d.backward()
    assert rel(d.detach(), gold["t1/dist"]) < 1e-6 and rel(a.grad, gold["t1/grad"]) < 1e-6


def test_oracle_depthmix(gold):
    x = train_ops_inputs()
    dn = O.normalize_depths(x["depths"])
    assert rel(dn, gold["t2/depths_norm"]) < 1e-6
    mask = O.depthcomp_mix_mask(dn, MARGIN, FT)
    assert torch.equal(mask, torch.as_tensor(gold["t2/mask"]))
    assert rel(O.mix(mask, x["imgs"]), gold["t2/mix_img"]) < 1e-6
    assert rel(O.mix(mask, x["softmax_t"]), gold["t2/mix_softmax"]) < 1e-6


def test_oracle_pseudo_label_loss(gold):
    x = train_ops_inputs()
    s = x["logits_s"].clone().requires_grad_()
    loss, label = O.calc_pseudo_label_loss(x["softmax_t"], s, CONSISTENCY)
    loss.backward()
    assert torch.equal(label, torch.as_tensor(gold["t3/label"]))
    assert rel(loss.detach(), gold["t3/loss"]) < 1e-6 and rel(s.grad, gold["t3/grad"]) < 1e-6


def test_oracle_ema(gold):
    x = train_ops_inputs()
    for it, alpha in EMA_CASES:
        ema = [t.clone() for t in x["ema"]]
        O.update_ema(ema, x["params"], alpha, it)
        for k, t in enumerate(ema):
            assert rel(t, gold["t4/it%d/%d" % (it, k)]) < 1e-6


# ------------------------------------------------------------------------------------------ GPU: kernels vs reference
def _ops():
    from improving_segmentation_with_selfsupervised_depth_b200 import train_ops
    return train_ops


@pytest.mark.gpu
def test_gpu_feature_distance(gold):
    T = _ops()
    x = train_ops_inputs()
    a = x["feats"][0].cuda().requires_grad_()
    b = x["feats"][1].cuda().requires_grad_()
    d = T.feature_distance(a, b)
    (2.0 * d).backward()
    assert rel(d.detach(), gold["t1/dist"]) < 3e-6
    assert rel(a.grad / 2.0, gold["t1/grad"]) < 3e-6 and rel(b.grad / -2.0, gold["t1/grad"]) < 3e-6
    # channels-last feature maps (what the encoder produces) and a zero distance
    ac = x["feats"][0].cuda().contiguous(memory_format=torch.channels_last)
    bc = x["feats"][1].cuda().contiguous(memory_format=torch.channels_last)
    assert rel(T.feature_distance(ac, bc), gold["t1/dist"]) < 3e-6
    z = torch.ones(4, 8, device="cuda", requires_grad=True)
    dz = T.feature_distance(z, torch.ones(4, 8, device="cuda"))
    dz.backward()
    assert float(dz.detach()) == 0.0 and float(z.grad.abs().max()) == 0.0


@pytest.mark.gpu
def test_gpu_depthmix(gold):
    T = _ops()
    x = train_ops_inputs()
    dn = T.normalize_depths(x["depths"].cuda())
    assert rel(dn, gold["t2/depths_norm"]) < 1e-6
    mask = T.depthcomp_mix_mask(dn, MARGIN, FT)
    assert mask.dtype == torch.int64 and torch.equal(mask.cpu(), torch.as_tensor(gold["t2/mask"]))
    img, _ = T.mix(mask, data=x["imgs"].cuda())
    assert rel(img, gold["t2/mix_img"]) < 1e-6
    sm = x["softmax_t"].cuda()
    for t in (sm, sm.contiguous(memory_format=torch.channels_last)):       # planar and channels-last teacher softmax
        out, _ = T.mix(mask, data=t)
        assert out.stride() == t.stride() and rel(out, gold["t2/mix_softmax"]) < 1e-6
    out, tgt = T.mix(mask.float(), data=sm, target=x["imgs"].cuda())          # fp32 mask, data + target
    assert rel(out, gold["t2/mix_softmax"]) < 1e-6 and rel(tgt, gold["t2/mix_img"]) < 1e-6
    # B = 4: pairing (i, (i+1) % B) against the oracle
    g = torch.Generator().manual_seed(3)
    d4 = torch.rand(4, 1, 16, 24, generator=g)
    m4 = T.depthcomp_mix_mask(d4.cuda(), 0.05, 0.1)
    assert torch.equal(m4.cpu(), O.depthcomp_mix_mask(d4, 0.05, 0.1))
    x4 = torch.rand(4, 5, 16, 24, generator=g)
    assert rel(T.mix(m4, data=x4.cuda())[0], O.mix(m4.cpu(), x4)) < 1e-6


@pytest.mark.gpu
def test_gpu_pseudo_label_loss(gold):
    T = _ops()
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    P.install_dropin()
    x = train_ops_inputs()
    for fmt in (torch.contiguous_format, torch.channels_last):
        s = x["logits_s"].cuda().requires_grad_()
        loss, label = T.calc_pseudo_label_loss(x["softmax_t"].cuda().contiguous(memory_format=fmt), s, CONSISTENCY)
        loss.backward()
        assert torch.equal(label.cpu(), torch.as_tensor(gold["t3/label"]))
        assert rel(loss.detach(), gold["t3/loss"]) < 3e-6
        assert rel(s.grad, gold["t3/grad"]) < 3e-6
    label, weight = T.pseudo_labels(x["softmax_t"].cuda())
    share = float((x["softmax_t"].max(1)[0] >= 0.968).sum()) / label.numel()
    assert abs(float(weight.min()) - share) < 1e-7 and abs(float(weight.max()) - share) < 1e-7


@pytest.mark.gpu
def test_gpu_ema_update(gold):
    T = _ops()
    x = train_ops_inputs()
    for it, alpha in EMA_CASES:
        ema = [torch.nn.Parameter(t.clone().cuda()) for t in x["ema"]]
        # a channels-last 4-D parameter, as the drop-in's conv weights are stored
        ema[0] = torch.nn.Parameter(x["ema"][0].clone().cuda().contiguous(memory_format=torch.channels_last))
        params = [torch.nn.Parameter(t.clone().cuda()) for t in x["params"]]
        params[0] = torch.nn.Parameter(x["params"][0].clone().cuda().contiguous(memory_format=torch.channels_last))
        a = T.update_ema_variables(ema, params, alpha, it)
        assert a == min(1 - 1 / (it + 1), alpha)
        for k, t in enumerate(ema):
            assert rel(t.detach(), gold["t4/it%d/%d" % (it, k)]) < 1e-6
    # many tensors: more than one table of 48 tensors / 320 chunks
    g = torch.Generator().manual_seed(9)
    dst = [torch.randn(int(n), generator=g) for n in torch.randint(1, 3000, (130,), generator=g)] + [torch.randn(700000, generator=g)]
    src = [torch.randn(t.shape, generator=g) for t in dst]
    want = [0.9 * d + 0.1 * s for d, s in zip(dst, src)]
    dg = [d.cuda() for d in dst]
    T.multi_axpby(dg, [s.cuda() for s in src], 0.9, 0.1)
    assert max(rel(a, b) for a, b in zip(dg, want)) < 1e-6
