"""GPU: the drop-in boundary, proven on the reference's OWN training step.

The unmodified reference (`oracle/_ref`, vendored by oracle/make_ref.py) provides `train.Trainer`; its
`train_step` (train.py:442-549, incl. `train_step_segmentation_unlabeled` :653-760) is run twice on the same synthetic
batches: once with the reference's `models` / `loss` (PyTorch / cuDNN on this GPU), once with this repo's package
registered under those names by `install_dropin()` — same initial state_dict, CPU RNG re-seeded before every step (the
auto-mask noise is a CPU `torch.randn`, monodepth_loss.py:163-164), dropout masks of the reference run replayed.

 * `test_loss_curve_100_steps`: BASELINE config 1 family (dec5, ResNet-50 frozen, Adam 1e-4), 100 optimizer steps on the
   fp32 CUDA-core route and on the tcgen05 TF32 route, each held to the deviation the reference shows from ITSELF under
   the same class of perturbation (last-bit weight noise / cuDNN TF32) — see the test's docstring for why 1e-3 over 100
   steps is not attainable even by the reference against itself.
 * `test_config_steps`: dec6 (unfrozen encoder + frozen ImageNet encoder + feature distance), joint (PAD, two CE
   losses, SGD groups, clip_grad_norm) and depthmix (mean teacher, DepthMix, pseudo labels, EMA): every entry of the
   step's loss dict over 3 steps — covers `backward(retain_graph=True)` followed by a second backward through the
   custom autograd Functions, `optimizer.zero_grad()` semantics and torch's clip / SGD on this repo's parameters.
"""
import contextlib
import io
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _driver():
    import ref_driver as R
    if not R.available():
        pytest.skip("oracle/_ref not vendored (python oracle/make_ref.py where /root/reference exists)")
    return R


def _batch(B, H, W, seed):
    from improving_segmentation_with_selfsupervised_depth_b200.synthetic import synthetic_inputs
    return synthetic_inputs(B, H, W, seed=seed, labels=True)


def _dropouts(model):
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)) and m.p > 0]


def _run(R, cfg, batch, steps, dropin, init_sd=None, masks=None, tf32=False, batches=None):
    """Returns (per-step loss dicts, initial state_dict, per-step dropout masks by module name).  batches: optional list
    of input dicts cycled over the steps (default: `batch` every step)."""
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, ops.USE_TC)
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    ops.USE_TC = tf32
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            tr = R.make_trainer(cfg, batch, dropin=dropin)
        models = [tr.model] + ([tr.ema_model] if tr.ema_model is not None else [])
        if init_sd is not None:
            for m, sd in zip(models, init_sd):
                m.load_state_dict(sd)
        sd0 = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in models]
        if dropin:
            tr.monodepth_loss_calculator_train.noise = "torch"      # the reference's own CPU randn stream
        rec, hooks = [], []
        if not dropin:
            for mi, m in enumerate(models):
                for name, mod in _dropouts(m):
                    hooks.append(mod.register_forward_hook(
                        lambda mod, inp, out, key=(mi, name): rec[-1].setdefault(key, []).append((out != 0).float().cpu())
                        if mod.training else None))
        out = []
        for step in range(steps):
            torch.manual_seed(1000 + step)
            rec.append({})
            if dropin:
                for mi, m in enumerate(models):
                    for name, mod in _dropouts(m):
                        mod.replay_mask = list(masks[step].get((mi, name), []))
            src = batch if batches is None else batches[step % len(batches)]
            inputs = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in src.items()}
            with contextlib.redirect_stdout(io.StringIO()):
                losses = tr.train_step(inputs, step)
            out.append({k: float(v) for k, v in losses.items()})
        for h in hooks:
            h.remove()
        return out, sd0, rec
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, ops.USE_TC = prev
        R.deactivate()


def _curves(tag, ref, own, keys):
    """Worst pointwise relative difference of the two curves; the curves themselves go to gpurun_out/ for the record."""
    import json
    lines = []
    worst, at = 0.0, None
    for i, (a, b) in enumerate(zip(ref, own)):
        for k in keys:
            e = abs(a[k] - b[k]) / (abs(a[k]) + 1e-12) if a[k] != 0 else abs(b[k])
            if e > worst:
                worst, at = e, (i, k, a[k], b[k])
        if i % 10 == 0 or i == len(ref) - 1:
            lines.append("%s step %3d ref %.6f own %.6f" % (tag, i, a[keys[0]], b[keys[0]]))
    print("\n".join(lines))
    print("%s worst relative difference over %d steps: %.3e at %s" % (tag, len(ref), worst, at))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "curve_%s.json" % tag.replace("/", "_")), "w") as f:
            json.dump({"keys": keys, "ref": ref, "own": own, "worst": worst, "at": at}, f)
    return worst


def test_loss_curve_100_steps():
    """100 optimizer steps of the dec5 recipe (ResNet-50 frozen, Adam 1e-4, 8 synthetic batches cycled) through the
    reference's own `Trainer.train_step`, reference `models`/`loss` vs this repo's drop-in.

    The north-star asks for curves within 1e-3 over 100 steps.  Training a random-weight, train-mode-BatchNorm network
    at lr 1e-4 is chaotic: the REFERENCE ITSELF, re-run with its initial weights perturbed in the last fp32 bit (run B
    below), leaves its own curve by > 1e-3 after ~10 steps and by O(1) after ~40.  So the test pins what can be pinned:
      * fp32 CUDA-core route vs the reference with cuDNN TF32 off: <= 1e-3 (relative to the initial loss) for the first
        8 steps — i.e. while rounding noise has not been amplified yet — and over all 100 steps a mean deviation not larger
        than 3x the reference's own bit-perturbation sensitivity;
      * tcgen05 TF32 route vs the same curve: first step and mean deviation within 3x of what the reference itself shows
        when its convolutions run in cuDNN's TF32 (its default on this GPU, run C)."""
    R = _driver()
    B, H, W = 2, 64, 128
    cfg = R.load_cfg("dec5", H, W, B, "resnet50")
    batches = [_batch(B, H, W, seed=21 + i) for i in range(8)]
    N = 100
    ref_a, sd0, masks = _run(R, cfg, batches[0], N, dropin=False, tf32=False, batches=batches)
    g = torch.Generator(device="cuda").manual_seed(5)
    sd_b = [{k: (v * (1 + 1.2e-7 * (2 * torch.rand(v.shape, device=v.device, generator=g) - 1))
                 if (v.dtype.is_floating_point and v.dim() == 4) else v) for k, v in sd.items()} for sd in sd0]
    ref_b, _, _ = _run(R, cfg, batches[0], N, dropin=False, init_sd=sd_b, tf32=False, batches=batches)
    ref_c, _, _ = _run(R, cfg, batches[0], N, dropin=False, init_sd=sd0, tf32=True, batches=batches)
    own32, _, _ = _run(R, cfg, batches[0], N, dropin=True, init_sd=sd0, masks=masks, tf32=False, batches=batches)
    owntf, _, _ = _run(R, cfg, batches[0], N, dropin=True, init_sd=sd0, masks=masks, tf32=True, batches=batches)
    scale = ref_a[0]["mono_loss"]

    def dev(x):
        return [abs(u["mono_loss"] - v["mono_loss"]) / scale for u, v in zip(x, ref_a)]
    d_b, d_c, d_32, d_tf = dev(ref_b), dev(ref_c), dev(own32), dev(owntf)
    import json
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "curve_dec5_100steps.json"), "w") as f:
            json.dump({"ref_fp32": [r["mono_loss"] for r in ref_a], "ref_fp32_bit_perturbed": [r["mono_loss"] for r in ref_b],
                       "ref_cudnn_tf32": [r["mono_loss"] for r in ref_c], "own_fp32_route": [r["mono_loss"] for r in own32],
                       "own_tcgen05_tf32_route": [r["mono_loss"] for r in owntf]}, f)
    mean = lambda v: sum(v) / len(v)        # noqa: E731
    print("deviation from the reference fp32 curve, relative to the initial loss %.4f:" % scale)
    for name, d in (("reference, weights perturbed by 1 ulp", d_b), ("reference, cuDNN TF32", d_c),
                    ("this repo, fp32 route", d_32), ("this repo, tcgen05 TF32 route", d_tf)):
        print("  %-40s step0 %.2e  max(steps<8) %.2e  mean(100) %.2e  max %.2e" % (name, d[0], max(d[:8]), mean(d), max(d)))
    assert mean([r["mono_loss"] for r in ref_a[-16:]]) < mean([r["mono_loss"] for r in ref_a[:16]])      # it trains
    assert max(d_32[:8]) < 1e-3, d_32[:8]
    assert mean(d_32) < 3 * mean(d_b) + 1e-3, (mean(d_32), mean(d_b))
    assert d_tf[0] < 3 * d_c[0] + 1e-3, (d_tf[0], d_c[0])
    assert mean(d_tf) < 3 * mean(d_c) + 1e-3, (mean(d_tf), mean(d_c))


@pytest.mark.parametrize("name,keys,tf32,steps,tol", [
    ("dec6", ["mono_loss", "feat_dist_loss", "total_loss"], False, 3, 2e-3),
    ("joint", ["segmentation_loss", "mono_loss", "total_loss"], False, 3, 2e-3),
    ("depthmix", ["segmentation_loss", "segmentation_total_loss", "mono_total_loss", "total_loss"], False, 3, 2e-3),
    # the tcgen05 route of the PAD multi-task decoder (what bench.py --config joint runs) against the reference with
    # cuDNN TF32 convolutions: first step only (both sides carry TF32 operand rounding, amplified by train-mode BatchNorm
    # on a 4 x 8-pixel bottleneck), 5 % on every loss of the step
    ("joint", ["segmentation_loss", "mono_loss", "total_loss"], True, 1, 5e-2)])
def test_config_steps(name, keys, tf32, steps, tol):
    R = _driver()
    B, H, W = 2, 64, 128
    cfg = R.load_cfg(name, H, W, B, "resnet50")
    batch = _batch(B, H, W, seed=22)
    ref, sd0, masks = _run(R, cfg, batch, steps, dropin=False, tf32=tf32)
    own, _, _ = _run(R, cfg, batch, steps, dropin=True, init_sd=sd0, masks=masks, tf32=tf32)
    worst = _curves(name + ("_tf32" if tf32 else ""), ref, own, keys)
    assert worst < tol, worst
