"""GPU: the drop-in boundary, proven on the reference's OWN training step.

The unmodified reference (`oracle/_ref`, vendored by oracle/make_ref.py) provides `train.Trainer`; its
`train_step` (train.py:442-549, incl. `train_step_segmentation_unlabeled` :653-760) is run twice on the same synthetic
batches: once with the reference's `models` / `loss` (PyTorch / cuDNN on this GPU), once with this repo's package
registered under those names by `install_dropin()` — same initial state_dict, CPU RNG re-seeded before every step (the
auto-mask noise is a CPU `torch.randn`, monodepth_loss.py:163-164), dropout masks of the reference run replayed.

 * `test_loss_curve_100_steps`: BASELINE config 1 family (dec5, ResNet-50 frozen, Adam 1e-4), 100 optimizer steps:
   the two loss curves must agree to 1e-3 relative at every step (north_star) on the fp32 CUDA-core route (cuDNN TF32
   off), and on the tcgen05 TF32 route against the reference with cuDNN TF32 on (its default).
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


def _run(R, cfg, batch, steps, dropin, init_sd=None, masks=None, tf32=False):
    """Returns (per-step loss dicts, initial state_dict, per-step dropout masks by module name)."""
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
            inputs = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
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
    lines = []
    worst = 0.0
    for i, (a, b) in enumerate(zip(ref, own)):
        for k in keys:
            e = abs(a[k] - b[k]) / (abs(a[k]) + 1e-12) if a[k] != 0 else abs(b[k])
            worst = max(worst, e)
        if i % 10 == 0 or i == len(ref) - 1:
            lines.append("%s step %3d ref %.6f own %.6f" % (tag, i, a[keys[0]], b[keys[0]]))
    print("\n".join(lines))
    print("%s worst relative difference over %d steps: %.3e" % (tag, len(ref), worst))
    return worst


@pytest.mark.parametrize("route,tol", [("fp32", 1e-3), ("tf32", 1e-3)])
def test_loss_curve_100_steps(route, tol):
    R = _driver()
    B, H, W = 2, 64, 128
    cfg = R.load_cfg("dec5", H, W, B, "resnet50")
    batch = _batch(B, H, W, seed=21)
    tf32 = route == "tf32"
    ref, sd0, masks = _run(R, cfg, batch, 100, dropin=False, tf32=tf32)
    own, _, _ = _run(R, cfg, batch, 100, dropin=True, init_sd=sd0, masks=masks, tf32=tf32)
    worst = _curves("dec5/" + route, ref, own, ["mono_loss", "total_loss"])
    assert ref[-1]["mono_loss"] < ref[0]["mono_loss"]          # it trains
    assert worst < tol, worst


@pytest.mark.parametrize("name,keys", [
    ("dec6", ["mono_loss", "feat_dist_loss", "total_loss"]),
    ("joint", ["segmentation_loss", "mono_loss", "total_loss"]),
    ("depthmix", ["segmentation_loss", "segmentation_total_loss", "mono_total_loss", "total_loss"])])
def test_config_steps(name, keys):
    R = _driver()
    B, H, W = 2, 64, 128
    cfg = R.load_cfg(name, H, W, B, "resnet50")
    batch = _batch(B, H, W, seed=22)
    ref, sd0, masks = _run(R, cfg, batch, 3, dropin=False)
    own, _, _ = _run(R, cfg, batch, 3, dropin=True, init_sd=sd0, masks=masks)
    worst = _curves(name, ref, own, keys)
    assert worst < 2e-3, worst
