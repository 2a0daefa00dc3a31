"""Shared test helpers (CPU oracle side + comparison utilities)."""
import numpy as np
import torch

import segsde_oracle as O

LOSS_KW = dict(min_depth=0.1, max_depth=100, test_min_depth=1e-3, test_max_depth=80, disparity_smoothness=1e-3,
               no_ssim=False, avg_reprojection=False, disable_automasking=False)


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def grad_close(a, b, tol, frac=3e-2, med=1e-5):
    """Gradient comparison for the photometric loss.  d loss / d disparity is discontinuous in fp32-rounding-
    sized events: (i) bilinear grid_sample has a slope jump wherever a sample coordinate crosses an integer
    (|ix| ~ 500 carries ~1e-4 px of rounding, so ~2e-4 of all samples land on the other side of a texel
    boundary than in the reference's evaluation order), (ii) border clipping masks (ix <= 0, ix >= W-1) and
    (iii) the per-pixel arg-min between nearly tied candidates.  Each event moves one pixel's gradient by O(1);
    a scale-s disparity element aggregates 4^s pixels x 2 frames, so up to ~3% of the coarsest map's elements
    can contain one.  Everything else must agree to `tol` (relative to the max) and the median error must be
    at rounding level."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    scale = b.abs().max() + 1e-30
    err = (a - b).abs() / scale
    return (err > tol).double().mean().item() <= frac and err.median().item() <= med


def loss_case(golden):
    """The loss-level case of tests/golden/make_golden.py: inputs, disparities, poses."""
    B, H, W = 2, 64, 96
    inputs = O.synthetic_inputs(B, H, W, seed=7)
    disps = [torch.from_numpy(golden["loss_disp%d" % s]).clone() for s in range(4)]
    Ts = {-1: torch.from_numpy(golden["loss_T-1"]).clone(), 1: torch.from_numpy(golden["loss_T1"]).clone()}
    return B, H, W, inputs, disps, Ts


def loss_noise(B, H, W, avg):
    torch.manual_seed(11)
    return [torch.randn(B, 1 if avg else 2, H, W) * 0.00001 for _ in range(4)]


def model_cfg_from_contract(c):
    return dict(c["cfg"])


def unpack_mask(golden, prefix):
    shape = tuple(int(v) for v in golden[prefix + "dropout_shape"])
    bits = np.unpackbits(golden[prefix + "dropout_mask"])[:int(np.prod(shape))]
    return torch.from_numpy(bits.reshape(shape).astype(np.float32))


def unpack_named_mask(golden, key):
    shape = tuple(int(v) for v in golden[key + "_shape"])
    bits = np.unpackbits(golden[key])[:int(np.prod(shape))]
    return torch.from_numpy(bits.reshape(shape).astype(np.float32))


def noise_floor_retry(fn):
    """For tests that hold this package to a multiple of the cuDNN-TF32 oracle's own distance from the fp32 oracle: that
    floor is itself a random variable (cuDNN picks algorithms by the workspace it can get, split-K atomics reorder), and a
    1-in-6 unlucky draw was observed in the full suite while the test passed 3/3 in isolation.  One re-draw on failure;
    a real regression fails both."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        try:
            return fn(*a, **k)
        except AssertionError as first:
            print("noise-floor assertion failed once (%s); re-drawing" % (str(first)[:200],))
            import torch
            torch.cuda.empty_cache()
            return fn(*a, **k)
    return wrapped

