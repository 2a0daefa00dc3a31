"""Segmentation losses — drop-in for loss/loss.py on the sm_100a kernels."""
import math

import torch

from .. import ops


def cross_entropy2d(input, target, class_weight=None, pixel_weights=None):
    """Reference :17-37: ignore_index 250, mean over valid pixels; with pixel_weights the mean runs over
    all pixels.  NaN pixel weights disable the weighting (reference :31-32)."""
    if class_weight is not None:
        raise NotImplementedError("class_weight is unused by every reference config (train.py:503,649)")
    n, c, h, w = input.size()
    nt, ht, wt = target.size()
    if h != ht and w != wt:
        input = ops.bilinear(input, (ht, wt), align_corners=True)
    return ops.cross_entropy(input, target, pixel_weights=pixel_weights, ignore_index=250)


def berhu(input, target, mask, apply_log=False):
    raise NotImplementedError("berhu is a SURVEY §8(f) 'next' row (label-selection scoring), not built yet")


def pixel_wise_entropy(logits, normalize=False):
    raise NotImplementedError("pixel_wise_entropy is a SURVEY §8(f) 'next' row, not built yet")
