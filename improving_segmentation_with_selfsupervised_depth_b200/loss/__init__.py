"""Factories of the reference's loss package (loss/__init__.py:1-37)."""
import functools
import logging

from .loss import cross_entropy2d
from .monodepth_loss import MonodepthLoss

logger = logging.getLogger("segsde")

# segmentation loss registry (the name `key2loss` is part of the reference's module surface)
key2loss = {"cross_entropy": cross_entropy2d}


def get_segmentation_loss_function(cfg):
    """`training.segmentation_loss` is either None (plain cross entropy) or {"name": <key2loss key>, **kwargs bound to
    the loss}; an unknown name raises NotImplementedError (reference :16-30)."""
    spec = cfg["training"]["segmentation_loss"]
    if spec is None:
        logger.info("Using default cross entropy loss")
        return cross_entropy2d
    bound = dict(spec)
    name = bound.pop("name")
    fn = key2loss.get(name)
    if fn is None:
        raise NotImplementedError("Loss {} not implemented".format(name))
    logger.info("Using {} with {} params".format(name, bound))
    return functools.partial(fn, **bound)


def get_monodepth_loss(cfg, is_train, batch_size=None):
    """MonodepthLoss from `training.monodepth_loss` (reference :32-37); the batch size defaults to the training one."""
    options = dict(cfg["training"]["monodepth_loss"])
    options["batch_size"] = cfg["training"]["batch_size"] if batch_size is None else batch_size
    return MonodepthLoss(is_train=is_train, **options)
