"""Factories of the reference's loss package (loss/__init__.py:1-37)."""
import functools
import logging

from .loss import cross_entropy2d
from .monodepth_loss import MonodepthLoss

logger = logging.getLogger("segsde")

key2loss = {"cross_entropy": cross_entropy2d}


def get_segmentation_loss_function(cfg):
    spec = cfg["training"]["segmentation_loss"]
    if spec is None:
        logger.info("Using default cross entropy loss")
        return cross_entropy2d
    name = spec["name"]
    params = {k: v for k, v in spec.items() if k != "name"}
    if name not in key2loss:
        raise NotImplementedError("Loss {} not implemented".format(name))
    logger.info("Using {} with {} params".format(name, params))
    return functools.partial(key2loss[name], **params)


def get_monodepth_loss(cfg, is_train, batch_size=None):
    if batch_size is None:
        batch_size = cfg["training"]["batch_size"]
    return MonodepthLoss(**cfg["training"]["monodepth_loss"], batch_size=batch_size, is_train=is_train)
