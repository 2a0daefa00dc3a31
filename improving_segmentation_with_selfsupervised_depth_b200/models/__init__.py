"""`models` package surface of the reference (models/__init__.py:1-23): `get_model(model_dict, n_classes)`."""
import copy

from .joint_segmentation_depth import joint_segmentation_depth

# architecture name (`model.arch` in the YAML configs) -> factory(name=, num_classes=, **remaining config keys)
ARCHITECTURES = {"joint_segmentation_depth": joint_segmentation_depth}


def _get_model_instance(name):
    """Factory for an architecture name; unknown names raise NotImplementedError like the reference (:17-23)."""
    factory = ARCHITECTURES.get(name)
    if factory is None:
        raise NotImplementedError("Model {} not available".format(name))
    return factory


def get_model(model_dict, n_classes):
    """Builds the network described by `model_dict`: `arch` picks the factory, every other key becomes a factory
    keyword (deep-copied, the caller's config is never mutated)."""
    arch = model_dict["arch"]
    kwargs = {k: copy.deepcopy(v) for k, v in model_dict.items() if k != "arch"}
    return _get_model_instance(arch)(name=arch, num_classes=n_classes, **kwargs)
