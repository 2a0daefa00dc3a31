"""B200-native (sm_100a) implementation of the monodepth / joint-segmentation hot path of
lhoyer/improving_segmentation_with_selfsupervised_depth, behind the reference's `models/` and
`loss/` Python API.  `install_dropin()` registers the sub-packages under the reference's
top-level names so that the reference's `train.py` imports them unchanged."""
import importlib
import sys

__version__ = "0.1.0"


def install_dropin():
    """Makes `import models` / `import loss` (train.py:26-29) resolve to this package."""
    pkg = __name__
    for name in ("models", "loss"):
        mod = importlib.import_module(pkg + "." + name)
        sys.modules[name] = mod
        prefix = pkg + "." + name + "."
        for k, v in list(sys.modules.items()):
            if k.startswith(prefix):
                sys.modules[name + "." + k[len(prefix):]] = v
    return sys.modules["models"], sys.modules["loss"]
