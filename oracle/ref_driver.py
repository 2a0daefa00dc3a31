"""Drives the vendored, UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__ and bench.py's reference arms (`--impl reference`, `--impl torch_gpu`) may import this.

What it does (SURVEY.md §8c recipe):
  * puts oracle/_ref first on sys.path and stubs the three packages the reference imports but this image lacks
    (matplotlib, kornia, ray) so that the reference's own `train.py` can be imported as-is;
  * replaces `train.build_loader` by a synthetic dataset (the reference's loaders need the Cityscapes files), after which
    the REAL `Trainer.__init__` / `Trainer.train_step` (reference train.py:157-343, 442-549) run unchanged;
  * with `dropin=True` the top-level names `models` / `loss` resolve to this repo's package
    (`install_dropin()`), everything else (`train`, `utils`, `loader`, `configs`, `evaluation`) stays the reference's:
    that is the drop-in boundary of DESIGN.md §1, exercised by tests/test_gpu_dropin_trainer.py.
"""
import copy
import importlib
import logging
import os
import sys
import types

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_TOP = ("models", "loss", "utils", "configs", "loader", "evaluation", "train", "experiments", "label_selection")


def available():
    return os.path.isfile(os.path.join(REF, "train.py")) and os.path.isdir(os.path.join(REF, "models"))


def _purge(names):
    for m in list(sys.modules):
        if m.split(".")[0] in names:
            del sys.modules[m]


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__segsde_stub__ = True
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, mod)
    return mod


def stub_missing_deps():
    """matplotlib / kornia / ray are imported at module level by train.py, loader/transformsgpu.py and
    utils/…; none of them is on the path measured here (plots, ColorJitter / GaussianBlur2d behind flags that the
    parity configs switch off, hyper-parameter search)."""
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("kornia")
    _stub("kornia.augmentation")
    _stub("kornia.filters")
    for n in ("ray", "ray.tune", "ray.tune.suggest", "ray.tune.suggest.variant_generator", "ray.tune.config_parser"):
        _stub(n)
    sys.modules["ray.tune"].__dict__.setdefault("TuneError", RuntimeError)
    sys.modules["ray.tune.suggest"].__dict__.setdefault("BasicVariantGenerator", object)
    for n in ("generate_variants", "flatten_resolved_vars", "format_vars"):
        sys.modules["ray.tune.suggest.variant_generator"].__dict__.setdefault(n, None)
    sys.modules["ray.tune.config_parser"].__dict__.setdefault("create_trial_from_spec", None)


def activate(dropin=False):
    """Make `import train`, `import models`, ... resolve to the vendored reference (and, with dropin=True, `models` /
    `loss` to this repo's package)."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    _purge(_TOP)
    if REF in sys.path:
        sys.path.remove(REF)
    sys.path.insert(0, REF)
    stub_missing_deps()
    if dropin:
        root = os.path.dirname(HERE)
        if root not in sys.path:
            sys.path.insert(1, root)
        import improving_segmentation_with_selfsupervised_depth_b200 as P
        P.install_dropin()
    _machine()


def _machine():
    """The reference's machine registry (configs/machine_config.py) with its relative directories pinned under
    oracle/_ref (machine-specific configuration, which the reference expects every user to provide)."""
    from configs.machine_config import MachineConfig
    MachineConfig("ws")
    MachineConfig.DOWNLOAD_MODEL_DIR = os.path.join(REF, "_model_dir") + os.sep
    MachineConfig.LOG_DIR = os.path.join(REF, "_logs") + os.sep
    return MachineConfig


def deactivate():
    _purge(_TOP)
    if REF in sys.path:
        sys.path.remove(REF)


# ----------------------------------------------------------------------------------------------------------------------
# configurations of BASELINE.json, restated from the reference's YAMLs + experiments.py
# ----------------------------------------------------------------------------------------------------------------------
def load_cfg(name, H, W, B, backbone="resnet50", n_workers=0, cudnn_benchmark=False):
    """name: dec5 | dec6 | joint | depthmix.  Returns the cfg dict `Trainer.__init__` expects (before its own merge of
    monodepth_options)."""
    yml = {"dec5": "cityscapes_monodepth_highres_dec5_crop.yml", "dec6": "cityscapes_monodepth_highres_dec6_crop.yml",
           "joint": "cityscapes_joint.yml", "depthmix": "cityscapes_joint.yml"}[name]
    with open(os.path.join(REF, "configs", yml)) as f:
        cfg = yaml.safe_load(f)
    cfg["seed"] = 1337
    m, t, mo = cfg["model"], cfg["training"], cfg["monodepth_options"]
    m.update(backbone_name=backbone, backbone_pretraining="none", depth_pretraining="none", pose_pretraining="none")
    mo.update(height=H, width=W, crop_h=H, crop_w=W)
    t.update(batch_size=B, val_batch_size=B, n_workers=n_workers, resume=None, n_tensorboard_trainimgs=0,
             benchmark=cudnn_benchmark)     # train.py:176 default is True (autotuned cuDNN algorithms); tests pin False
    t.setdefault("save_monodepth_ema", False)
    cfg["data"].setdefault("dataset_seed", 42)
    if name in ("dec5", "dec6"):
        m["depth_args"]["max_scale_size"] = [H, W]
        if name == "dec6":
            t["amp"] = False          # the YAML says True; fp32/TF32 here, the AMP-equivalent mode is reported separately
    else:
        # experiments.py:373-404 (exp 212): PAD multi-task decoder + dec-6 depth args, SGD with per-group learning rates
        m.update(segmentation_name="mtl_pad",
                 segmentation_args={"weights": "none", "output_stride": 1, "distillation_layer": 7, "side_output": True,
                                    "final_layer": 9},
                 depth_args={"intermediate_aspp": True, "aspp_rates": [6, 12, 18], "n_upconv": 4,
                             "num_ch_dec": [64, 128, 128, 256, 256], "max_scale_size": [H, W]},
                 disable_monodepth=False, disable_pose=False, freeze_backbone=False, freeze_depth=False,
                 freeze_pose=False, freeze_segmentation=False, enable_imnet_encoder=False)
        t.update(monodepth_lambda=1.0, segmentation_lambda=1.0, feat_dist_lambda=0.0, clip_grad_norm=10, amp=False,
                 optimizer={"name": "sgd", "lr": 1e-2, "backbone_lr": 1e-3, "pose_lr": 1e-6, "momentum": 0.9,
                            "weight_decay": 5e-4})
        t["unlabeled_segmentation"] = None
        cfg["data"].pop("generated_depth_dir", None)     # PAD predicts its own depth (no precomputed depth estimates)
        if name == "depthmix":
            t["unlabeled_segmentation"] = {
                "mix_mask": "depthcomp", "depthmix_online_depth": True, "depthcomp_margin": 0.03,
                "depthcomp_foreground_threshold": 0.0, "consistency_weight": 1, "only_unlabeled": False,
                "mix_use_gt": True, "backward_first_pseudo_label": False, "color_jitter": False, "blur": False}
    return cfg


def provide_imnet_weights(backbone="resnet50", seed=0):
    """There is no network: the ImageNet checkpoint that `enable_imnet_encoder` / `backbone_pretraining: imnet` load
    (torchvision hub cache for the reference, <DOWNLOAD_MODEL_DIR>/imagenet/<backbone>.pth for this repo) is replaced
    by ONE seeded random-weight torchvision state_dict written to both places, so both sides load identical tensors.
    TORCH_HOME is pointed into oracle/_ref so the real hub cache stays untouched."""
    import torchvision
    home = os.path.join(REF, "_torch_home")
    os.environ["TORCH_HOME"] = home
    torch.hub.set_dir(os.path.join(home, "hub"))
    wenum = {"resnet18": torchvision.models.ResNet18_Weights, "resnet50": torchvision.models.ResNet50_Weights,
             "resnet101": torchvision.models.ResNet101_Weights}[backbone].IMAGENET1K_V1
    fn = os.path.join(home, "hub", "checkpoints", os.path.basename(wenum.url))
    os.makedirs(os.path.dirname(fn), exist_ok=True)
    g = torch.Generator().manual_seed(seed)
    net = getattr(torchvision.models, backbone)(weights=None)
    sd = net.state_dict()
    for k, v in sd.items():
        if v.dtype.is_floating_point and "running" not in k:
            sd[k] = v + 0.01 * torch.randn(v.shape, generator=g)
    if not os.path.exists(fn):
        torch.save(sd, fn)
    own = os.path.join(_machine().DOWNLOAD_MODEL_DIR, "imagenet", "%s.pth" % backbone)
    os.makedirs(os.path.dirname(own), exist_ok=True)
    if not os.path.exists(own):
        torch.save(sd, own)
    return fn, own


class SyntheticLoader(torch.utils.data.Dataset):
    """Stands in for loader.build_loader(...): seeded synthetic samples of SURVEY.md §8(d) with the attributes the
    Trainer reads (`n_classes`, `ignore_index`)."""
    n_classes = 19
    ignore_index = 250

    def __init__(self, batch, n=64, labels=True, onehot=False):
        self.batch, self.n, self.labels, self.onehot = batch, n, labels, onehot

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        out = {}
        b = i % next(iter(self.batch.values())).shape[0]
        for k, v in self.batch.items():
            if k == "lbl" and not self.labels:
                continue
            out[k] = v[b]
        if self.onehot and "lbl" in self.batch:
            lbl = self.batch["lbl"][b]
            oh = torch.zeros(self.n_classes, *lbl.shape, device=lbl.device)
            valid = lbl != self.ignore_index
            oh.scatter_(0, lbl.clamp(0, self.n_classes - 1)[None], 1.0)
            oh *= valid[None]
            out["onehot_lbl"] = oh
            out["is_labeled"] = torch.tensor(i % 2 == 0)        # half of the samples keep the teacher's own softmax
        out["filename"] = "synthetic_%d" % i
        return out

    def decode_segmap_tocolor(self, x):
        return x


def make_trainer(cfg, batch, dropin, device=None):
    """The reference's real Trainer over a synthetic dataset.  `batch`: dict of CPU tensors with leading dim B (the
    samples every loader serves).  device: None = the reference's own choice (cuda if available)."""
    activate(dropin=dropin)
    train = importlib.import_module("train")
    cfg = copy.deepcopy(cfg)
    if cfg["model"].get("enable_imnet_encoder") or cfg["model"].get("backbone_pretraining") == "imnet":
        provide_imnet_weights(cfg["model"]["backbone_name"])

    def build_loader(data_cfg, split="train", load_labels=True, load_sequence=True):
        return SyntheticLoader(batch, labels=load_labels, onehot=bool(data_cfg.get("load_onehot", False)))
    train.build_loader = build_loader
    logger = logging.getLogger("segsde")
    tr = train.Trainer(cfg, None, None, logger, "synthetic")
    if device is not None and torch.device(device) != tr.device:
        tr.device = torch.device(device)
        tr.model.to(tr.device)
        if tr.ema_model is not None:
            tr.ema_model.to(tr.device)
    return tr
