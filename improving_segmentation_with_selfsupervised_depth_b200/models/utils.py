"""Factories / checkpoint plumbing — drop-in for models/utils.py."""
import os
import re

import torch
from torch import nn

from .depth_decoder import DepthDecoder
from .pose_decoder import PoseDecoder
from .resnet_encoder import ResnetEncoder


def _download_dir():
    try:   # when used inside the reference tree its machine registry decides (models/utils.py:11)
        from configs.machine_config import MachineConfig
        return MachineConfig.DOWNLOAD_MODEL_DIR
    except Exception:
        return os.environ.get("SEGSDE_MODEL_DIR", os.path.expanduser("~/.cache/segsde_models"))


def _device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def get_resnet_backbone(backbone_name, backbone_pretraining="none", replace_stride_with_dilation=None,
                        use_intermediate_layer_getter=False, num_input_images=1):
    """Reference :18-55."""
    if backbone_name not in ["resnet18", "resnet50", "resnet101"]:
        raise NotImplementedError
    n_res = int(re.match(r"([a-z]+)([0-9]+)", backbone_name, re.I).groups()[-1])
    kw = {} if num_input_images > 1 else {"replace_stride_with_dilation": replace_stride_with_dilation}
    if backbone_pretraining in ("none", "imnet") or "mono" in backbone_pretraining:
        backbone = ResnetEncoder(n_res, False, num_input_images=num_input_images, **kw)
    else:
        raise NotImplementedError
    if backbone_pretraining == "imnet":
        path = os.path.join(_download_dir(), "imagenet", "%s.pth" % backbone_name)
        if os.path.exists(path):
            sd = torch.load(path, map_location="cpu")
            if num_input_images > 1:   # reference resnet_encoder.py:57-59
                sd["conv1.weight"] = torch.cat([sd["conv1.weight"]] * num_input_images, 1) / num_input_images
            # strict except for the classifier, which the encoder drops (avgpool / fc become Identity below)
            own = backbone.encoder.state_dict()
            missing = [k for k in own if k not in sd and not k.startswith("fc.")]
            unexpected = [k for k in sd if k not in own and not k.startswith("fc.")]
            if missing or unexpected:
                raise RuntimeError("ImageNet checkpoint %s does not match %s: missing %s, unexpected %s"
                                   % (path, backbone_name, missing[:4], unexpected[:4]))
            backbone.encoder.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        elif os.environ.get("SEGSDE_ALLOW_RANDOM_IMNET", "0") == "1":
            print("WARNING: ImageNet weights for %s not found at %s — SEGSDE_ALLOW_RANDOM_IMNET=1: random init "
                  "(the feature-distance target is then a random network)" % (backbone_name, path))
        else:
            # the reference downloads torchvision's ImageNet weights here (resnet_encoder.py:52-60); there is no network,
            # and silently training against a random `imnet_encoder` would not reproduce the reference
            raise FileNotFoundError("ImageNet weights for %s not found at %s; place torchvision's %s state_dict there "
                                    "or set SEGSDE_ALLOW_RANDOM_IMNET=1 to continue with random init"
                                    % (backbone_name, path, backbone_name))
    elif "mono" in backbone_pretraining:
        print('Load ' + backbone_pretraining + 'weights')
        download_model_if_doesnt_exist(backbone_pretraining)
        path = os.path.join(_download_dir(), backbone_pretraining, "encoder.pth")
        loaded = torch.load(path, map_location=_device())
        own = backbone.state_dict()
        backbone.load_state_dict({k: v for k, v in loaded.items() if k in own}, strict=False)
    backbone.encoder.avgpool = nn.Identity()
    backbone.encoder.fc = nn.Identity()
    if use_intermediate_layer_getter:
        raise NotImplementedError("IntermediateLayerGetter backbones are not used by any shipped config")
    return backbone


def get_depth_decoder(depth_pretraining, num_ch_enc, scales=range(4), **kwargs):
    """Reference :58-73."""
    decoder = DepthDecoder(num_ch_enc, scales, **kwargs)
    decoder.to(_device())
    if depth_pretraining not in (None, 'none'):
        print('Load ' + depth_pretraining + 'depth weights')
        download_model_if_doesnt_exist(depth_pretraining)
        path = os.path.join(_download_dir(), depth_pretraining, "depth.pth")
        decoder.load_state_dict(torch.load(path, map_location=_device()))
    return decoder


def get_posenet(backbone_name, backbone_pretraining, pose_pretraining, num_pose_frames):
    """Reference :76-97."""
    models = {"pose_encoder": get_resnet_backbone(
        backbone_name=backbone_name, backbone_pretraining="imnet" if backbone_pretraining == "imnet" else "none",
        num_input_images=num_pose_frames)}
    models["pose"] = PoseDecoder(models["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
    if "mono" in pose_pretraining:
        for mn in ["pose_encoder", "pose"]:
            download_model_if_doesnt_exist(pose_pretraining)
            path = os.path.join(_download_dir(), pose_pretraining, "{}.pth".format(mn))
            loaded = torch.load(path, map_location=_device())
            own = models[mn].state_dict()
            models[mn].load_state_dict({k: v for k, v in loaded.items() if k in own})
    return models


def _get_layer(encoder, decoder, layer):
    """Reference :100-105."""
    return encoder[layer] if layer <= 4 else decoder[("upconv", 9 - layer)]


def download_model_if_doesnt_exist(model_name, download_dir=None):
    """Reference :108-171 downloads from Google Drive; there is no network here, so the checkpoint
    directory must already exist."""
    download_dir = os.path.expandvars(download_dir or _download_dir()).replace('$SLURM_JOB_ID/', '')
    model_path = os.path.join(download_dir, model_name)
    if not os.path.exists(os.path.join(model_path, "depth.pth")):
        raise FileNotFoundError("pretrained model %s not found under %s (downloading is out of scope: no network)"
                                % (model_name, download_dir))
