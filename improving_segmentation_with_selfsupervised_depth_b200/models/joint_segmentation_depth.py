"""JointSegmentationMonodepth and its factory — drop-in for models/joint_segmentation_depth.py.

The network is a dictionary of sub-modules (`.models`, keys a subset of encoder / imnet_encoder / pose_encoder / pose /
depth / segmentation / mtl_decoder — train.py indexes them by name) plus the routing between them; the factory turns
the YAML `model` section into that dictionary.
"""
import torch
from torch import nn

from .joint_segmentation_depth_decoder import PAD, JointSegDepthDecoder
from .monodepth_layers import transformation_from_parameters
from .utils import get_depth_decoder, get_posenet, get_resnet_backbone

SEGMENTATION_DECODERS = {"joint_seg_depth_dec": JointSegDepthDecoder, "mtl_pad": PAD}


class JointSegmentationMonodepth(nn.Module):
    def __init__(self, models, frame_ids, use_pose_net, num_pose_frames, provide_uncropped_for_pose):
        super().__init__()
        self.models = nn.ModuleDict(models)
        self.frame_ids = frame_ids
        self.use_pose_net = use_pose_net            # train.py:664 toggles this on the EMA teacher
        self.num_pose_frames = num_pose_frames
        self.provide_uncropped_for_pose = provide_uncropped_for_pose
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    # ---------------------------------------------------------------------------------------------------- poses
    def _pose_input_key(self):
        return "color_full_aug" if self.provide_uncropped_for_pose else "color_aug"

    def _poses_from_pairs(self, inputs):
        """One pose-net pass per source frame on the (earlier, later) frame pair; the relative pose of a past frame is
        the inverse of the predicted one (reference :24-50)."""
        key, out = self._pose_input_key(), {}
        enc, dec = self.models["pose_encoder"], self.models["pose"]
        for f in self.frame_ids[1:]:
            if f == "s":
                continue
            target, source = inputs[key, 0, 0], inputs[key, f, 0]
            first, second = (source, target) if f < 0 else (target, source)
            feats = enc.forward_pair(first, second) if hasattr(enc, "forward_pair") else enc(torch.cat([first, second], 1))
            axisangle, translation = dec([feats])
            out[("axisangle", 0, f)], out[("translation", 0, f)] = axisangle, translation
            out[("cam_T_cam", 0, f)] = transformation_from_parameters(axisangle[:, 0], translation[:, 0], invert=f < 0)
        return out

    def _poses_from_all_frames(self, inputs):
        """All frames through the pose net at once; output slot k belongs to source frame k (reference :52-68)."""
        key, out = self._pose_input_key(), {}
        frames = [inputs[(key, f, 0)] for f in self.frame_ids if f != "s"]
        axisangle, translation = self.models["pose"]([self.models["pose_encoder"](torch.cat(frames, 1))])
        for k, f in enumerate(self.frame_ids[1:]):
            if f == "s":
                continue
            out[("axisangle", 0, f)], out[("translation", 0, f)] = axisangle, translation
            out[("cam_T_cam", 0, f)] = transformation_from_parameters(axisangle[:, k], translation[:, k])
        return out

    def predict_poses(self, inputs, features):
        return self._poses_from_pairs(inputs) if self.num_pose_frames == 2 else self._poses_from_all_frames(inputs)

    # -------------------------------------------------------------------------------------------------- forward
    def predict_test_disp(self, x):
        return self.models["depth"](self.models["encoder"](x[("color", 0, 0)]))

    def forward(self, x):
        """inputs dict (tuple keys, loader contract) -> outputs dict (reference :77-100)."""
        m = self.models
        image = x["color_aug", 0, 0]
        features = m["encoder"](image)
        out = {"bottleneck": features[-1]}
        if "mtl_decoder" in m:                      # PAD: depth + segmentation from one multi-task decoder
            out.update(m["mtl_decoder"](features))
        else:
            if "depth" in m:
                out.update(m["depth"](features))
            if "segmentation" in m:
                out["semantics"] = m["segmentation"](features)
        if "imnet_encoder" in m:                    # frozen ImageNet twin for the feature-distance loss (train.py:480)
            m["imnet_encoder"].eval()
            with torch.no_grad():
                reference_features = m["imnet_encoder"](image)[-1].detach()
            out["encoder_features"], out["imnet_features"] = features[-1], reference_features
        if self.use_pose_net:
            out.update(self.predict_poses(x, features))
        return out


JointSegmentationDepth = JointSegmentationMonodepth   # name used by BASELINE.json


def get_segmentation_network(segmentation_name, num_ch_enc, segmentation_size, num_classes, segmentation_args,
                             depth_args):
    decoder_cls = SEGMENTATION_DECODERS[segmentation_name]
    num_ch_dec = depth_args.get("num_ch_dec", [16, 32, 64, 128, 256])
    return decoder_cls(num_ch_enc, num_ch_dec, num_classes, depth_args=depth_args, **segmentation_args)


def _freeze(module, what=None):
    if what:
        print("Freeze %s weights" % what)
    for p in module.parameters():
        p.requires_grad = False


def joint_segmentation_depth(name, backbone_name, segmentation_name, segmentation_args,
                             num_classes, backbone_pretraining,
                             depth_pretraining, pose_pretraining, freeze_backbone,
                             freeze_segmentation,
                             freeze_depth, freeze_pose, replace_stride_with_dilation,
                             frame_ids, num_scales, pose_model_input, provide_uncropped_for_pose,
                             height, width, depth_args, disable_monodepth, enable_imnet_encoder,
                             disable_pose, imnet_encoder_dilation=True, **kwargs):
    """The YAML `model` section -> network (reference :116-183; same keywords, unknown ones are ignored as there)."""
    assert frame_ids[0] == 0
    monodepth = not disable_monodepth
    num_pose_frames = 2 if pose_model_input == "pairs" else len(frame_ids)
    use_pose_net = not disable_pose and frame_ids != (0, "s")

    # ---- sub-networks
    encoder = get_resnet_backbone(backbone_name, backbone_pretraining, replace_stride_with_dilation,
                                  use_intermediate_layer_getter=False)
    parts = {"encoder": encoder}
    if enable_imnet_encoder:
        parts["imnet_encoder"] = get_resnet_backbone(
            backbone_name, "imnet", use_intermediate_layer_getter=False,
            replace_stride_with_dilation=replace_stride_with_dilation if imnet_encoder_dilation else None)
    if use_pose_net and monodepth:
        parts.update(get_posenet("resnet18", backbone_pretraining, pose_pretraining, num_pose_frames))
    seg_args = (segmentation_name, encoder.num_ch_enc, (height, width), num_classes, segmentation_args, depth_args)
    if segmentation_name == "mtl_pad":
        parts["mtl_decoder"] = get_segmentation_network(*seg_args)
    else:
        if monodepth:
            parts["depth"] = get_depth_decoder(depth_pretraining, encoder.num_ch_enc, range(num_scales), **depth_args)
        if segmentation_name is not None:
            parts["segmentation"] = get_segmentation_network(*seg_args)

    # ---- freezing (requires_grad only; BatchNorm modes are the trainer's business, train.py:433-440)
    if "imnet_encoder" in parts:
        _freeze(parts["imnet_encoder"])
    if freeze_backbone:
        _freeze(encoder, "backbone")
    if monodepth and freeze_depth:
        _freeze(parts["depth"], "depth decoder")
    if monodepth and freeze_pose:
        for k in ("pose_encoder", "pose"):
            if k in parts or k == "pose":
                _freeze(parts[k], "pose decoder" if k == "pose" else None)
    if freeze_segmentation and "segmentation" in parts:
        _freeze(parts["segmentation"], "segmentation decoder")
    return JointSegmentationMonodepth(parts, frame_ids, use_pose_net, num_pose_frames, provide_uncropped_for_pose)
