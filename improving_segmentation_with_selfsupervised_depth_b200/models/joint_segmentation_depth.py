"""JointSegmentationMonodepth + factory — drop-in for models/joint_segmentation_depth.py."""
import torch
from torch import nn

from .. import ops
from .joint_segmentation_depth_decoder import PAD, JointSegDepthDecoder
from .monodepth_layers import transformation_from_parameters
from .utils import get_depth_decoder, get_posenet, get_resnet_backbone


class JointSegmentationMonodepth(nn.Module):
    def __init__(self, models, frame_ids, use_pose_net, num_pose_frames, provide_uncropped_for_pose):
        super().__init__()
        self.frame_ids = frame_ids
        self.use_pose_net = use_pose_net
        self.num_pose_frames = num_pose_frames
        self.provide_uncropped_for_pose = provide_uncropped_for_pose
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.models = nn.ModuleDict(models)

    def predict_poses(self, inputs, features):
        """Reference :20-70."""
        outputs = {}
        key = "color_full_aug" if self.provide_uncropped_for_pose else "color_aug"
        enc, dec = self.models["pose_encoder"], self.models["pose"]
        if self.num_pose_frames == 2:
            for f_i in self.frame_ids[1:]:
                if f_i == "s":
                    continue
                # frames always enter the pose network in temporal order
                a, b = (inputs[key, f_i, 0], inputs[key, 0, 0]) if f_i < 0 else (inputs[key, 0, 0], inputs[key, f_i, 0])
                feats = [enc.forward_pair(a, b)] if hasattr(enc, "forward_pair") else [enc(torch.cat([a, b], 1))]
                axisangle, translation = dec(feats)
                outputs[("axisangle", 0, f_i)] = axisangle
                outputs[("translation", 0, f_i)] = translation
                outputs[("cam_T_cam", 0, f_i)] = transformation_from_parameters(
                    axisangle[:, 0], translation[:, 0], invert=(f_i < 0))
        else:
            frames = [inputs[(key, i, 0)] for i in self.frame_ids if i != "s"]
            axisangle, translation = dec([enc(torch.cat(frames, 1))])
            for i, f_i in enumerate(self.frame_ids[1:]):
                if f_i != "s":
                    outputs[("axisangle", 0, f_i)] = axisangle
                    outputs[("translation", 0, f_i)] = translation
                    outputs[("cam_T_cam", 0, f_i)] = transformation_from_parameters(axisangle[:, i], translation[:, i])
        return outputs

    def predict_test_disp(self, x):
        return self.models["depth"](self.models["encoder"](x[("color", 0, 0)]))

    def forward(self, x):
        """Reference :77-100."""
        outputs, inputs = {}, x
        features = self.models["encoder"](inputs["color_aug", 0, 0])
        outputs["bottleneck"] = features[-1]
        if "mtl_decoder" in self.models:
            outputs.update(self.models["mtl_decoder"](features))
        else:
            if "depth" in self.models:
                outputs.update(self.models["depth"](features))
            if "segmentation" in self.models:
                outputs["semantics"] = self.models["segmentation"](features)
        if "imnet_encoder" in self.models:
            outputs["encoder_features"] = features[-1]
            self.models["imnet_encoder"].eval()
            with torch.no_grad():
                outputs["imnet_features"] = self.models["imnet_encoder"](inputs["color_aug", 0, 0])[-1].detach()
        if self.use_pose_net:
            outputs.update(self.predict_poses(inputs, features))
        return outputs


JointSegmentationDepth = JointSegmentationMonodepth   # name used by BASELINE.json


def get_segmentation_network(segmentation_name, num_ch_enc, segmentation_size, num_classes, segmentation_args,
                             depth_args):
    model_map = {'joint_seg_depth_dec': JointSegDepthDecoder, 'mtl_pad': PAD}
    num_ch_dec = depth_args.get("num_ch_dec", [16, 32, 64, 128, 256])
    return model_map[segmentation_name](num_ch_enc, num_ch_dec, num_classes, **segmentation_args,
                                        depth_args=depth_args)


def _freeze(module):
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
    """Reference :116-183 (same kwargs; unknown ones are swallowed like the reference does)."""
    num_pose_frames = 2 if pose_model_input == "pairs" else len(frame_ids)
    assert frame_ids[0] == 0
    use_pose_net = not (frame_ids == (0, "s")) and not disable_pose
    models = {"encoder": get_resnet_backbone(backbone_name, backbone_pretraining, replace_stride_with_dilation,
                                             use_intermediate_layer_getter=False)}
    num_ch_enc = models["encoder"].num_ch_enc
    if enable_imnet_encoder:
        models["imnet_encoder"] = get_resnet_backbone(
            backbone_name, 'imnet',
            replace_stride_with_dilation=replace_stride_with_dilation if imnet_encoder_dilation else None,
            use_intermediate_layer_getter=False)
        _freeze(models["imnet_encoder"])
    if use_pose_net and not disable_monodepth:
        models.update(get_posenet("resnet18", backbone_pretraining, pose_pretraining, num_pose_frames))
    if segmentation_name in ["mtl_pad"]:
        models["mtl_decoder"] = get_segmentation_network(segmentation_name, num_ch_enc, (height, width),
                                                         num_classes, segmentation_args, depth_args)
    else:
        if not disable_monodepth:
            models["depth"] = get_depth_decoder(depth_pretraining, num_ch_enc, range(num_scales), **depth_args)
        if segmentation_name is not None:
            models["segmentation"] = get_segmentation_network(segmentation_name, num_ch_enc, (height, width),
                                                              num_classes, segmentation_args, depth_args)
    if freeze_backbone:
        print('Freeze backbone weights')
        _freeze(models["encoder"])
    if not disable_monodepth and freeze_depth:
        print('Freeze depth decoder weights')
        _freeze(models["depth"])
    if not disable_monodepth and freeze_pose:
        print('Freeze pose decoder weights')
        if "pose_encoder" in models:
            _freeze(models["pose_encoder"])
        _freeze(models["pose"])
    if "segmentation" in models and freeze_segmentation:
        print('Freeze segmentation decoder weights')
        _freeze(models["segmentation"])
    return JointSegmentationMonodepth(models, frame_ids, use_pose_net, num_pose_frames, provide_uncropped_for_pose)
