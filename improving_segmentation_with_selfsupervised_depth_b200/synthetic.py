"""Seeded synthetic training samples of SURVEY.md §8(d): smooth sinusoid textures, source frames =
shifted target (+1 % noise), Cityscapes intrinsics scaled to the image size, block-constant labels.
Host-side data plumbing (the reference's loader is out of scope); tensors come back on the CPU."""
import math

import torch
import torch.nn.functional as F


def synthetic_inputs(B, H, W, seed=1234, num_scales=4, frame_ids=(0, -1, 1), labels=False):
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(H + 8, dtype=torch.float32), torch.arange(W + 8, dtype=torch.float32),
                            indexing="ij")
    tex = torch.zeros(B, 3, H + 8, W + 8)
    for _ in range(8):
        fx = (torch.rand(B, 3, 1, 1, generator=g) * 6 + 1) * 2 * math.pi / W
        fy = (torch.rand(B, 3, 1, 1, generator=g) * 6 + 1) * 2 * math.pi / H
        ph = torch.rand(B, 3, 1, 1, generator=g) * 2 * math.pi
        tex += torch.sin(xs * fx + ys * fy + ph) / 8
    tex = (0.5 + 0.4 * tex + 0.05 * torch.rand(B, 3, H + 8, W + 8, generator=g)).clamp(0, 1)
    inputs = {}
    shifts = {0: (0, 0), -1: (3, 1), 1: (-3, -1)}
    for f in frame_ids:
        dx, dy = shifts[f]
        img = tex[:, :, 4 + dy:4 + dy + H, 4 + dx:4 + dx + W].clone()
        if f != 0:
            img = (img + 0.01 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1)
        inputs[("color", f, 0)] = img.contiguous()
        inputs[("color_aug", f, 0)] = inputs[("color", f, 0)]
        for s in range(1, num_scales):
            inputs[("color", f, s)] = F.interpolate(img, size=(H // 2 ** s, W // 2 ** s), mode="area")
    for s in range(num_scales):
        K = torch.eye(4)
        K[0, 0], K[1, 1] = 2262.52 / 2048 * W, 2265.3017905988554 / 1024 * H
        K[0, 2], K[1, 2] = 1096.98 / 2048 * W, 513.137 / 1024 * H
        K[0] /= 2 ** s
        K[1] /= 2 ** s
        K[0, 3] = 0
        inputs[("K", s)] = K[None].repeat(B, 1, 1)
        inputs[("inv_K", s)] = torch.linalg.pinv(K)[None].repeat(B, 1, 1)
    if labels:
        blk = torch.randint(0, 19, (B, (H + 31) // 32, (W + 31) // 32), generator=g)
        ign = torch.rand(B, (H + 31) // 32, (W + 31) // 32, generator=g) < 0.1
        blk[ign] = 250
        inputs["lbl"] = blk.repeat_interleave(32, 1).repeat_interleave(32, 2)[:, :H, :W].contiguous()
    return inputs


def mono_config(backbone="resnet50", H=512, W=1024, freeze_backbone=True, enable_imnet_encoder=False,
                segmentation_name=None, segmentation_args=None):
    """kwargs of models.get_model for the dec5/dec6 architecture (configs/cityscapes_monodepth_highres_dec5_crop.yml
    with the backbone of BASELINE.json)."""
    return {
        "arch": "joint_segmentation_depth", "backbone_name": backbone,
        "replace_stride_with_dilation": [False, False, True],
        "segmentation_name": segmentation_name, "segmentation_args": segmentation_args,
        "depth_args": {"intermediate_aspp": True, "aspp_rates": [6, 12, 18], "n_upconv": 4,
                       "num_ch_dec": [64, 128, 128, 256, 256], "max_scale_size": [H, W]},
        "pose_model_input": "pairs", "backbone_pretraining": "none", "depth_pretraining": "none",
        "pose_pretraining": "none", "freeze_backbone": freeze_backbone, "freeze_depth": False,
        "freeze_pose": False, "freeze_segmentation": segmentation_name is None, "disable_monodepth": False,
        "disable_pose": False, "enable_imnet_encoder": enable_imnet_encoder, "provide_uncropped_for_pose": False,
        "frame_ids": [0, -1, 1], "num_scales": 4, "height": H, "width": W, "crop_h": H, "crop_w": W,
    }


MONO_LOSS_KW = dict(num_scales=4, frame_ids=[0, -1, 1], min_depth=0.1, max_depth=100, test_min_depth=1e-3,
                    test_max_depth=80, disparity_smoothness=1e-3, no_ssim=False, avg_reprojection=False,
                    disable_automasking=False)
