"""CPU: the drop-in modules expose the reference's state_dict keys, shapes and trainable set."""
import contextlib
import io

import pytest
import torch

import improving_segmentation_with_selfsupervised_depth_b200 as P


@pytest.mark.parametrize("name", ["mono_r50", "mono_r18", "pad_r50", "segdec_r50"])
def test_state_dict_contract(contracts, name):
    models, _ = P.install_dropin()
    c = contracts[name]
    with contextlib.redirect_stdout(io.StringIO()):
        m = models.get_model(dict(c["cfg"]), 19)
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == c["state_dict"]
    assert [k for k, p in m.named_parameters() if p.requires_grad] == c["trainable"]
    # loading a reference-layout (contiguous OIHW) checkpoint keeps the kernels' channels_last weights
    sd = {k: torch.zeros(s) for k, s in c["state_dict"].items()}
    m.load_state_dict(sd)
    w = m.models["encoder"].encoder.layer1[0].conv2.weight
    assert w.permute(0, 2, 3, 1).is_contiguous()


def test_dropin_import_surface():
    P.install_dropin()
    import loss
    import models
    from loss import get_monodepth_loss, get_segmentation_loss_function, key2loss   # noqa
    from loss.loss import berhu, cross_entropy2d, pixel_wise_entropy           # noqa
    from models import get_model                                               # noqa
    from models.joint_segmentation_depth_decoder import PAD, JointSegDepthDecoder   # noqa
    from models.joint_segmentation_depth import JointSegmentationDepth, JointSegmentationMonodepth   # noqa
    from models.model_parts import ASPP, SelfAttention                         # noqa
    from models.monodepth_layers import (SSIM, BackprojectDepth, Conv3x3, ConvBlock, Project3D,   # noqa
                                         disp_to_depth, get_smooth_loss, get_translation_matrix,
                                         rot_from_axisangle, transformation_from_parameters, upsample)
    from models.utils import (_get_layer, download_model_if_doesnt_exist, get_depth_decoder,   # noqa
                              get_posenet, get_resnet_backbone)
    with pytest.raises(NotImplementedError):
        models.get_model({"arch": "nope"}, 19)
    cfg = {"training": {"segmentation_loss": {"name": "cross_entropy"}, "batch_size": 2,
                        "monodepth_loss": dict(num_scales=4, frame_ids=[0, -1, 1], height=64, width=96,
                                               min_depth=0.1, max_depth=100, test_min_depth=1e-3, test_max_depth=80,
                                               disparity_smoothness=1e-3, no_ssim=False, avg_reprojection=False,
                                               disable_automasking=False, crop_h=32, crop_w=48)}}
    ml = loss.get_monodepth_loss(cfg, is_train=True)
    assert (ml.height, ml.width, ml.batch_size) == (32, 48, 2)
    assert callable(loss.get_segmentation_loss_function(cfg))
