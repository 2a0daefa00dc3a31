"""Segmentation decoders — drop-in for models/joint_segmentation_depth_decoder.py (JointSegDepthDecoder
:11-75, PAD :78-184) on the sm_100a kernels."""
import numpy as np
from torch import nn

from .. import _cabi as A
from .. import ops
from .layers import BatchNorm2d, Conv2d, Dropout
from .model_parts import SelfAttention
from .utils import _get_layer, get_depth_decoder


class _Head(nn.Sequential):
    """[Dropout] -> [3x3 conv -> BN -> ReLU -> Dropout] -> 1x1 classifier, keys as the reference's nn.Sequential."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, Conv2d) and i + 1 < len(mods) and isinstance(mods[i + 1], BatchNorm2d):
                x = mods[i + 1](m(x), act=A.ACT_RELU)
                i += 3          # conv, bn, relu
            else:
                x = m(x)
                i += 1
        return x


class JointSegDepthDecoder(nn.Module):
    first_iter = True

    def __init__(self, num_ch_enc, num_ch_dec, num_classes, layers=None,
                 head_inter_channels=64, weights='none',
                 head_dropout=0.1, layer_dropout=0, output_stride=1, layer_out_channels=64,
                 depth_args=None, head_inter=True):
        super().__init__()
        if layers is None:
            layers = [9]
        self.output_stride = output_stride
        self.num_ch_enc, self.num_ch_dec, self.num_classes, self.layers = num_ch_enc, num_ch_dec, num_classes, layers
        assert len(num_ch_enc) == 5
        assert len(num_ch_dec) == 5
        self.last_layer = len(num_ch_enc) + len(num_ch_dec) - 1
        self.unet_dec = get_depth_decoder(weights, num_ch_enc, **depth_args)
        accumulated_ch = 0
        project = {}
        for layer in layers:
            ch = num_ch_enc[layer] if layer <= 4 else num_ch_dec[self.last_layer - layer]
            accumulated_ch += layer_out_channels
            project[f"seg{layer}"] = nn.Sequential(Conv2d(int(ch), layer_out_channels, 1, bias=False))
        self.project = nn.ModuleDict(project)
        if head_inter:
            head_conv = [Conv2d(accumulated_ch, head_inter_channels, 3, padding=1, bias=False),
                         BatchNorm2d(head_inter_channels), nn.ReLU(), Dropout(head_dropout)]
        else:
            head_conv = [nn.Identity()]
        self.head = _Head(Dropout(layer_dropout) if layer_dropout > 0 else nn.Identity(), *head_conv,
                          Conv2d(head_inter_channels, self.num_classes, 1))

    def forward(self, encoder_features):
        seg_features = self.unet_dec(encoder_features)
        segmentation_size = tuple(_get_layer(encoder_features, seg_features, self.last_layer).shape[2:])
        last_layer_size = tuple(int(v) for v in np.array(segmentation_size) // self.output_stride)
        stacked = []
        for layer in self.layers:
            if f"seg{layer}" in self.project:
                proj = self.project[f"seg{layer}"][0](_get_layer(encoder_features, seg_features, layer))
                if tuple(proj.shape[2:]) != last_layer_size:
                    proj = ops.bilinear(proj, last_layer_size, align_corners=False)
                stacked.append(proj)
        stacked = stacked[0] if len(stacked) == 1 else ops.cat_channels(stacked)
        score = self.head(stacked)
        if last_layer_size != segmentation_size:
            score = ops.bilinear(score, segmentation_size, align_corners=False)
        JointSegDepthDecoder.first_iter = False
        return score


class PAD(nn.Module):
    first_iter = True

    def __init__(self, num_ch_enc, num_ch_dec, num_classes, final_layer=9,
                 weights=None, output_stride=1, depth_args=None, distillation_layer=7, side_output=True):
        super().__init__()
        self.output_stride = output_stride
        self.num_ch_enc, self.num_ch_dec, self.num_classes = num_ch_enc, num_ch_dec, num_classes
        self.side_output = side_output
        assert len(num_ch_enc) == 5
        assert len(num_ch_dec) == 5
        self.final_layer = final_layer
        self.last_layer = len(num_ch_enc) + len(num_ch_dec) - 1
        self.distillation_layer = distillation_layer
        self.dec_n_upconv = depth_args.get("n_upconv", 4)
        distillation_ch = int(self.layer_channels(self.distillation_layer))
        final_ch = int(self.layer_channels(self.final_layer))
        num_scales = 4
        self.depth_dec = get_depth_decoder(weights, num_ch_enc, range(num_scales), **depth_args)
        self.seg_dec = get_depth_decoder(weights, num_ch_enc, range(num_scales), **depth_args)
        self.seg_dec.enable_disparity = False
        for s in range(num_scales):
            self.seg_dec.convs[("dispconv", s)] = nn.Identity()
        self.sa_depth = SelfAttention(distillation_ch, distillation_ch)
        self.sa_seg = SelfAttention(distillation_ch, distillation_ch)
        if self.side_output:
            self.seg_intermediate_head = nn.Sequential(Conv2d(distillation_ch, self.num_classes, 1))
        self.seg_final_head = nn.Sequential(Conv2d(final_ch, self.num_classes, 1))

    def layer_channels(self, layer):
        return self.num_ch_enc[layer] if layer <= 4 else self.num_ch_dec[self.last_layer - layer]

    def depth_params(self):
        return [*self.depth_dec.parameters(), *self.sa_seg.parameters()]

    def segmentation_params(self):
        params = [*self.seg_dec.parameters(), *self.sa_depth.parameters(), *self.seg_final_head.parameters()]
        if self.side_output:
            params.extend(self.seg_intermediate_head.parameters())
        return params

    # The decoder pair runs as a two-phase schedule around one exchange point (reference :135-184):
    #   phase A: both U-Net decoders from the bottleneck down to the distillation layer,
    #   exchange: each branch receives the other's features through a sigmoid gate (SelfAttention),
    #   phase B: both decoders continue from the exchanged tensors to full resolution.
    def _schedule(self):
        split = self.last_layer - self.distillation_layer            # index of the "upconv" stage at the exchange
        return split, list(range(self.dec_n_upconv, split - 1, -1)), list(range(split - 1, -1, -1))

    def _exchange(self, depth_mid, seg_mid):
        gated_depth = self.sa_depth(depth_mid)      # depth features offered to the segmentation branch
        gated_seg = self.sa_seg(seg_mid)            # segmentation features offered to the depth branch
        return ops.add(depth_mid, gated_seg), ops.add(seg_mid, gated_depth), gated_depth

    def _to_label_size(self, logits, size):
        small = tuple(int(v) for v in np.array(size) // self.output_stride)
        return logits if small == size else ops.bilinear(logits, size, align_corners=False)

    def forward(self, encoder_features):
        size = tuple(encoder_features[0].shape[2:])
        split, phase_a, phase_b = self._schedule()
        key = ("upconv", split)
        branches = {"depth": self.depth_dec, "seg": self.seg_dec}
        feats = {name: dec(encoder_features, exec_layer=phase_a) for name, dec in branches.items()}
        side = self.seg_intermediate_head[0](feats["seg"][key]) if self.side_output else None
        to_depth, to_seg, gated_depth = self._exchange(feats["depth"][key], feats["seg"][key])
        out = dict(feats["depth"])
        out.update(self.depth_dec(encoder_features, x=to_depth, exec_layer=phase_b))
        seg_tail = self.seg_dec(encoder_features, x=to_seg, exec_layer=phase_b)
        out["semantics"] = self._to_label_size(
            self.seg_final_head[0](_get_layer(gated_depth, seg_tail, self.final_layer)), size)
        if side is not None:
            out["intermediate_semantics"] = self._to_label_size(side, size)
        PAD.first_iter = False
        return out
