"""ORACLE — test infrastructure only (never imported by the product package).

A CPU, fp32, plain-PyTorch *functional restatement* of the reference hot path
(lhoyer/improving_segmentation_with_selfsupervised_depth): the network forward is written as pure
functions over the reference's own `state_dict` (same keys / OIHW tensors), the loss as pure functions
over tensors.  Gradients come from torch autograd on CPU.  Every function cites the reference
file:line it restates.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module.

Pinning: `oracle/validate_against_reference.py` (run in the build container, where /root/reference is
importable) checks every function here against the unmodified reference on seeded inputs, and
`tests/golden/make_golden.py` stores reference outputs as fixtures which `tests/test_oracle_golden.py`
replays against this file on any machine.  Third-party arithmetic the reference delegates to
(torch conv/BN/grid_sample/interpolate, torchvision ResNet/ASPP) is called here through the same
torch functional ops — the reference pins none of it with its own tests.
"""
import math

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# loss path
# ------------------------------------------------------------------------------------------------


def disp_to_depth(disp, min_depth, max_depth):
    """models/monodepth_layers.py:18-27."""
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1.0 / scaled


def backproject(depth, inv_K):
    """BackprojectDepth.forward, models/monodepth_layers.py:169-174 (pixel grid built at :155-167)."""
    B, _, H, W = depth.shape
    dev, dt = depth.device, depth.dtype       # the reference keeps these buffers on the module's device (:155-167)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt, device=dev), torch.arange(W, dtype=dt, device=dev), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=dt, device=dev)], 0).unsqueeze(0).expand(B, 3, H * W)
    cam = torch.matmul(inv_K[:, :3, :3], pix) * depth.reshape(B, 1, -1)
    return torch.cat([cam, torch.ones(B, 1, H * W, dtype=dt, device=dev)], 1)


def project(points, K, T, H, W, eps=1e-7):
    """Project3D.forward, models/monodepth_layers.py:188-199."""
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    pix = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + eps)
    pix = pix.view(B, 2, H, W).permute(0, 2, 3, 1)
    norm = torch.tensor([W - 1, H - 1], dtype=pix.dtype, device=pix.device)
    return (pix / norm - 0.5) * 2


def ssim(x, y):
    """SSIM.forward, models/monodepth_layers.py:240-254 (C1=0.01^2, C2=0.03^2)."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sig_x = F.avg_pool2d(x * x, 3, 1) - mu_x ** 2
    sig_y = F.avg_pool2d(y * y, 3, 1) - mu_y ** 2
    sig_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + 0.01 ** 2) * (2 * sig_xy + 0.03 ** 2)
    d = (mu_x ** 2 + mu_y ** 2 + 0.01 ** 2) * (sig_x + sig_y + 0.03 ** 2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def photometric(pred, target, no_ssim=False):
    """compute_reprojection_loss, loss/monodepth_loss.py:104-116."""
    l1 = (target - pred).abs().mean(1, True)
    if no_ssim:
        return l1
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def smoothness(disp, img):
    """get_smooth_loss, models/monodepth_layers.py:208-221."""
    gdx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    gdy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs()
    gix = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True)
    giy = (img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True)
    return (gdx * torch.exp(-gix)).mean() + (gdy * torch.exp(-giy)).mean()


def warp(inputs, disp, T, frame_id, H, W, min_depth, max_depth):
    """One (scale, frame) iteration of generate_images_pred, loss/monodepth_loss.py:70-98."""
    up = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
    _, depth = disp_to_depth(up, min_depth, max_depth)
    grid = project(backproject(depth, inputs[("inv_K", 0)]), inputs[("K", 0)], T, H, W)
    pred = F.grid_sample(inputs[("color", frame_id, 0)], grid, padding_mode="border", align_corners=True)
    return depth, grid, pred


def monodepth_loss(inputs, disps, cam_T_cam, frame_ids, H, W, min_depth=0.1, max_depth=100.0,
                   disparity_smoothness=1e-3, no_ssim=False, avg_reprojection=False,
                   disable_automasking=False, noise=None, extras=None):
    """generate_images_pred + compute_losses, loss/monodepth_loss.py:64-192.

    disps: list over scales of Bx1xhxw; cam_T_cam: {frame_id: Bx4x4}; noise: list over scales of the
    tie-break noise ALREADY multiplied by 1e-5 (the reference draws torch.randn on the CPU at :163-164).
    """
    S = len(disps)
    losses, total = {}, 0
    target = inputs[("color", 0, 0)]
    for s in range(S):
        reproj = []
        for f in frame_ids[1:]:
            T = inputs["stereo_T"] if f == "s" else cam_T_cam[f]
            depth, grid, pred = warp(inputs, disps[s], T, f, H, W, min_depth, max_depth)
            if extras is not None:
                extras[("depth", 0, s)], extras[("sample", f, s)], extras[("color", f, s)] = depth, grid, pred
            reproj.append(photometric(pred, target, no_ssim))
        reproj = torch.cat(reproj, 1)
        if avg_reprojection:
            reproj = reproj.mean(1, keepdim=True)
        if not disable_automasking:
            ident = torch.cat([photometric(inputs[("color", f, 0)], target, no_ssim) for f in frame_ids[1:]], 1)
            if avg_reprojection:
                ident = ident.mean(1, keepdim=True)
            ident = ident + noise[s]
            combined = torch.cat((ident, reproj), dim=1)
        else:
            combined = reproj
        if combined.shape[1] == 1:
            chosen = combined
        else:
            chosen, idx = torch.min(combined, dim=1)
            if extras is not None and not disable_automasking:
                extras["identity_selection/%d" % s] = (idx > ident.shape[1] - 1).float()
        loss = chosen.mean()
        d = disps[s]
        norm = d / (d.mean(2, True).mean(3, True) + 1e-7)
        loss = loss + disparity_smoothness * smoothness(norm, inputs[("color", 0, s)]) / (2 ** s)
        total = total + loss
        losses["loss/%d" % s] = loss
    losses["loss"] = total / S
    return losses


def cross_entropy2d(logits, target, class_weight=None, pixel_weights=None):
    """loss/loss.py:17-37."""
    n, c, h, w = logits.shape
    _, ht, wt = target.shape
    if h != ht and w != wt:
        logits = F.interpolate(logits, size=(ht, wt), mode="bilinear", align_corners=True)
    flat = logits.permute(0, 2, 3, 1).reshape(-1, c)
    loss = F.cross_entropy(flat, target.reshape(-1), weight=class_weight,
                           reduction="mean" if pixel_weights is None else "none", ignore_index=250)
    if pixel_weights is not None:
        loss = (pixel_weights.reshape(-1).detach() * loss).mean()
    return loss


def berhu(pred, target, mask, apply_log=False, threshold=0.2):
    """loss/loss.py:5-15: reverse Huber on the masked absolute difference; the switch point C = threshold * max is a
    plain number (no gradient flows through it)."""
    if apply_log:
        pred, target = torch.log(1 + pred), torch.log(1 + target)
    a = (target - pred).abs() * mask
    C = threshold * float(a.detach().max())
    return torch.where(a <= C, a, (a * a + C * C) / (2 * C)).mean()


def pixel_wise_entropy(logits, normalize=False):
    """loss/loss.py:40-47: per-pixel entropy of the channel softmax, in units of log2(C)."""
    p = F.softmax(logits, dim=1)
    ent = -(p * torch.log2(p + 1e-30)).sum(1) / math.log2(logits.shape[1])
    if normalize:
        ent = (ent - ent.min()) / (ent.max() - ent.min())
    return ent


# ------------------------------------------------------------------------------------------------
# step-level ops of train.py
# ------------------------------------------------------------------------------------------------
def feature_distance(a, b):
    """train.py:482."""
    return torch.dist(a, b, p=2)


def normalize_depths(disp):
    """train.py:688-692 (the clamp to [min, max] there is the identity)."""
    d = disp.detach().clone()
    for j in range(d.shape[0]):
        lo, hi = d[j].min(), d[j].max()
        d[j] = (d[j] - lo) / (hi - lo)
    return d


def depthcomp_mix_mask(depths, margin, foreground_threshold):
    """train.py:585-604 with the pairing (i, (i+1) % B) — for B = 2 the reference's (0,1), (1,0)."""
    B = depths.shape[0]
    out = []
    for i in range(B):
        own, other = depths[i], depths[(i + 1) % B]
        m = torch.ge(own, other - margin).long() * torch.ge(own, foreground_threshold).long()
        out.append(m)
    return torch.cat(out)


def mix(mask, data):
    """loader/transformsgpu.py:33-47, branch mask.shape[0] == data.shape[0]."""
    B = data.shape[0]
    return torch.stack([mask[i] * data[i] + (1 - mask[i]) * data[(i + 1) % B] for i in range(B)])


def calc_pseudo_label_loss(teacher_softmax, student_logits, consistency_weight=1.0, ignore_index=250):
    """train.py:644-651."""
    max_probs, label = torch.max(teacher_softmax, dim=1)
    label = label.clone()
    label[max_probs == 0] = ignore_index
    w = float((max_probs >= 0.968).sum()) / label.numel()
    weights = w * torch.ones_like(max_probs)
    return consistency_weight * cross_entropy2d(student_logits, label, pixel_weights=weights), label


def update_ema(ema_params, params, alpha_teacher, iteration):
    """train.py:346-358 (in place on the list of ema tensors)."""
    alpha = min(1 - 1 / (iteration + 1), alpha_teacher)
    for e, p in zip(ema_params, params):
        e[:] = alpha * e + (1 - alpha) * p
    return alpha


# ------------------------------------------------------------------------------------------------
# pose geometry
# ------------------------------------------------------------------------------------------------


def rot_from_axisangle(vec):
    """models/monodepth_layers.py:66-105 (vec: Bx1x3)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    Cc = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    rows = [
        [x * x * Cc + ca, x * y * Cc - z * sa, z * x * Cc + y * sa],
        [x * y * Cc + z * sa, y * y * Cc + ca, y * z * Cc - x * sa],
        [z * x * Cc - y * sa, y * z * Cc + x * sa, z * z * Cc + ca],
    ]
    B = vec.shape[0]
    top = torch.stack([torch.stack([e.reshape(B) for e in r], 1) for r in rows], 1)   # B,3,3
    R = torch.zeros(B, 4, 4, device=vec.device)
    R[:, :3, :3] = top
    R[:, 3, 3] = 1
    return R


def transformation_from_parameters(axisangle, translation, invert=False):
    """models/monodepth_layers.py:30-63."""
    R = rot_from_axisangle(axisangle)
    t = translation.reshape(-1, 3, 1)
    if invert:
        R = R.transpose(1, 2)
        t = -t
    B = t.shape[0]
    Tm = torch.cat([torch.cat([torch.eye(3, device=t.device).expand(B, 3, 3), t], 2),
                    torch.tensor([0., 0., 0., 1.], device=t.device).expand(B, 1, 4)], 1)
    return torch.matmul(R, Tm) if invert else torch.matmul(Tm, R)


# ------------------------------------------------------------------------------------------------
# network path: pure functions over the reference state_dict
# ------------------------------------------------------------------------------------------------

RESNET_BLOCKS = {18: ("basic", [2, 2, 2, 2]), 34: ("basic", [3, 4, 6, 3]), 50: ("bottleneck", [3, 4, 6, 3]),
                 101: ("bottleneck", [3, 4, 23, 3]), 152: ("bottleneck", [3, 8, 36, 3])}


class BNMode:
    """How BatchNorm layers behave: training=True uses batch statistics and (like torch) updates the
    running buffers of `sd` in place with momentum 0.1."""

    def __init__(self, training=True, momentum=0.1, eps=1e-5):
        self.training, self.momentum, self.eps = training, momentum, eps


def _bn(sd, p, x, mode):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        mode.training, mode.momentum, mode.eps)


def _conv(sd, p, x, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride, padding, dilation)


def resnet_features(sd, prefix, x, num_layers, replace_stride_with_dilation=None, mode=None):
    """ResnetEncoder.forward (models/resnet_encoder.py:90-101) over torchvision ResNet v1.5 blocks
    (stride on the 3x3; replace_stride_with_dilation as torchvision `_make_layer`)."""
    mode = mode or BNMode()
    kind, blocks = RESNET_BLOCKS[num_layers]
    rswd = replace_stride_with_dilation or [False, False, False]
    e = prefix
    feats = []
    x = (x - 0.45) / 0.225
    x = F.relu(_bn(sd, e + "bn1", _conv(sd, e + "conv1", x, 2, 3), mode))
    feats.append(x)
    x = F.max_pool2d(x, 3, 2, 1)
    dilation = 1
    for li in range(4):
        stride = 1 if li == 0 else 2
        prev_dil = dilation
        if li > 0 and rswd[li - 1]:
            dilation *= stride
            stride = 1
        for bi in range(blocks[li]):
            p = "%slayer%d.%d." % (e, li + 1, bi)
            s = stride if bi == 0 else 1
            dil = prev_dil if bi == 0 else dilation
            identity = x
            if kind == "bottleneck":
                out = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x), mode))
                out = F.relu(_bn(sd, p + "bn2", _conv(sd, p + "conv2", out, s, dil, dil), mode))
                out = _bn(sd, p + "bn3", _conv(sd, p + "conv3", out), mode)
            else:
                out = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x, s, dil, dil), mode))
                out = _bn(sd, p + "bn2", _conv(sd, p + "conv2", out, 1, dilation, dilation), mode)
            if (p + "downsample.0.weight") in sd:
                identity = _bn(sd, p + "downsample.1", _conv(sd, p + "downsample.0", x, s), mode)
            x = F.relu(out + identity)
        feats.append(x)
    return feats


def _conv3x3_refl(sd, p, x):
    """Conv3x3, models/monodepth_layers.py:127-142."""
    return _conv(sd, p + ".conv", F.pad(x, (1, 1, 1, 1), mode="reflect"))


def _convblock(sd, p, x, mode):
    """ConvBlock, models/monodepth_layers.py:108-124 (bn optional, dropout2d off in all shipped configs)."""
    x = _conv3x3_refl(sd, p + ".block.0", x)
    if (p + ".block.1.weight") in sd:
        x = _bn(sd, p + ".block.1", x, mode)
    return F.elu(x)


def _aspp(sd, p, x, rates, pooling, mode, dropout_mask=None):
    """ASPP, models/model_parts.py:5-32 with torchvision ASPPConv / ASPPPooling."""
    res = [F.relu(_bn(sd, p + ".convs.0.1", _conv(sd, p + ".convs.0.0", x), mode))]
    for i, r in enumerate(rates):
        q = "%s.convs.%d" % (p, i + 1)
        res.append(F.relu(_bn(sd, q + ".1", _conv(sd, q + ".0", x, 1, r, r), mode)))
    if pooling:
        q = "%s.convs.%d" % (p, len(rates) + 1)
        g = F.adaptive_avg_pool2d(x, 1)
        g = F.relu(_bn(sd, q + ".2", _conv(sd, q + ".1", g), mode))
        res.append(F.interpolate(g, size=x.shape[-2:], mode="bilinear", align_corners=False))
    y = F.relu(_bn(sd, p + ".project.1", _conv(sd, p + ".project.0", torch.cat(res, 1)), mode))
    if mode.training:      # nn.Dropout(0.5), model_parts.py:25 — mask supplied by the caller (replay)
        if dropout_mask is None:
            dropout_mask = (torch.rand_like(y) >= 0.5).float()
        y = y * dropout_mask / 0.5
    return y


def depth_decoder(sd, prefix, feats, scales=(0, 1, 2, 3), n_upconv=4, intermediate_aspp=True,
                  aspp_rates=(6, 12, 18), aspp_pooling=True, use_skips=True, mode=None, dropout_mask=None,
                  x=None, exec_layer=None, enable_disparity=True):
    """DepthDecoder.forward, models/depth_decoder.py:75-116.  Module indices inside `decoder` follow the
    construction order at :43-73: per level i=n..0 -> upconv(i,0), [skip_proj(i)], upconv(i,1); then dispconvs."""
    mode = mode or BNMode()
    out = {}
    idx = {}
    k = 0
    for i in range(n_upconv, -1, -1):
        idx[("upconv", i, 0)] = k; k += 1
        if use_skips and i > 0:
            idx[("skip_proj", i)] = k; k += 1
        idx[("upconv", i, 1)] = k; k += 1
    for s in scales:
        idx[("dispconv", s)] = k; k += 1
    if x is None:
        x = feats[-1]
    for i in range(n_upconv, -1, -1):
        if exec_layer is not None and i not in exec_layer:
            continue
        p = "%sdecoder.%d" % (prefix, idx[("upconv", i, 0)])
        if i == n_upconv and intermediate_aspp:
            x = _aspp(sd, p, x, aspp_rates, aspp_pooling, mode, dropout_mask)
        else:
            x = _convblock(sd, p, x, mode)
        if x.shape[-1] < feats[i - 1].shape[-1] or i == 0:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        if use_skips and i > 0:
            x = torch.cat([x, feats[i - 1]], 1)
        x = _convblock(sd, "%sdecoder.%d" % (prefix, idx[("upconv", i, 1)]), x, mode)
        out[("upconv", i)] = x
        if i in scales and enable_disparity:
            out[("disp", i)] = torch.sigmoid(_conv3x3_refl(sd, "%sdecoder.%d" % (prefix, idx[("dispconv", i)]), x))
    return out


def pose_decoder(sd, prefix, last_feature, num_frames_to_predict_for=2):
    """PoseDecoder.forward, models/pose_decoder.py:39-58 (num_input_features=1)."""
    out = F.relu(_conv(sd, prefix + "net.0", last_feature))
    out = F.relu(_conv(sd, prefix + "net.1", out, 1, 1))
    out = F.relu(_conv(sd, prefix + "net.2", out, 1, 1))
    out = _conv(sd, prefix + "net.3", out)
    out = 0.01 * out.mean(3).mean(2).view(-1, num_frames_to_predict_for, 1, 6)
    return out[..., :3], out[..., 3:]


def predict_poses(sd, inputs, frame_ids, mode=None):
    """JointSegmentationMonodepth.predict_poses, 'pairs' mode, models/joint_segmentation_depth.py:24-50."""
    out = {}
    for f in frame_ids[1:]:
        if f == "s":
            continue
        pair = [inputs[("color_aug", f, 0)], inputs[("color_aug", 0, 0)]] if f < 0 else \
            [inputs[("color_aug", 0, 0)], inputs[("color_aug", f, 0)]]
        feats = resnet_features(sd, "models.pose_encoder.encoder.", torch.cat(pair, 1), 18, None, mode)
        aa, tr = pose_decoder(sd, "models.pose.", feats[-1])
        out[("axisangle", 0, f)], out[("translation", 0, f)] = aa, tr
        out[("cam_T_cam", 0, f)] = transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0))
    return out


def model_forward(sd, inputs, cfg, mode=None, dropout_mask=None):
    """JointSegmentationMonodepth.forward (models/joint_segmentation_depth.py:77-100) for the monodepth
    configurations (encoder + depth decoder + pose net).  cfg: dict(num_layers, rswd, frame_ids, depth_args)."""
    mode = mode or BNMode()
    feats = resnet_features(sd, "models.encoder.encoder.", inputs[("color_aug", 0, 0)], cfg["num_layers"],
                            cfg.get("rswd"), mode)
    out = {"bottleneck": feats[-1], "features": feats}
    da = cfg.get("depth_args", {})
    out.update(depth_decoder(sd, "models.depth.", feats, mode=mode, dropout_mask=dropout_mask,
                             intermediate_aspp=da.get("intermediate_aspp", True),
                             aspp_rates=da.get("aspp_rates", (6, 12, 18)),
                             aspp_pooling=da.get("aspp_pooling", True)))
    if cfg.get("use_pose_net", True):
        out.update(predict_poses(sd, inputs, cfg["frame_ids"], mode))
    return out


# ------------------------------------------------------------------------------------------------
# deterministic synthetic data / weights shared by the oracle, the tests and bench.py
# ------------------------------------------------------------------------------------------------


def synthetic_inputs(B, H, W, seed=1234, num_scales=4, frame_ids=(0, -1, 1), labels=False):
    """smooth sinusoid textures, source frames = shifted target + noise, Cityscapes
    intrinsics scaled to WxH (loader/cityscapes_loader.py:127-130, sequence_segmentation_loader.py:276-286)."""
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


def synthetic_state_dict(template, seed=0):
    """Fills a {name: tensor} template (e.g. a module's state_dict) with deterministic values:
    fan-in-scaled normal conv weights, BN gamma~U(0.5,1.5), small biases, positive running_var.
    The recipe depends only on (sorted key order, shapes, seed) so any implementation of the same
    architecture gets identical weights without shipping them."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(template.keys()):
        v = template[k]
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = torch.randn(shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            out[k] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            out[k] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif k.endswith(".weight"):      # BN gamma / linear
            out[k] = torch.rand(shape, generator=g) + 0.5 if len(shape) == 1 else \
                torch.randn(shape, generator=g) * 0.01
        else:                            # biases
            out[k] = torch.randn(shape, generator=g) * 0.05
    return out


# ------------------------------------------------------------------------------------------------
# segmentation decoders (joint seg + depth configurations)
# ------------------------------------------------------------------------------------------------


def _self_attention(sd, p, x):
    """SelfAttention.forward, models/model_parts.py:43-45."""
    return _conv(sd, p + ".conv", x, 1, 1) * torch.sigmoid(_conv(sd, p + ".attention", x, 1, 1))


def joint_seg_depth_decoder(sd, prefix, feats, layers=(9,), output_stride=1, head_inter=True, mode=None,
                            dropout_masks=None, depth_args=None):
    """JointSegDepthDecoder.forward, models/joint_segmentation_depth_decoder.py:55-75 (layer_dropout = 0).
    dropout_masks: {"aspp": mask, "head": mask} replayed 0/1 masks (nn.Dropout(0.5) / nn.Dropout(head_dropout=0.1))."""
    mode = mode or BNMode()
    dm = dropout_masks or {}
    da = depth_args or {}
    dec = depth_decoder(sd, prefix + "unet_dec.", feats, mode=mode, dropout_mask=dm.get("aspp"),
                        intermediate_aspp=da.get("intermediate_aspp", True), aspp_rates=da.get("aspp_rates", (6, 12, 18)),
                        aspp_pooling=da.get("aspp_pooling", True))
    last_layer = 9
    seg_size = tuple((feats[last_layer] if last_layer <= 4 else dec[("upconv", 9 - last_layer)]).shape[2:])
    out_size = tuple(int(v) // output_stride for v in seg_size)
    stacked = []
    for layer in layers:
        src = feats[layer] if layer <= 4 else dec[("upconv", 9 - layer)]
        proj = _conv(sd, "%sproject.seg%d.0" % (prefix, layer), src)
        stacked.append(F.interpolate(proj, size=out_size, mode="bilinear", align_corners=False))
    x = torch.cat(stacked, 1)
    if head_inter:      # head = [Identity, conv3x3, BN, ReLU, Dropout, conv1x1] -> indices 1,2,5
        x = F.relu(_bn(sd, prefix + "head.2", _conv(sd, prefix + "head.1", x, 1, 1), mode))
        if mode.training:
            x = x * dm["head"] / 0.9
        x = _conv(sd, prefix + "head.5", x)
    else:               # head = [Identity, Identity, conv1x1]
        x = _conv(sd, prefix + "head.2", x)
    if out_size != seg_size:
        x = F.interpolate(x, size=seg_size, mode="bilinear", align_corners=False)
    return x


def pad_decoder(sd, prefix, feats, distillation_layer=7, final_layer=9, side_output=True, output_stride=1, mode=None,
                dropout_masks=None, depth_args=None):
    """PAD.forward, models/joint_segmentation_depth_decoder.py:134-184.
    dropout_masks: {"depth": aspp mask of depth_dec, "seg": aspp mask of seg_dec}."""
    mode = mode or BNMode()
    dm = dropout_masks or {}
    da = depth_args or {}
    kw = dict(mode=mode, intermediate_aspp=da.get("intermediate_aspp", True), aspp_rates=da.get("aspp_rates", (6, 12, 18)),
              aspp_pooling=da.get("aspp_pooling", True))
    n_up = da.get("n_upconv", 4)
    di = 9 - distillation_layer
    first, second = list(range(n_up, di - 1, -1)), list(range(di - 1, -1, -1))
    depth = depth_decoder(sd, prefix + "depth_dec.", feats, exec_layer=first, dropout_mask=dm.get("depth"), **kw)
    seg = depth_decoder(sd, prefix + "seg_dec.", feats, exec_layer=first, dropout_mask=dm.get("seg"),
                        enable_disparity=False, scales=(), **kw)
    mid = ("upconv", di)
    out = {}
    if side_output:
        inter = _conv(sd, prefix + "seg_intermediate_head.0", seg[mid])
    sa_depth = _self_attention(sd, prefix + "sa_depth", depth[mid])
    sa_seg = _self_attention(sd, prefix + "sa_seg", seg[mid])
    merged_seg, merged_depth = seg[mid] + sa_depth, depth[mid] + sa_seg
    depth.update(depth_decoder(sd, prefix + "depth_dec.", feats, x=merged_depth, exec_layer=second, **kw))
    seg2 = depth_decoder(sd, prefix + "seg_dec.", feats, x=merged_seg, exec_layer=second, enable_disparity=False,
                         scales=(), **kw)
    final = _conv(sd, prefix + "seg_final_head.0", seg2[("upconv", 9 - final_layer)])
    seg_size = tuple(feats[0].shape[2:])
    out_size = tuple(int(v) // output_stride for v in seg_size)
    if out_size != seg_size:
        final = F.interpolate(final, size=seg_size, mode="bilinear", align_corners=False)
        if side_output:
            inter = F.interpolate(inter, size=seg_size, mode="bilinear", align_corners=False)
    out.update(depth)
    out["semantics"] = final
    if side_output:
        out["intermediate_semantics"] = inter
    return out
