"""GPU: every network-path kernel against plain PyTorch fp32 on the CPU (the same ATen ops the reference calls).
fp32 CUDA-core kernels: 1e-4 relative (summation-order differences only)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = 2e-4


def _ops():
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    ops.USE_TC = False
    return A, ops


def _ref_conv(x1, x2, w, b, stride, pad, dil, reflect, up1, act):
    x = F.interpolate(x1, scale_factor=2, mode="nearest") if up1 else x1
    if x2 is not None:
        x = torch.cat([x, x2], 1)
    if reflect:
        x = F.pad(x, (pad,) * 4, mode="reflect")
        y = F.conv2d(x, w, b, stride, 0, dil)
    else:
        y = F.conv2d(x, w, b, stride, pad, dil)
    return {0: lambda t: t, 1: F.relu, 2: F.elu, 3: torch.sigmoid}[act](y)


CASES = [
    # cin1, cin2, cout, k, stride, pad, dil, reflect, up1, act, bias, H, W
    (16, 0, 24, 3, 1, 1, 1, False, False, 1, False, 13, 17),
    (8, 0, 8, 1, 1, 0, 1, False, False, 0, False, 9, 11),
    (8, 0, 16, 1, 2, 0, 1, False, False, 0, False, 10, 12),
    (8, 0, 12, 3, 2, 1, 1, False, False, 0, True, 11, 14),
    (8, 0, 8, 3, 1, 2, 2, False, False, 0, False, 12, 12),
    (8, 0, 8, 3, 1, 6, 6, False, False, 1, False, 8, 16),
    (16, 0, 8, 3, 1, 1, 1, True, False, 2, True, 10, 14),
    (12, 20, 16, 3, 1, 1, 1, True, True, 2, True, 6, 9),
    (12, 20, 16, 3, 1, 1, 1, True, False, 2, True, 12, 18),
    (16, 0, 1, 3, 1, 1, 1, True, False, 3, True, 16, 24),
    (32, 0, 12, 1, 1, 0, 1, False, False, 0, True, 4, 8),
    (70, 0, 66, 3, 1, 1, 1, False, False, 0, False, 9, 9),
    # 1x1 conv over a handful of pixels (ASPP image pooling, model_parts.py:28-40): warp-per-output kernel
    (96, 0, 40, 1, 1, 0, 1, False, False, 1, True, 1, 1),
    (200, 0, 24, 1, 1, 0, 1, False, False, 0, False, 2, 3),
    # pose head shape class (pose_decoder.py:33): few output channels, many pixels -> finely split wgrad
    (64, 0, 12, 1, 1, 0, 1, False, False, 0, True, 16, 32),
    # segmentation-head shape class (joint_segmentation_depth_decoder.py:106-107): 1x1, Cout = 19 classes on >= 4096
    # pixels -> the few-output-channel kernels of conv_fewcout.cu (fprop / dgrad / wgrad); partial channel chunks too
    (64, 0, 19, 1, 1, 0, 1, False, False, 0, True, 48, 64),
    (128, 0, 19, 1, 1, 0, 1, False, False, 0, False, 40, 60),
    (36, 0, 5, 1, 1, 0, 1, False, False, 2, True, 50, 50),
    (256, 0, 12, 1, 1, 0, 1, False, False, 0, True, 48, 48),
    (96, 0, 32, 1, 1, 0, 1, False, False, 1, False, 47, 53),
]


@pytest.mark.parametrize("case", CASES)
def test_conv_fwd_bwd(case):
    A, ops = _ops()
    c1, c2, co, k, stride, pad, dil, reflect, up1, act, bias, H, W = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x1 = torch.randn(2, c1, H, W, generator=g).requires_grad_()
    H2, W2 = (2 * H, 2 * W) if up1 else (H, W)
    x2 = torch.randn(2, c2, H2, W2, generator=g).requires_grad_() if c2 else None
    w = (torch.randn(co, c1 + c2, k, k, generator=g) * 0.1).requires_grad_()
    b = torch.randn(co, generator=g).requires_grad_() if bias else None
    y = _ref_conv(x1, x2, w, b, stride, pad, dil, reflect, up1, act)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    gx1 = x1.detach().cuda().requires_grad_()
    gx2 = x2.detach().cuda().requires_grad_() if c2 else None
    gw = w.detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    gb = b.detach().cuda().requires_grad_() if bias else None
    gy = ops.conv2d(gx1, gw, gb, x2=gx2, stride=stride, pad=pad, dil=dil,
                    pad_mode=A.PAD_REFLECT if reflect else A.PAD_ZERO, up1=up1, act=act)
    assert gy.shape == y.shape
    assert rel_err(gy, y) < TOL
    gy.backward(dy.cuda())
    assert rel_err(gx1.grad, x1.grad) < TOL
    if c2:
        assert rel_err(gx2.grad, x2.grad) < TOL
    assert rel_err(gw.grad, w.grad) < TOL
    if bias:
        assert rel_err(gb.grad, b.grad) < TOL


@pytest.mark.parametrize("cin", [3, 6])
def test_stem_conv_nchw_normalised(cin):
    A, ops = _ops()
    g = torch.Generator().manual_seed(cin)
    x = torch.rand(2, cin, 20, 28, generator=g)
    w = (torch.randn(64, cin, 7, 7, generator=g) * 0.1).requires_grad_()
    y = F.conv2d((x - 0.45) / 0.225, w, None, 2, 3)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    gw = w.detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    if cin == 3:
        gy = ops.conv2d(x.cuda(), gw, stride=2, pad=3, nchw_norm_in=True)
    else:
        gy = ops.conv2d(x[:, :3].contiguous().cuda(), gw, x2=x[:, 3:].contiguous().cuda(), stride=2, pad=3, nchw_norm_in=True)
    assert rel_err(gy, y) < TOL
    gy.backward(dy.cuda())
    assert rel_err(gw.grad, w.grad) < TOL


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("with_res", [False, True])
def test_batch_norm(training, with_res):
    A, ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(3, 20, 7, 9, generator=g) * 2 + 0.5).requires_grad_()
    res = torch.randn(3, 20, 7, 9, generator=g).requires_grad_() if with_res else None
    wt, bs = (torch.rand(20, generator=g) + 0.5).requires_grad_(), torch.randn(20, generator=g).requires_grad_()
    rm, rv = torch.randn(20, generator=g) * 0.1, torch.rand(20, generator=g) + 0.5
    rm2, rv2 = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm2, rv2, wt, bs, training, 0.1, 1e-5)
    y = F.relu(y + res) if with_res else F.relu(y)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    gx, gwt, gbs = x.detach().cuda().requires_grad_(), wt.detach().cuda().requires_grad_(), bs.detach().cuda().requires_grad_()
    gres = res.detach().cuda().requires_grad_() if with_res else None
    grm, grv = rm.cuda(), rv.cuda()
    gy = ops.batch_norm(gx, gwt, gbs, grm, grv, training, 0.1, 1e-5, residual=gres, act=A.ACT_RELU)
    assert rel_err(gy, y) < 1e-5
    assert rel_err(grm, rm2) < 1e-5 and rel_err(grv, rv2) < 1e-5
    gy.backward(dy.cuda())
    assert rel_err(gx.grad, x.grad) < 1e-4
    assert rel_err(gwt.grad, wt.grad) < 1e-4 and rel_err(gbs.grad, bs.grad) < 1e-4
    if with_res:
        assert rel_err(gres.grad, res.grad) < 1e-6


def test_pool_resize_misc():
    A, ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 12, 11, 14, generator=g).requires_grad_()
    # maxpool
    y = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    gx = x.detach().cuda().requires_grad_()
    gy = ops.maxpool3x3s2(gx)
    assert rel_err(gy, y) == 0
    gy.backward(dy.cuda())
    assert rel_err(gx.grad, x.grad) < 1e-6
    # spatial mean + broadcast
    x.grad = None
    y = 0.01 * x.mean(3).mean(2)
    y.sum().backward()
    gx = x.detach().cuda().requires_grad_()
    gy = ops.spatial_mean(gx, 0.01)
    assert rel_err(gy.reshape(2, 12), y) < 1e-5
    gy.sum().backward()
    assert rel_err(gx.grad, x.grad) < 1e-5
    v = torch.randn(2, 12, 1, 1, generator=g).requires_grad_()
    yb = F.interpolate(v, size=(5, 7), mode="bilinear", align_corners=False)
    dyb = torch.randn(yb.shape, generator=g)
    yb.backward(dyb)
    gv = v.detach().cuda().requires_grad_()
    gyb = ops.broadcast_hw(gv, 5, 7)
    assert rel_err(gyb, yb) < 1e-6
    gyb.backward(dyb.cuda())
    assert rel_err(gv.grad, v.grad) < 1e-5
    # bilinear, both conventions, up and down
    for size, align in (((22, 28), False), ((22, 28), True), ((5, 9), False), ((33, 14), True)):
        x.grad = None
        yr = F.interpolate(x, size=size, mode="bilinear", align_corners=align)
        dyr = torch.randn(yr.shape, generator=g)
        yr.backward(dyr)
        gx = x.detach().cuda().requires_grad_()
        gyr = ops.bilinear(gx, size, align)
        assert rel_err(gyr, yr) < 1e-5, (size, align)
        gyr.backward(dyr.cuda())
        assert rel_err(gx.grad, x.grad) < 1e-5, (size, align)
    # cat / add / gate / activation
    a, b = torch.randn(2, 5, 4, 6, generator=g).requires_grad_(), torch.randn(2, 7, 4, 6, generator=g).requires_grad_()
    yc = torch.cat([a, b], 1)
    wc = torch.randn(yc.shape, generator=g)
    (yc * wc).sum().backward()
    ga, gb2 = a.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
    gyc = ops.cat_channels([ga, gb2])
    assert rel_err(gyc, yc) == 0
    (gyc * wc.cuda()).sum().backward()
    assert rel_err(ga.grad, a.grad) == 0 and rel_err(gb2.grad, b.grad) == 0
    f, t = torch.randn(2, 8, 4, 6, generator=g).requires_grad_(), torch.randn(2, 8, 4, 6, generator=g).requires_grad_()
    yg = f * torch.sigmoid(t)
    wg = torch.randn(yg.shape, generator=g)
    (yg * wg).sum().backward()
    gf, gt = f.detach().cuda().requires_grad_(), t.detach().cuda().requires_grad_()
    gyg = ops.gate(gf, gt)
    assert rel_err(gyg, yg) < 1e-6
    (gyg * wg.cuda()).sum().backward()
    assert rel_err(gf.grad, f.grad) < 1e-5 and rel_err(gt.grad, t.grad) < 1e-5
    assert rel_err(ops.add(gf, gt), f + t) < 1e-7
    assert rel_err(ops.activation(gf, A.ACT_ELU), F.elu(f)) < 1e-6
    # dropout: replayed mask and statistical check of the Philox mask
    m = (torch.rand(2, 8, 4, 6, generator=g) > 0.5).float()
    yd = ops.dropout(gf, 0.5, True, replay_mask=m)
    assert rel_err(yd, f * m / 0.5) < 1e-7
    big = torch.ones(4, 64, 32, 32).cuda()
    kept = (ops.dropout(big, 0.3, True) != 0).float().mean().item()
    assert abs(kept - 0.7) < 0.01
    # NCHW-contiguous input goes through the layout kernel
    xc = torch.randn(2, 6, 5, 7, generator=g)
    assert rel_err(ops.as_cl(xc.cuda()), xc) == 0
