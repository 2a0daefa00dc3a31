"""GPU: the tcgen05/TMA implicit-GEMM convolution family against the generic CUDA-core kernels (same inputs,
same C-ABI) and against PyTorch fp32 on the CPU.  TF32 operands (10-bit mantissa) with fp32 accumulation:
tolerance 3e-3 relative to the output scale, the numerics class of the reference's own cuDNN path."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = 3e-3


def _mods():
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    return A, ops


def _ref(x1, x2, w, b, pad, dil, reflect, up1, act):
    x = F.interpolate(x1, scale_factor=2, mode="nearest") if up1 else x1
    if x2 is not None:
        x = torch.cat([x, x2], 1)
    if reflect:
        y = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, b, 1, 0, dil)
    else:
        y = F.conv2d(x, w, b, 1, pad, dil)
    return {0: lambda t: t, 1: F.relu, 2: F.elu, 3: torch.sigmoid}[act](y)


CASES = [
    # c1, c2, cout, k, pad, dil, reflect, up1, act, bias, N, H, W
    (64, 0, 64, 1, 0, 1, False, False, 0, False, 2, 16, 32),
    (64, 0, 128, 3, 1, 1, False, False, 1, False, 2, 16, 32),
    (128, 0, 64, 3, 1, 1, False, False, 0, True, 1, 8, 128),
    (32, 0, 64, 3, 2, 2, False, False, 0, False, 2, 12, 40),
    (64, 0, 64, 3, 6, 6, False, False, 1, False, 2, 8, 16),
    (64, 0, 64, 3, 1, 1, True, False, 2, True, 2, 16, 32),
    (64, 32, 128, 3, 1, 1, True, True, 2, True, 2, 8, 16),
    (64, 64, 64, 3, 1, 1, True, False, 2, True, 1, 32, 64),
    (256, 0, 192, 1, 0, 1, False, False, 0, False, 2, 9, 13),
    (96, 0, 256, 3, 1, 1, False, False, 0, False, 1, 20, 160),
    # >= 128 pixels wide: row-halo kernel (2 output rows per tile, shifted-descriptor taps)
    (64, 64, 64, 3, 1, 1, True, False, 2, True, 1, 6, 256),
    (32, 32, 128, 3, 1, 1, True, True, 2, True, 2, 5, 64),
    (64, 0, 64, 3, 2, 2, False, False, 0, False, 1, 7, 160),
    (64, 0, 128, 3, 1, 1, False, False, 0, True, 2, 9, 130),
    # enough tiles for the generic kernel's 256-pixel tiles (>= 2 per SM): 1x1 wide, 3x3 narrow, concat + odd height
    (64, 0, 128, 1, 0, 1, False, False, 0, True, 4, 128, 160),
    (64, 0, 64, 3, 1, 1, False, False, 1, False, 16, 96, 64),
    (64, 64, 128, 1, 0, 1, False, False, 2, True, 5, 127, 144),
    # nearest x2 upsample + reflect pad + 3x3 WITHOUT a skip: four 2x2 phase convolutions on the low-res input
    (64, 0, 64, 3, 1, 1, True, True, 2, True, 2, 16, 32),
    (128, 0, 64, 3, 1, 1, True, True, 0, False, 1, 7, 64),
]
# The kernel every launch of a case must take, in launch order: fprop, dgrad (one entry per source that gets a
# gradient; two sources outside the family share ONE generic launch), wgrad.  "generic" = the CUDA-core kernels: the
# documented fall-back for shapes outside the tensor-core family (dgrad needs Cin % 64 == 0 because Cin is the GEMM N of
# dgrad-as-fprop; wgrad needs Wo % 32 == 0 because pixels are its GEMM K in 32-pixel boxes).
ROUTE = [
    ["tc:conv", "tc:conv", "tc:wgrad"],
    ["tc:conv", "tc:conv", "tc:wgrad3x3"],
    ["tc:rowhalo", "tc:rowhalo", "tc:wgrad3x3"],
    ["tc:conv", "generic", "generic"],                      # Cin = 32; Wo = 40
    ["tc:conv", "tc:conv", "generic"],                      # Wo = 16
    ["tc:conv", "tc:conv", "tc:wgrad3x3"],
    ["tc:conv", "tc:conv", "generic", "tc:wgrad3x3"],       # second source has 32 channels
    ["tc:conv", "tc:conv", "tc:conv", "tc:wgrad3x3"],
    ["tc:conv", "tc:conv", "generic"],                      # Wo = 13
    ["tc:rowhalo", "generic", "tc:wgrad3x3"],               # Cin = 96
    ["tc:rowhalo", "tc:rowhalo", "tc:rowhalo", "tc:wgrad3x3"],
    ["tc:rowhalo", "generic", "tc:wgrad3x3"],               # both sources 32 channels: one generic dgrad launch
    ["tc:rowhalo", "tc:rowhalo", "tc:wgrad3x3"],
    ["tc:rowhalo", "tc:rowhalo", "generic"],                # Wo = 130
    ["tc:conv256", "tc:conv256", "tc:wgrad"],
    ["tc:conv256", "tc:conv256", "tc:wgrad3x3"],
    ["tc:conv256", "tc:conv256", "tc:conv256", "generic"],  # Wo = 144 is not a multiple of 32
    ["tc:conv"] * 4 + ["tc:conv"] * 4 + ["tc:wgrad"] * 4,    # fprop, dgrad, wgrad: one launch per phase
    ["tc:conv"] * 4 + ["tc:conv"] * 4 + ["tc:wgrad"] * 4,
]


def test_tc_is_available():
    A, _ = _mods()
    assert A.lib().segsde_tc_available() == 1


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_tc_conv_fwd_bwd(ci):
    """Every case asserts WHICH kernel ran (ops.ROUTES): a silent fall-back to the fp32 CUDA-core kernels would pass the
    TF32 tolerance trivially."""
    A, ops = _mods()
    ops.USE_TC = True
    case = CASES[ci]
    c1, c2, co, k, pad, dil, reflect, up1, act, bias, N, H, W = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 997)
    x1 = torch.randn(N, c1, H, W, generator=g).requires_grad_()
    H2, W2 = (2 * H, 2 * W) if up1 else (H, W)
    x2 = torch.randn(N, c2, H2, W2, generator=g).requires_grad_() if c2 else None
    w = (torch.randn(co, c1 + c2, k, k, generator=g) / ((c1 + c2) * k * k) ** 0.5).requires_grad_()
    b = torch.randn(co, generator=g).requires_grad_() if bias else None
    y = _ref(x1, x2, w, b, pad, dil, reflect, up1, act)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    gx1 = x1.detach().cuda().requires_grad_()
    gx2 = x2.detach().cuda().requires_grad_() if c2 else None
    gw = w.detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    gb = b.detach().cuda().requires_grad_() if bias else None
    n0 = A.launch_count()
    ops.PROFILE, ops.ROUTES = [], []
    try:
        gy = ops.conv2d(gx1, gw, gb, x2=gx2, stride=1, pad=pad, dil=dil,
                        pad_mode=A.PAD_REFLECT if reflect else A.PAD_ZERO, up1=up1, act=act)
        gy.backward(dy.cuda())
        torch.cuda.synchronize()
        routes = [r for _, r in ops.ROUTES]
    finally:
        ops.PROFILE, ops.ROUTES = None, None
    assert routes == ROUTE[ci], (case, routes)
    assert gy.shape == y.shape
    assert rel_err(gy, y) < TOL, "fprop"

    def close(a, b, what):
        # ReLU's mask is taken from the TF32 output: elements whose fp32 pre-activation is within TF32 rounding
        # of zero flip, which moves isolated gradient entries (allowed: <= 0.5 % of them)
        # flip; each flipped dz entry perturbs every gradient element it feeds by ~1/sqrt(fan) of its scale, so
        # the check is a relative L2 bound instead of an element-wise one
        if act == 1:
            a, b = a.double().cpu(), b.double()
            assert ((a - b).norm() / b.norm()).item() < 2e-2, what
        else:
            assert rel_err(a, b) < TOL, what
    close(gx1.grad, x1.grad, "dgrad x1")
    if c2:
        close(gx2.grad, x2.grad, "dgrad x2")
    close(gw.grad, w.grad, "wgrad")
    if bias:
        close(gb.grad, b.grad, "dbias")
    assert A.launch_count() > n0


@pytest.mark.parametrize("k,cin,cout,H,W", [(3, 64, 128, 32, 64), (1, 64, 128, 32, 64), (3, 128, 256, 18, 36), (1, 256, 512, 16, 32),
                                            (3, 64, 64, 17, 35)])
def test_tc_conv_stride2(k, cin, cout, H, W):
    """ResNet stride-2 convolutions: 3x3 through TMA element strides (fprop) + zero-stuffed dy (dgrad/wgrad),
    1x1 through the ::2 strided view."""
    A, ops = _mods()
    ops.USE_TC = True
    g = torch.Generator().manual_seed(k * 100 + cin)
    x = torch.randn(2, cin, H, W, generator=g).requires_grad_()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).requires_grad_()
    y = F.conv2d(x, w, None, 2, k // 2)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    gx = x.detach().cuda().requires_grad_()
    gw = w.detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    ops.PROFILE = []
    ops.PROFILE_DESC = []
    ops.ROUTES = []
    gy = ops.conv2d(gx, gw, None, stride=2, pad=k // 2)
    gy.backward(dy.cuda())
    torch.cuda.synchronize()
    routes = [r for _, r in ops.ROUTES]
    ops.PROFILE, ops.PROFILE_DESC, ops.ROUTES = None, None, None
    # fprop and dgrad on the tensor cores (3x3: dgrad = four phase convolutions of 1, 2, 2 and 4 taps; 1x1: the ::2 view);
    # wgrad too when the output width is a multiple of its 32-pixel GEMM-K boxes
    n_dgrad = 4 if k == 3 else 1
    assert routes == ["tc:conv"] + ["tc:conv"] * n_dgrad + ["tc:wgrad" if (W // 2) % 32 == 0 else "generic"], routes
    assert gy.shape == y.shape
    assert rel_err(gy, y) < TOL, "fprop"
    assert rel_err(gx.grad, x.grad) < TOL, "dgrad"
    assert rel_err(gw.grad, w.grad) < TOL, "wgrad"


@pytest.mark.parametrize("few", [True, False])
@pytest.mark.parametrize("cin,H,W,reflect", [(64, 16, 32, True), (128, 9, 64, True), (64, 12, 32, False),
                                             (256, 21, 96, True), (64, 2, 32, True), (64, 48, 64, True),
                                             (128, 40, 96, False)])
def test_tc_disparity_head(cin, H, W, reflect, few, monkeypatch):
    """C -> 1 sigmoid heads (fwd, dgrad, wgrad, dbias): tap-plane GEMMs + stencil — on the few-output-channel fp32
    kernels (default; the two larger cases have >= 4096 pixels and reach conv_fewcout.cu) and on the tcgen05 kernels."""
    A, ops = _mods()
    from improving_segmentation_with_selfsupervised_depth_b200 import conv_op
    monkeypatch.setattr(conv_op, "HEAD_FEWCOUT", few)
    ops.USE_TC = True
    g = torch.Generator().manual_seed(cin + H)
    x = torch.randn(2, cin, H, W, generator=g).requires_grad_()
    w = (torch.randn(1, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).requires_grad_()
    b = torch.randn(1, generator=g).requires_grad_()
    y = _ref(x, None, w, b, 1, 1, reflect, False, 3)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    gx = x.detach().cuda().requires_grad_()
    gw = w.detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    gb = b.detach().cuda().requires_grad_()
    ops.PROFILE, ops.PROFILE_DESC = [], []
    gy = ops.conv2d(gx, gw, gb, pad=1, pad_mode=A.PAD_REFLECT if reflect else A.PAD_ZERO, act=A.ACT_SIGMOID)
    gy.backward(dy.cuda())
    torch.cuda.synchronize()
    descs = list(ops.PROFILE_DESC)
    ops.PROFILE, ops.PROFILE_DESC = None, None
    assert any("head" in (d or "") for d in descs), descs
    assert rel_err(gy, y) < TOL
    assert rel_err(gx.grad, x.grad) < TOL
    assert rel_err(gw.grad, w.grad) < TOL
    assert rel_err(gb.grad, b.grad) < TOL


@pytest.mark.parametrize("route", ["band", "im2col"])
@pytest.mark.parametrize("c2,H,W", [(0, 32, 64), (3, 20, 128), (0, 64, 256)])
def test_tc_stem(route, c2, H, W, monkeypatch):
    """7x7 / s2 / pad 3 stem on NCHW frames with the (x-0.45)/0.225 normalisation (resnet_encoder.py:92-93): the
    row-band route (overlapping TMA view of the packed frames, no im2col) and the im2col fallback, forward, fused
    BatchNorm statistics and weight gradient against PyTorch fp32."""
    A, ops = _mods()
    monkeypatch.setenv("SEGSDE_STEM_BAND", "1" if route == "band" else "0")
    ops.USE_TC = True
    try:
        g = torch.Generator().manual_seed(11 + c2 + H)
        n = 2
        x1 = torch.rand(n, 3, H, W, generator=g)
        x2 = torch.rand(n, c2, H, W, generator=g) if c2 else None
        w = torch.randn(64, 3 + c2, 7, 7, generator=g) * 0.05
        xin = torch.cat([x1, x2], 1) if c2 else x1
        wr = w.clone().requires_grad_()
        ref = F.conv2d((xin - 0.45) / 0.225, wr, None, 2, 3)
        gy = torch.randn(ref.shape, generator=g)
        ref.backward(gy)
        wt = w.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
        sums = torch.zeros(3 * 64, device="cuda", dtype=torch.float64)
        ops.PROFILE, ops.PROFILE_DESC = [], []
        y = ops.conv2d(x1.cuda(), wt, None, x2=x2.cuda() if c2 else None, stride=2, pad=3, nchw_norm_in=True, bn_stats=sums)
        y.backward(gy.cuda())
        assert any(route in d for d in ops.PROFILE_DESC), ops.PROFILE_DESC
        assert rel_err(y, ref) < TOL
        assert rel_err(wt.grad, wr.grad) < TOL
        cnt = ref.numel() / 64
        mean = ref.detach().double().mean((0, 2, 3))
        # sums = (sum(y - shift), sum((y - shift)^2), shift) per channel
        s = sums.cpu().view(3, 64)
        assert rel_err(s[0] / cnt + s[2], mean) < TOL
        var = ref.detach().double().var((0, 2, 3), unbiased=False)
        assert rel_err(s[1] / cnt - (s[0] / cnt) ** 2, var) < 2 * TOL
    finally:
        ops.PROFILE, ops.PROFILE_DESC = None, None
        ops.USE_TC = False
