"""CPU: which C-ABI entry points the convolution autograd layer (`conv_op.py`) calls for each layer class of the
network, in which order, and with which geometry — checked with a recording stand-in for the library (no kernel
runs; outputs are uninitialised memory).  This pins the host-side plumbing the GPU kernels depend on: halo buffers of
the right size, descriptors with the right padding / stride after each rewrite (dgrad as a forward conv on dy,
zero-stuffing, the `::2` view), the overlapping row-band view of the stem, and that every gradient is produced."""
import ctypes as C

import pytest
import torch


class Recorder:
    """Replaces `_cabi._invoke`: records (name, args), answers 0 (success)."""

    def __init__(self, unsupported=()):
        self.calls, self.unsupported = [], set(unsupported)

    def __call__(self, name, args):
        self.calls.append((name, args))
        return -3 if name in self.unsupported else 0          # SEGSDE_E_UNSUPPORTED

    def names(self):
        return [n for n, _ in self.calls]

    def args_of(self, name, k=0):
        return [a for n, a in self.calls if n == name][k]

    def clear(self):
        self.calls = []


def struct(arg):
    """The ctypes structure behind a byref() argument."""
    return arg._obj


@pytest.fixture()
def rec(monkeypatch):
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import conv_op, ops
    assert A.E_UNSUPPORTED == -3
    r = Recorder()
    monkeypatch.setattr(A, "_invoke", r)
    monkeypatch.setattr(A, "require_cuda", lambda *ts: None)
    monkeypatch.setattr(A, "stream_ptr", lambda: C.c_void_p(0))
    monkeypatch.setattr(conv_op, "_tc_enabled", lambda: True)
    monkeypatch.setattr(ops, "USE_TC", True)
    return r


def cl(*shape, grad=True):
    return torch.randn(*shape).contiguous(memory_format=torch.channels_last).requires_grad_(grad)


def test_decoder_block_with_upsampled_skip(rec):
    """ConvBlock on cat(upsample(x), skip) with reflection padding + ELU (depth_decoder.py:93-101)."""
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
    x, skip = cl(2, 128, 8, 16), cl(2, 64, 16, 32)
    w, b = cl(128, 192, 3, 3), torch.zeros(128, requires_grad=True)
    y = ops.conv2d(x, w, b, x2=skip, pad=1, pad_mode=A.PAD_REFLECT, up1=True, act=A.ACT_ELU)
    assert tuple(y.shape) == (2, 128, 16, 32)
    assert rec.names() == ["segsde_pad_prep", "segsde_pad_prep", "segsde_conv2d_fwd_tc"]
    up_args, skip_args = rec.args_of("segsde_pad_prep", 0), rec.args_of("segsde_pad_prep", 1)
    assert (struct(up_args[1]).h, struct(up_args[1]).w, up_args[2].value) == (18, 34, 1)        # upsampled + halo
    assert (struct(skip_args[1]).h, struct(skip_args[1]).w, skip_args[2].value) == (18, 34, 0)
    fwd = rec.args_of("segsde_conv2d_fwd_tc")
    v1, v2, vy, d = struct(fwd[0]), struct(fwd[1]), struct(fwd[4]), struct(fwd[5])
    assert (v1.c, v2.c, vy.c, vy.h, vy.w) == (128, 64, 128, 16, 32)
    assert (d.kh, d.kw, d.stride, d.pad, d.pad_mode, d.up1, d.act) == (3, 3, 1, 0, A.PAD_ZERO, 0, A.ACT_ELU)
    assert (v1.sw, v1.sh) == (128, 34 * 128) and v1.sn == 18 * 34 * 128                          # dense NHWC halo buffer
    rec.clear()
    y.backward(torch.ones_like(y))
    n = rec.names()
    assert n[0] == "segsde_act_bwd_bias"                                   # ELU' and the bias gradient in one pass
    # one dgrad (forward kernel on dy with flipped weights) + fold per source, one wgrad over both sources
    assert n.count("segsde_weight_transpose_flip") == 2 and n.count("segsde_conv2d_fwd_tc") == 2
    assert n.count("segsde_pad_fold") == 2 and n.count("segsde_conv2d_wgrad_tc") == 1
    dg = struct(rec.args_of("segsde_conv2d_fwd_tc", 0)[5])
    assert (dg.pad, dg.stride, dg.act) == (2, 1, A.ACT_NONE)               # full correlation: pad = k - 1 - 0
    folds = [rec.args_of("segsde_pad_fold", k) for k in range(2)]
    assert sorted(f[2].value for f in folds) == [0, 1]                     # the upsampled source folds 2x2 -> 1
    assert x.grad.shape == x.shape and skip.grad.shape == skip.shape and w.grad.shape == w.shape
    assert w.grad.stride() == w.stride() and b.grad.shape == (128,)


def test_skipless_upconv_runs_as_four_phase_convolutions(rec, monkeypatch):
    """upsample -> ReflectionPad2d(1) -> Conv3x3 without a skip (depth_decoder.py:93-100, the 64 -> 64 layer): four 2x2
    convolutions on windows of the replicate-padded LOW-res input, each writing one [a::2, b::2] phase of y; backward =
    four 2x2 dgrads + one fold, four wgrads + one weight fold.  SEGSDE_PHASE_UPCONV=0 restores pad_prep + 3x3."""
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, conv_op, ops
    x, w, b = cl(2, 64, 8, 32), cl(64, 64, 3, 3), torch.zeros(64, requires_grad=True)
    y = ops.conv2d(x, w, b, pad=1, pad_mode=A.PAD_REFLECT, up1=True, act=A.ACT_ELU)
    assert tuple(y.shape) == (2, 64, 16, 64)
    n = rec.names()
    assert n[0] == "segsde_pad_replicate" and n.count("segsde_weight_phase_up") == 4 and n.count("segsde_conv2d_fwd_tc") == 4
    assert "segsde_pad_prep" not in n
    xp = struct(rec.args_of("segsde_pad_replicate")[1])
    assert (xp.h, xp.w, xp.c) == (10, 34, 64)                               # 1-pixel border on the low-res tensor
    seen = set()
    for k in range(4):
        f = rec.args_of("segsde_conv2d_fwd_tc", k)
        v1, vy, d = struct(f[0]), struct(f[4]), struct(f[5])
        assert (d.kh, d.kw, d.stride, d.pad, d.act) == (2, 2, 1, 0, A.ACT_ELU)
        assert (v1.h, v1.w) == (9, 33) and (vy.h, vy.w, vy.c) == (8, 32, 64)
        assert vy.sw == 2 * 64 and vy.sh == 2 * 64 * 64                     # a stride-2 phase view of the 16 x 64 output
        seen.add(vy.ptr)
    assert len(seen) == 4                                                   # four different phases
    rec.clear()
    y.backward(torch.ones_like(y))
    n = rec.names()
    assert n[0] == "segsde_act_bwd_bias"
    assert n.count("segsde_conv2d_fwd_tc") == 4 and n.count("segsde_phase_up_fold") == 1      # dgrad
    assert n.count("segsde_conv2d_wgrad_tc") == 4 and n.count("segsde_weight_phase_up_fold") == 1
    assert "segsde_pad_fold" not in n
    assert x.grad.shape == x.shape and w.grad.shape == w.shape and w.grad.stride() == w.stride()
    # the switch
    rec.clear()
    monkeypatch.setattr(conv_op, "PHASE_UPCONV", False)
    ops.conv2d(x, w, b, pad=1, pad_mode=A.PAD_REFLECT, up1=True, act=A.ACT_ELU)
    assert rec.names() == ["segsde_pad_prep", "segsde_conv2d_fwd_tc"]


def test_second_backward_over_a_retained_graph_reuses_the_flipped_weights(rec):
    """train.py:486 (`backward(retain_graph=True)`) then :510: the transposed / tap-flipped dgrad weights stay on the
    graph node, so the second traversal launches no second flip."""
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    x, w = cl(2, 64, 8, 32), cl(128, 64, 3, 3)
    y = ops.conv2d(x, w, None, pad=1)
    rec.clear()
    y.sum().backward(retain_graph=True)
    assert rec.names().count("segsde_weight_transpose_flip") == 1
    rec.clear()
    (2 * y).sum().backward()
    n = rec.names()
    assert n.count("segsde_weight_transpose_flip") == 0 and n.count("segsde_conv2d_fwd_tc") == 1      # dgrad still runs


def test_stride2_3x3_backward_rewrites(rec, monkeypatch):
    """3x3 / stride 2 (ResNet downsampling blocks): forward strided through the TMA map; backward: dgrad as four phase
    convolutions (1, 2, 2, 4 taps) of dy writing the (row, column)-parity sub-grids of dx, wgrad directly on the strided
    dy.  With SEGSDE_PHASE_DGRAD off: dgrad on the zero-stuffed dy as a stride-1 problem (round 1's form)."""
    from improving_segmentation_with_selfsupervised_depth_b200 import conv_op, ops
    x, w = cl(2, 64, 16, 32), cl(128, 64, 3, 3)
    y = ops.conv2d(x, w, None, stride=2, pad=1)
    assert tuple(y.shape) == (2, 128, 8, 16) and rec.names() == ["segsde_conv2d_fwd_tc"]
    assert struct(rec.args_of("segsde_conv2d_fwd_tc")[5]).stride == 2
    rec.clear()
    y.backward(torch.ones_like(y), retain_graph=True)
    n = rec.names()
    assert n == ["segsde_weight_phase_s2", "segsde_conv2d_fwd_tc"] * 4 + ["segsde_conv2d_wgrad_tc"]
    taps, subgrids = [], []
    for call in rec.calls:
        if call[0] == "segsde_conv2d_fwd_tc":
            d, out, src = struct(call[1][5]), struct(call[1][4]), struct(call[1][0])
            taps.append((d.kh, d.kw, d.stride, d.pad))
            subgrids.append((out.h, out.w, out.sh, out.sw))
            assert (src.h, src.w, src.c) == (8, 16, 128)                                        # the dense dy
    assert taps == [(1, 1, 1, 0), (1, 2, 1, 0), (2, 1, 1, 0), (2, 2, 1, 0)]
    assert subgrids == [(8, 16, 2 * 32 * 64, 2 * 64)] * 4                                       # [:, a::2, b::2] of dx
    wg = rec.args_of("segsde_conv2d_wgrad_tc")
    assert (struct(wg[2]).h, struct(wg[2]).w, struct(wg[5]).stride) == (8, 16, 2)                # the dense strided dy
    rec.clear()
    monkeypatch.setattr(conv_op, "PHASE_DGRAD_S2", False)
    y.backward(torch.ones_like(y))
    n = rec.names()
    assert n == ["segsde_copy_nhwc", "segsde_weight_transpose_flip", "segsde_conv2d_fwd_tc", "segsde_conv2d_wgrad_tc"]
    stuffed = struct(rec.args_of("segsde_copy_nhwc")[1])
    assert (stuffed.h, stuffed.w) == (8, 16) and stuffed.sw == 2 * 128 and stuffed.sh == 2 * 32 * 128   # ::2 view of 16x32
    dg_in, dg = struct(rec.args_of("segsde_conv2d_fwd_tc")[0]), struct(rec.args_of("segsde_conv2d_fwd_tc")[5])
    assert (dg_in.h, dg_in.w, dg.stride, dg.pad) == (16, 32, 1, 1)


def test_1x1_stride2_runs_on_the_subsampled_view(rec):
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    x, w = cl(2, 64, 16, 32), cl(128, 64, 1, 1)
    y = ops.conv2d(x, w, None, stride=2)
    fwd = rec.args_of("segsde_conv2d_fwd_tc")
    v, d = struct(fwd[0]), struct(fwd[5])
    assert (v.h, v.w, v.sw, v.sh, d.stride) == (8, 16, 128, 2 * 32 * 64, 1)
    rec.clear()
    y.backward(torch.ones_like(y))
    assert x.grad.shape == x.shape and "segsde_conv2d_wgrad_tc" in rec.names()


@pytest.mark.parametrize("c2,P", [(0, 4), (3, 8)])
def test_stem_row_band_view(rec, c2, P):
    """7x7/s2 stem: packed NHWC-P frames and the overlapping view {c = 8P, w = W/2 at stride 2P, h = H + 6}."""
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    H, W = 32, 64
    x1 = torch.rand(2, 3, H, W)
    x2 = torch.rand(2, c2, H, W) if c2 else None
    w = cl(64, 3 + c2, 7, 7)
    sums = torch.zeros(3 * 64, dtype=torch.float64)
    y = ops.conv2d(x1, w, None, x2=x2, stride=2, pad=3, nchw_norm_in=True, bn_stats=sums)
    assert tuple(y.shape) == (2, 64, H // 2, W // 2)
    assert rec.names() == ["segsde_stem_pack", "segsde_stem_pack_w", "segsde_conv2d_fwd_tc_stats"]
    pack = rec.args_of("segsde_stem_pack")
    assert [a.value for a in pack[2:10]] == [3, c2, 2, H, W, 3, W + 8, P]
    fwd = rec.args_of("segsde_conv2d_fwd_tc_stats")
    v, d = struct(fwd[0]), struct(fwd[5])
    assert (v.c, v.w, v.h, v.sw, v.sh, v.sn) == (8 * P, W // 2, H + 6, 2 * P, (W + 8) * P, (H + 6) * (W + 8) * P)
    assert (d.kh, d.kw, d.stride, d.stride_w, d.pad) == (7, 1, 2, 1, 0)
    rec.clear()
    y.backward(torch.ones_like(y))
    assert rec.names() == ["segsde_conv2d_wgrad_tc", "segsde_stem_pack_w"] and w.grad.shape == w.shape
    assert rec.args_of("segsde_stem_pack_w")[7].value == 1                 # direction 1: unpack the packed gradient


def test_stem_falls_back_to_im2col_for_odd_sizes(rec):
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    y = ops.conv2d(torch.rand(1, 3, 31, 64), cl(64, 3, 7, 7), None, stride=2, pad=3, nchw_norm_in=True)
    assert rec.names()[0] == "segsde_stem_im2col" and tuple(y.shape) == (1, 64, 16, 32)


def test_disparity_head_routes(rec, monkeypatch):
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, conv_op, ops
    x, w, b = cl(2, 64, 16, 32), cl(1, 64, 3, 3), torch.zeros(1, requires_grad=True)
    # SEGSDE_HEAD_FEWCOUT=1: 12 tap planes (9 used) through the few-output-channel CUDA-core 1x1 kernels
    monkeypatch.setattr(conv_op, "HEAD_FEWCOUT", True)
    y = ops.conv2d(x, w, b, pad=1, pad_mode=A.PAD_REFLECT, act=A.ACT_SIGMOID)
    assert tuple(y.shape) == (2, 1, 16, 32)
    assert rec.names() == ["segsde_copy_rows", "segsde_conv2d_fwd", "segsde_head_stencil_fwd"]
    assert struct(rec.args_of("segsde_conv2d_fwd")[4]).c == 12
    rec.clear()
    y.backward(torch.ones_like(y))
    n = rec.names()
    assert n[0] == "segsde_act_bwd_bias" and "segsde_head_gcol" in n
    assert n.count("segsde_conv2d_dgrad") == 1 and n.count("segsde_conv2d_wgrad") == 1
    assert not any(k.endswith("_tc") for k in n)
    # default: 32 planes on the tcgen05 kernels
    rec.clear()
    monkeypatch.setattr(conv_op, "HEAD_FEWCOUT", False)
    y = ops.conv2d(x, w, b, pad=1, pad_mode=A.PAD_REFLECT, act=A.ACT_SIGMOID)
    assert rec.names() == ["segsde_copy_rows", "segsde_conv2d_fwd_tc", "segsde_head_stencil_fwd"]       # tap-plane route
    assert struct(rec.args_of("segsde_conv2d_fwd_tc")[4]).c == 32                                       # 9 planes padded to 32
    rec.clear()
    y.backward(torch.ones_like(y))
    n = rec.names()
    assert n[0] == "segsde_act_bwd_bias" and "segsde_head_gcol" in n and n.count("segsde_conv2d_wgrad_tc") == 1


def test_unsupported_tensor_core_shape_takes_the_generic_kernels(monkeypatch):
    """A shape the tensor-core family refuses (SEGSDE_E_UNSUPPORTED) must fall through to the generic entry points —
    never to anything outside the library."""
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import conv_op, ops
    r = Recorder(unsupported={"segsde_conv2d_fwd_tc", "segsde_conv2d_wgrad_tc"})
    monkeypatch.setattr(A, "_invoke", r)
    monkeypatch.setattr(A, "require_cuda", lambda *ts: None)
    monkeypatch.setattr(A, "stream_ptr", lambda: C.c_void_p(0))
    monkeypatch.setattr(conv_op, "_tc_enabled", lambda: True)
    x, w = cl(1, 64, 8, 8), cl(64, 64, 3, 3)
    y = ops.conv2d(x, w, None, pad=1)
    assert r.names() == ["segsde_conv2d_fwd_tc", "segsde_conv2d_fwd"]
    r.clear()
    y.backward(torch.ones_like(y))
    n = r.names()
    # dgrad = the forward entry on dy with flipped weights: tensor-core attempt, then the generic forward kernel
    assert n == ["segsde_weight_transpose_flip", "segsde_conv2d_fwd_tc", "segsde_conv2d_fwd", "segsde_conv2d_wgrad_tc",
                 "segsde_conv2d_wgrad"]


def test_batch_norm_train_uses_the_fused_entry(rec):
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
    x = cl(2, 64, 8, 8)
    g, b = torch.ones(64, requires_grad=True), torch.zeros(64, requires_grad=True)
    rm, rv = torch.zeros(64), torch.ones(64)
    y = ops.batch_norm(x, g, b, rm, rv, True, 0.1, 1e-5, act=A.ACT_RELU)
    assert rec.names() == ["segsde_bn_stats", "segsde_bn_apply_train"]
    rec.clear()
    ops.batch_norm(x, g, b, rm, rv, False, 0.1, 1e-5)
    assert rec.names() == ["segsde_bn_eval_prepare", "segsde_bn_apply"]
    rec.clear()
    y.backward(torch.ones_like(y))
    assert rec.names() == ["segsde_bn_bwd_reduce", "segsde_bn_bwd_apply"]
    assert g.grad.shape == (64,) and x.grad.shape == x.shape
