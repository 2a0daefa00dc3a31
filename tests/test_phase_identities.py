"""CPU: the two algebraic rewrites the convolution dispatch relies on, restated in plain PyTorch and
checked against the ops they replace.  The CUDA kernels that build the phase weights (`segsde_weight_phase_up`,
`segsde_weight_phase_s2`) and fold the results are checked against PyTorch in the GPU suite; this file pins the
identities themselves, tap sets included, so that a change of either side shows up without a GPU."""
import pytest
import torch
import torch.nn.functional as F

R_UP = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}       # taps of phase a summed into low-res tap d


def phase_up_weights(w, a, b):
    """[Cout, Cin, 2, 2] weights of output phase (a, b): wp[dr, ds] = sum_{r in R(a,dr)} sum_{s in R(b,ds)} w[r, s]."""
    wp = w.new_zeros(w.shape[0], w.shape[1], 2, 2)
    for dr in range(2):
        for ds in range(2):
            for r in R_UP[(a, dr)]:
                for s in R_UP[(b, ds)]:
                    wp[:, :, dr, ds] += w[:, :, r, s]
    return wp


@pytest.mark.parametrize("h,w_", [(2, 3), (5, 8), (7, 4)])
def test_upsample_reflectpad_conv3x3_is_four_2x2_convs_on_the_replicate_padded_input(h, w_):
    """y = conv3x3(reflect_pad1(nearest_up2(x))) (depth_decoder.py:93-100) == for each output phase (a, b) a 2x2
    convolution of the replicate-padded low-res x over the window starting at (a, b)."""
    g = torch.Generator().manual_seed(h * 10 + w_)
    x = torch.randn(2, 5, h, w_, generator=g, dtype=torch.float64)
    w = torch.randn(4, 5, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.pad(F.interpolate(x, scale_factor=2, mode="nearest"), (1, 1, 1, 1), mode="reflect"), w)
    xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
    y = torch.empty_like(ref)
    for a in range(2):
        for b in range(2):
            y[:, :, a::2, b::2] = F.conv2d(xp[:, :, a:a + h + 1, b:b + w_ + 1], phase_up_weights(w, a, b))
    torch.testing.assert_close(y, ref, rtol=1e-12, atol=1e-12)


def phase_s2_weights(wt, a, b):
    """Taps of dx phase (a, b) for a 3x3 / stride-2 / pad-1 convolution: rows a = 0 -> {w[1] at dy offset 0};
    a = 1 -> {w[2] at offset 0, w[0] at offset +1}; columns alike.  Returned as a correlation kernel over dy."""
    rows = [(1,)] if a == 0 else [(2,), (0,)]
    cols = [(1,)] if b == 0 else [(2,), (0,)]
    k = wt.new_zeros(wt.shape[1], wt.shape[0], len(rows), len(cols))       # [Cin, Cout, th, tw]
    for i, (r,) in enumerate(rows):
        for j, (s,) in enumerate(cols):
            k[:, :, i, j] = wt[:, :, r, s].t()
    return k


@pytest.mark.parametrize("H,W", [(8, 12), (6, 6)])
def test_stride2_3x3_dgrad_is_four_phase_convolutions_of_dy(H, W):
    """dx of conv3x3/s2/p1 == for each phase (a, b) of dx a (1+a) x (1+b)-tap stride-1 correlation of dy (zero beyond the
    last row / column) - 9/4 of the useful MACs instead of the 4x of zero-stuffing."""
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(2, 3, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(4, 3, 3, 3, generator=g, dtype=torch.float64)
    y = F.conv2d(x, w, stride=2, padding=1)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    dx = torch.zeros_like(x)
    for a in range(2):
        for b in range(2):
            k = phase_s2_weights(w, a, b)
            dyp = F.pad(dy, (0, b, 0, a))                  # the +1 taps read one row / column past the end: zeros
            dx[:, :, a::2, b::2] = F.conv2d(dyp, k)[:, :, :dx[:, :, a::2, b::2].shape[2], :dx[:, :, a::2, b::2].shape[3]]
    torch.testing.assert_close(dx, x.grad, rtol=1e-12, atol=1e-12)
