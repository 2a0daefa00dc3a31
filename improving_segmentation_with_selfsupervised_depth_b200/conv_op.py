"""Convolution autograd op: routes every eligible conv to the tcgen05/TMA kernels (conv_tc.cu) and the
rest (3/6-channel stem, C->1 disparity heads, 12-channel pose output) to the generic CUDA-core kernels
(conv_simt.cu / conv_small.cu).

Tensor-core route, by input-side fusion:
  * reflection padding / nearest x2 upsampling (decoder): materialised once into a halo-padded NHWC buffer
    (`segsde_pad_prep`), after which the conv is a plain "valid" one whose taps are TMA boxes at shifted
    coordinates; the concat stays virtual (second tensor map).  The halo buffer's gradient is folded back with
    `segsde_pad_fold`; wgrad reads the saved padded buffers.
  * stride 2, 1x1 (ResNet downsample): a stride-1 conv on the ::2 strided view (the TMA map just gets doubled
    pixel strides); its dgrad writes through the same strided view of a zero-filled gradient.
  * stride 2, 3x3: fprop through TMA element strides; for dgrad/wgrad the output gradient is zero-stuffed to
    the stride-1 output grid, which turns both into the stride-1 kernels.
  * dgrad everywhere = the fprop kernel on dy with the transposed / tap-flipped weights.
"""
import ctypes as C
import os

import torch

from . import _cabi as A
from . import ops

TC_STRIDE2_FPROP = True      # 3x3/s2 forward on the tensor cores via TMA element strides

# stride-2 3x3 dgrad as four phase convolutions instead of zero-stuffing (SEGSDE_PHASE_DGRAD=0 restores round 1's form)
PHASE_DGRAD_S2 = os.environ.get("SEGSDE_PHASE_DGRAD", "1") != "0"


# nearest x2 upsample + reflect pad + 3x3 conv (no skip) as four 2x2 phase convolutions on the low-res input
# (SEGSDE_PHASE_UPCONV=0 restores the materialised form)
# disparity heads: 32-plane tcgen05 GEMMs (default; measured 1.99 ms fwd+bwd at 12x512x1024) or 12-plane fp32 CUDA-core
# GEMMs of conv_fewcout.cu (SEGSDE_HEAD_FEWCOUT=1; 3.06 ms - latency-bound, kept as the fp32 cross-check of the heads)
HEAD_FEWCOUT = os.environ.get("SEGSDE_HEAD_FEWCOUT", "0") == "1"
PHASE_UPCONV = os.environ.get("SEGSDE_PHASE_UPCONV", "1") != "0"


def _tc_enabled():
    return ops.USE_TC and A.lib().segsde_tc_available() == 1


def _prep(x, up, pad):
    n, c, h, w = x.shape
    f = 2 if up else 1
    y = ops.cl_empty(n, c, h * f + 2 * pad, w * f + 2 * pad, x.device)
    A.call("segsde_pad_prep", C.byref(ops.view(x)), C.byref(ops.view(y)), C.c_int(int(up)), C.c_int(pad),
           A.stream_ptr())
    return y


def _fwd(x1, x2, w, b, y, d, kind, flops, desc=None, stats=None):
    """stats: optional [tensor(3C, float64, zero-filled), done-flag] box — BatchNorm batch statistics of y are
    accumulated by the tensor-core epilogue when that route is taken (done-flag set), else left to the caller."""
    st = A.stream_ptr()
    v1, v2, vy = ops.view(x1), (ops.view(x2) if x2 is not None else None), ops.view(y)
    if d.nchw_norm_in:
        v1.sn = v1.sh = v1.sw = 0

    def launch():
        if _tc_enabled():
            if stats is not None and A.try_call("segsde_conv2d_fwd_tc_stats", C.byref(v1), ops._ref(v2), A.ptr(w), A.ptr(b),
                                                C.byref(vy), C.byref(d), A.ptr(stats[0]), st):
                stats[1] = True
                return ops.log_route(kind, True)
            if A.try_call("segsde_conv2d_fwd_tc", C.byref(v1), ops._ref(v2), A.ptr(w), A.ptr(b), C.byref(vy), C.byref(d), st):
                return ops.log_route(kind, True)
        A.call("segsde_conv2d_fwd", C.byref(v1), ops._ref(v2), A.ptr(w), A.ptr(b), C.byref(vy), C.byref(d), st)
        ops.log_route(kind, False)
    ops._timed(kind, flops, launch, desc)


class _Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, weight, bias, stride, pad, dil, pad_mode, up1, act, nchw, stats=None):
        A.require_cuda(x1, weight)
        if nchw:
            x1 = x1.contiguous()
            x2 = x2.contiguous() if x2 is not None else None
        else:
            x1 = ops.as_cl(x1)
            x2 = ops.as_cl(x2) if x2 is not None else None
        w = ops.ohwi(weight.detach())
        cout, ctot, kh, kw = w.shape
        full_shape1 = tuple(x1.shape)
        c1 = x1.shape[1]
        c2 = x2.shape[1] if x2 is not None else 0
        tc_ch = (not nchw and c1 % 32 == 0 and c2 % 32 == 0 and cout % 64 == 0 and _tc_enabled())
        sub2 = False
        if stride == 2 and kh == 1 and kw == 1 and pad == 0 and x2 is None and tc_ch:
            x1 = x1[:, :, ::2, ::2]          # 1x1 / stride 2 == 1x1 / stride 1 on the subsampled view
            stride, sub2 = 1, True
        n, _, h1, w1 = x1.shape
        hc, wc = (h1 * 2, w1 * 2) if up1 else (h1, w1)
        ho, wo = ops._conv_out_hw(hc, wc, kh, kw, stride, pad, dil)
        reflect = pad_mode == A.PAD_REFLECT
        prepped = tc_ch and stride == 1 and (reflect or up1) and pad > 0
        if prepped:          # materialise padding (+ upsampling); the conv becomes a plain valid one
            x1e = _prep(x1, up1, pad)
            x2e = _prep(x2, False, pad) if x2 is not None else None
            pad_e, mode_e, up_e = 0, A.PAD_ZERO, False
        else:
            x1e, x2e, pad_e, mode_e, up_e = x1, x2, pad, pad_mode, up1
        y = ops.cl_empty(n, cout, ho, wo, x1.device)
        d = ops._desc(kh, kw, stride, pad_e, dil, mode_e, up_e, act, nchw)
        b = bias.detach() if bias is not None else None
        desc = "%d+%d->%d k%d s%d d%d out %dx%d%s%s" % (c1, c2, cout, kh, stride, dil, ho, wo,
                                                        " prep" if prepped else "", " sub2" if sub2 else "")
        _fwd(x1e, x2e, w, b, y, d, "fprop", 2.0 * n * ho * wo * cout * kh * kw * ctot, desc, stats)
        ctx.save_for_backward(x1e, x2e, w, y if act != A.ACT_NONE else None)
        ctx.cfg = (stride, pad, dil, pad_mode, up1, act, nchw, bias is not None, prepped, pad_e, mode_e, up_e,
                   tuple(x1.shape), tuple(x2.shape) if x2 is not None else None, sub2, full_shape1, tc_ch)
        ctx.desc = desc
        return y

    @staticmethod
    def backward(ctx, dy):
        x1e, x2e, w, y = ctx.saved_tensors
        (stride, pad, dil, pad_mode, up1, act, nchw, has_bias, prepped, pad_e, mode_e, up_e, shp1, shp2, sub2,
         full_shape1, tc_ch) = ctx.cfg
        cout, ctot, kh, kw = w.shape
        st = A.stream_ptr()
        dy = ops.as_cl(dy)
        need1, need2, needw, needb = ctx.needs_input_grad[:4]
        dev = dy.device
        db = ops.zeros_f32(cout, dev) if (has_bias and needb) else None
        if act != A.ACT_NONE or db is not None:
            dz = ops.cl_empty(*dy.shape, dev) if act != A.ACT_NONE else None
            A.call("segsde_act_bwd_bias", ops._ref(ops.view(y)) if y is not None else None, C.byref(ops.view(dy)),
                   ops._ref(ops.view(dz)) if dz is not None else None, C.c_int(act), A.ptr(db), st)
            if dz is None:
                dz = dy
        else:
            dz = dy
        n, _, ho, wo = dz.shape
        c1 = x1e.shape[1]
        c2 = x2e.shape[1] if x2e is not None else 0
        flops_scale = 1.0
        dz_s, d_s = None, None      # the strided dz / descriptor: wgrad takes them directly (no zero-stuffing)
        phase_dx1 = None
        if (stride == 2 and tc_ch and mode_e == A.PAD_ZERO and not up_e and not nchw and need1 and x2e is None and kh == 3
                and kw == 3 and pad_e == 1 and dil == 1 and c1 % 64 == 0 and cout % 32 == 0 and wo >= 8 and _tc_enabled()
                and PHASE_DGRAD_S2):
            # dgrad of a 3x3 / stride-2 convolution as four phase convolutions (1, 2, 2 and 4 taps) of dz, each writing
            # one (row parity, column parity) sub-grid of dx: 9/4 taps per input pixel instead of 9 over the zero-stuffed
            # gradient, and no fill + strided copy of the stuffed buffer
            gx = ops.cl_empty(*x1e.shape, dev)
            ok = True
            for a in (0, 1):
                for b_ in (0, 1):
                    sub = gx[:, :, a::2, b_::2]
                    if sub.shape[2] == 0 or sub.shape[3] == 0:
                        continue
                    wt = torch.empty(c1 * (1 + a) * (1 + b_) * cout, device=dev, dtype=torch.float32)
                    A.call("segsde_weight_phase_s2", A.ptr(w), A.ptr(wt), C.c_int(cout), C.c_int(ctot), C.c_int(0), C.c_int(c1),
                           C.c_int(a), C.c_int(b_), st)
                    dd = ops._desc(1 + a, 1 + b_, 1, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False)
                    fl = 2.0 * n * sub.shape[2] * sub.shape[3] * cout * (1 + a) * (1 + b_) * c1
                    done = []

                    def launch(sub=sub, wt=wt, dd=dd):
                        if A.try_call("segsde_conv2d_fwd_tc", C.byref(ops.view(dz)), None, A.ptr(wt), None,
                                      C.byref(ops.view(sub)), C.byref(dd), st):
                            done.append(1)
                            ops.log_route("dgrad", True)
                    ops._timed("dgrad", fl, launch, ctx.desc + " phase%d%d" % (a, b_))
                    ok = ok and bool(done)
            if ok:
                phase_dx1 = gx
            need1 = need1 and not ok           # a refused phase: fall back to the zero-stuffed form below
        if stride == 2 and tc_ch and mode_e == A.PAD_ZERO and not up_e and not nchw and (need1 or need2 or needw):
            dz_s, d_s = dz, ops._desc(kh, kw, 2, pad_e, dil, mode_e, up_e, A.ACT_NONE, nchw)
            if need1 or need2:
                # zero-stuff dz onto the stride-1 output grid: dgrad (and a wgrad the strided form refuses) become
                # stride-1 problems
                hs = x1e.shape[2] + 2 * pad_e - dil * (kh - 1)
                ws = x1e.shape[3] + 2 * pad_e - dil * (kw - 1)
                dzu = ops.cl_empty(n, cout, hs, ws, dev, zero=True)
                A.call("segsde_copy_nhwc", C.byref(ops.view(dz)), C.byref(ops.view(dzu[:, :, ::2, ::2][:, :, :ho, :wo])), st)
                dz, stride, flops_scale = dzu, 1, 0.25        # algorithmic flops stay those of the strided conv
        nz, _, hz, wz = dz.shape
        d = ops._desc(kh, kw, stride, pad_e, dil, mode_e, up_e, A.ACT_NONE, nchw)
        dx1 = dx2 = dw = None

        # ---- dgrad ---------------------------------------------------------------------------------
        if (need1 or (need2 and x2e is not None)) and not nchw:
            tc_dgrad = (_tc_enabled() and stride == 1 and mode_e == A.PAD_ZERO and not up_e and cout % 32 == 0
                        and dil * (kh - 1) - pad_e >= 0)
            results = []
            full1 = None
            for idx, (need, xe, c0, cn) in enumerate(((need1, x1e, 0, c1), (need2 and x2e is not None, x2e, c1, c2))):
                if not need:
                    results.append(None)
                    continue
                if tc_dgrad and cn % 64 == 0:
                    # the transposed / tap-flipped weights are kept on the graph node: the second backward over a
                    # retained graph (train.py:486 then :510) reuses them; they die with the graph
                    cache = ctx.__dict__.setdefault("_wt_cache", {})
                    wt = cache.get((c0, cn))
                    if wt is None:
                        wt = torch.empty(cn * kh * kw * cout, device=dev, dtype=torch.float32)
                        A.call("segsde_weight_transpose_flip", A.ptr(w), A.ptr(wt), C.c_int(cout), C.c_int(kh),
                               C.c_int(kw), C.c_int(ctot), C.c_int(c0), C.c_int(cn), st)
                        cache[(c0, cn)] = wt
                    if sub2 and idx == 0:      # gradient of the ::2 view: written through the view, zeros elsewhere
                        full1 = ops.cl_empty(*full_shape1, dev, zero=True)
                        gx = full1[:, :, ::2, ::2]
                    else:
                        gx = ops.cl_empty(*xe.shape, dev)
                    dd = ops._desc(kh, kw, 1, dil * (kh - 1) - pad_e, dil, A.PAD_ZERO, False, A.ACT_NONE, False)
                    _fwd(dz, None, wt, None, gx, dd, "dgrad", flops_scale * 2.0 * nz * hz * wz * cout * kh * kw * cn, ctx.desc)
                    results.append(gx)
                else:
                    results.append("generic")
            if "generic" in results:
                folded = mode_e == A.PAD_REFLECT or up_e
                g1 = ops.cl_empty(*x1e.shape, dev, zero=folded) if results[0] == "generic" else None
                g2 = ops.cl_empty(*x2e.shape, dev, zero=folded) if (x2e is not None and results[1] == "generic") else None
                v1 = ops.view(g1) if g1 is not None else ops.view(x1e, null=True)
                v2 = (ops.view(g2) if g2 is not None else ops.view(x2e, null=True)) if x2e is not None else None
                cneed = (c1 if g1 is not None else 0) + (c2 if g2 is not None else 0)
                ops._timed("dgrad", flops_scale * 2.0 * nz * hz * wz * cout * kh * kw * cneed,
                           lambda: A.call("segsde_conv2d_dgrad", C.byref(ops.view(dz)), A.ptr(w), C.byref(v1),
                                          ops._ref(v2), C.byref(d), st), ctx.desc + " generic")
                ops.log_route("dgrad", False)
                if g1 is not None:
                    results[0] = g1
                    if sub2:
                        full1 = ops.cl_empty(*full_shape1, dev, zero=True)
                        A.call("segsde_copy_nhwc", C.byref(ops.view(g1)), C.byref(ops.view(full1[:, :, ::2, ::2])), st)
                if g2 is not None:
                    results[1] = g2
            dx1, dx2 = results[0], results[1]
            if sub2 and dx1 is not None:
                dx1 = full1
            if prepped:          # fold the halo / upsampling back onto the original tensors
                if dx1 is not None:
                    f = ops.cl_empty(*shp1, dev)
                    A.call("segsde_pad_fold", C.byref(ops.view(dx1)), C.byref(ops.view(f)), C.c_int(int(up1)),
                           C.c_int(pad), st)
                    dx1 = f
                if dx2 is not None:
                    f = ops.cl_empty(*shp2, dev)
                    A.call("segsde_pad_fold", C.byref(ops.view(dx2)), C.byref(ops.view(f)), C.c_int(0), C.c_int(pad), st)
                    dx2 = f

        # ---- wgrad --------------------------------------------------------------------------------------
        if needw:
            dw = ops.zeros_like_w(w)
            v1, v2 = ops.view(x1e), (ops.view(x2e) if x2e is not None else None)
            if nchw:
                v1.sn = v1.sh = v1.sw = 0
            vdz = ops.view(dz)

            def launch_w():
                if dz_s is not None and A.try_call("segsde_conv2d_wgrad_tc", C.byref(v1), ops._ref(v2),
                                                   C.byref(ops.view(dz_s)), A.ptr(dw), None, C.byref(d_s), st):
                    return ops.log_route("wgrad", True)
                if _tc_enabled() and not nchw and A.try_call("segsde_conv2d_wgrad_tc", C.byref(v1), ops._ref(v2),
                                                             C.byref(vdz), A.ptr(dw), None, C.byref(d), st):
                    return ops.log_route("wgrad", True)
                A.call("segsde_conv2d_wgrad", C.byref(v1), ops._ref(v2), C.byref(vdz), A.ptr(dw), None, C.byref(d), st)
                ops.log_route("wgrad", False)
            ops._timed("wgrad", flops_scale * 2.0 * nz * hz * wz * cout * kh * kw * ctot, launch_w, ctx.desc)
        if phase_dx1 is not None:
            dx1 = phase_dx1
        return dx1, dx2, dw, db, None, None, None, None, None, None, None, None


class _UpConv3x3Fn(torch.autograd.Function):
    """y = act(conv3x3(reflect_pad1(nearest_up2(x))) + bias) — the decoder's `upconv(i,1)` without a skip connection
    (depth_decoder.py:93-101 at the finest stage) — as FOUR 2x2 "phase" convolutions on the replicate-padded low-res
    input (csrc/conv_aux.cu: output pixel (2i+a, 2j+b) sees the low-res rows {i-1,i,i} or {i,i,i+1}): 9/4 of the MACs in
    fprop, dgrad and wgrad, and neither the upsampled + padded copy (4x the low-res bytes) nor its gradient exist."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        A.require_cuda(x, weight)
        x = ops.as_cl(x)
        w = ops.ohwi(weight.detach())
        cout, c1 = w.shape[0], w.shape[1]
        n, _, h, wd = x.shape
        dev, st = x.device, A.stream_ptr()
        xp = ops.cl_empty(n, c1, h + 2, wd + 2, dev)
        A.call("segsde_pad_replicate", C.byref(ops.view(x)), C.byref(ops.view(xp)), st)
        y = ops.cl_empty(n, cout, 2 * h, 2 * wd, dev)
        b = bias.detach() if bias is not None else None
        d22 = ops._desc(2, 2, 1, 0, 1, A.PAD_ZERO, False, act, False)
        desc = "%d->%d k3 up2 out %dx%d phases" % (c1, cout, 2 * h, 2 * wd)
        for a in (0, 1):
            for b_ in (0, 1):
                wp = torch.empty(cout * 4 * c1, device=dev, dtype=torch.float32)
                A.call("segsde_weight_phase_up", A.ptr(w), A.ptr(wp), C.c_int(cout), C.c_int(c1), C.c_int(0), C.c_int(c1),
                       C.c_int(a), C.c_int(b_), st)
                _fwd(xp[:, :, a:a + h + 1, b_:b_ + wd + 1], None, wp, b, y[:, :, a::2, b_::2], d22, "fprop",
                     2.0 * n * h * wd * cout * 9 * c1, desc)
        ctx.save_for_backward(xp, w, y if act != A.ACT_NONE else None)
        ctx.cfg = (act, bias is not None, desc)
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, w, y = ctx.saved_tensors
        act, has_bias, desc = ctx.cfg
        cout, c1 = w.shape[0], w.shape[1]
        n, _, hp, wp_ = xp.shape
        h, wd = hp - 2, wp_ - 2
        dev, st = dy.device, A.stream_ptr()
        dy = ops.as_cl(dy)
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        db = ops.zeros_f32(cout, dev) if (has_bias and need_b) else None
        dz = dy
        if act != A.ACT_NONE or db is not None:
            dzb = ops.cl_empty(*dy.shape, dev) if act != A.ACT_NONE else None
            A.call("segsde_act_bwd_bias", ops._ref(ops.view(y)) if y is not None else None, C.byref(ops.view(dy)),
                   ops._ref(ops.view(dzb)) if dzb is not None else None, C.c_int(act), A.ptr(db), st)
            dz = dy if dzb is None else dzb
        fl = 2.0 * n * h * wd * cout * 9 * c1            # one phase's share of the algorithmic FLOPs of the 3x3 convolution
        dx = dw = None
        wps = {}
        for a in (0, 1):
            for b_ in (0, 1):
                wp = torch.empty(cout * 4 * c1, device=dev, dtype=torch.float32)
                A.call("segsde_weight_phase_up", A.ptr(w), A.ptr(wp), C.c_int(cout), C.c_int(c1), C.c_int(0), C.c_int(c1),
                       C.c_int(a), C.c_int(b_), st)
                wps[(a, b_)] = wp
        if need_x:
            gs = []
            dd = ops._desc(2, 2, 1, 1, 1, A.PAD_ZERO, False, A.ACT_NONE, False)
            for a in (0, 1):
                for b_ in (0, 1):
                    wt = torch.empty(c1 * 4 * cout, device=dev, dtype=torch.float32)
                    A.call("segsde_weight_transpose_flip", A.ptr(wps[(a, b_)]), A.ptr(wt), C.c_int(cout), C.c_int(2), C.c_int(2),
                           C.c_int(c1), C.c_int(0), C.c_int(c1), st)
                    g = ops.cl_empty(n, c1, h + 1, wd + 1, dev)
                    _fwd(dz[:, :, a::2, b_::2], None, wt, None, g, dd, "dgrad", fl, desc)
                    gs.append(g)
            dx = ops.cl_empty(n, c1, h, wd, dev)
            A.call("segsde_phase_up_fold", A.ptr(gs[0]), A.ptr(gs[1]), A.ptr(gs[2]), A.ptr(gs[3]), C.byref(ops.view(dx)), st)
        if need_w:
            dwp = ops.zeros_f32(4 * cout * 4 * c1, dev)
            d22 = ops._desc(2, 2, 1, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False)
            for a in (0, 1):
                for b_ in (0, 1):
                    part = dwp[(a * 2 + b_) * cout * 4 * c1:(a * 2 + b_ + 1) * cout * 4 * c1]
                    v1, vdz = ops.view(xp[:, :, a:a + h + 1, b_:b_ + wd + 1]), ops.view(dz[:, :, a::2, b_::2])

                    def launch_w(v1=v1, vdz=vdz, part=part):
                        if _tc_enabled() and A.try_call("segsde_conv2d_wgrad_tc", C.byref(v1), None, C.byref(vdz), A.ptr(part),
                                                        None, C.byref(d22), st):
                            return ops.log_route("wgrad", True)
                        A.call("segsde_conv2d_wgrad", C.byref(v1), None, C.byref(vdz), A.ptr(part), None, C.byref(d22), st)
                        ops.log_route("wgrad", False)
                    ops._timed("wgrad", fl, launch_w, desc)
            dw = ops.zeros_like_w(w)
            A.call("segsde_weight_phase_up_fold", A.ptr(dwp), A.ptr(dw), C.c_int(cout), C.c_int(c1), C.c_int(0), C.c_int(c1), st)
        return dx, dw, db, None


def _band_view(xp, n, hp, wo, wp, P):
    """The overlapping row-band view of a segsde_stem_pack buffer: pixel (y, ox) exposes the 8*P contiguous floats
    that start at padded pixel (y, 2*ox) — c = 8*P, horizontal stride 2*P floats (include/segsde_b200.h)."""
    v = A.NHWC()
    v.ptr = xp.data_ptr()
    v.n, v.h, v.w, v.c = n, hp, wo, 8 * P
    v.sn, v.sh, v.sw = hp * wp * P, wp * P, 2 * P
    return v


class _StemConvFn(torch.autograd.Function):
    """7x7/s2 stem on NCHW frames (resnet_encoder.py:92-93) on the tensor cores; the input normalisation
    (x-0.45)/0.225 happens inside the packing pass.  Only the weight gets a gradient.
    Route 1 ("band"): zero-haloed NHWC-4/8 copy of the frames + a (7 x 1)-tap implicit GEMM over the overlapping
    row-band view (K = 7*8*P) — the frames are read once.  Route 2 (fallback for odd sizes): im2col + 1x1 GEMM."""

    @staticmethod
    def forward(ctx, x1, x2, weight, stride, pad, stats=None):
        A.require_cuda(x1, weight)
        x1 = x1.contiguous().float()
        x2 = x2.contiguous().float() if x2 is not None else None
        w = ops.ohwi(weight.detach())
        cout, ctot, kh, kw = w.shape
        n, c1, h, wd = x1.shape
        c2 = x2.shape[1] if x2 is not None else 0
        k = kh * kw * ctot
        ho, wo = ops._conv_out_hw(h, wd, kh, kw, stride, pad, 1)
        st = A.stream_ptr()
        dev = x1.device
        y = ops.cl_empty(n, cout, ho, wo, dev)
        flops = 2.0 * n * ho * wo * cout * k
        if (kh == 7 and kw == 7 and stride == 2 and pad == 3 and h % 2 == 0 and wd % 2 == 0 and ctot <= 8
                and wo % 32 == 0 and cout % 32 == 0 and os.environ.get("SEGSDE_STEM_BAND", "1") != "0"):
            P = 4 if ctot <= 4 else 8
            wp = wd + 8
            xp = torch.empty(n * (h + 6) * wp * P, device=dev, dtype=torch.float32)
            A.call("segsde_stem_pack", A.ptr(x1), A.ptr(x2), C.c_int(c1), C.c_int(c2), C.c_int(n), C.c_int(h), C.c_int(wd),
                   C.c_int(3), C.c_int(wp), C.c_int(P), A.ptr(xp), st)
            wpk = torch.empty(cout * kh * 8 * P, device=dev, dtype=torch.float32)
            A.call("segsde_stem_pack_w", A.ptr(w), A.ptr(wpk), C.c_int(cout), C.c_int(kh), C.c_int(kw), C.c_int(ctot),
                   C.c_int(P), C.c_int(0), st)
            band, vy = _band_view(xp, n, h + 6, wo, wp, P), ops.view(y)
            d = ops._desc(kh, 1, 2, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False, stride_w=1)
            desc = "stem %d+%d->%d k%d s%d out %dx%d band" % (c1, c2, cout, kh, stride, ho, wo)
            done = []

            def launch():
                if stats is not None and A.try_call("segsde_conv2d_fwd_tc_stats", C.byref(band), None, A.ptr(wpk), None,
                                                    C.byref(vy), C.byref(d), A.ptr(stats[0]), st):
                    stats[1] = True
                    done.append(1)
                    ops.log_route("fprop", True)
                elif A.try_call("segsde_conv2d_fwd_tc", C.byref(band), None, A.ptr(wpk), None, C.byref(vy), C.byref(d), st):
                    done.append(1)
                    ops.log_route("fprop", True)
            ops._timed("fprop", flops, launch, desc)
            if done:
                ctx.save_for_backward(xp, w)
                ctx.cfg = ("band", (n, h + 6, wo, wp, P, kh, kw, ctot), flops, desc)
                return y
        kpad = (k + 31) // 32 * 32
        cols = ops.cl_empty(n, kpad, ho, wo, dev)
        A.call("segsde_stem_im2col", A.ptr(x1), A.ptr(x2), C.c_int(c1), C.c_int(c2), C.c_int(n), C.c_int(h), C.c_int(wd),
               C.c_int(kh), C.c_int(kw), C.c_int(stride), C.c_int(pad), C.c_int(kpad), A.ptr(cols), st)
        wpad = torch.empty(cout * kpad, device=dev, dtype=torch.float32)
        A.call("segsde_copy_rows", A.ptr(w), C.c_int(k), A.ptr(wpad), C.c_int(kpad), C.c_int(cout), C.c_int(k), st)
        d = ops._desc(1, 1, 1, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False)
        desc = "stem %d+%d->%d k%d s%d out %dx%d im2col" % (c1, c2, cout, kh, stride, ho, wo)
        _fwd(cols, None, wpad, None, y, d, "fprop", flops, desc, stats)
        ctx.save_for_backward(cols, w)
        ctx.cfg = ("im2col", (k, kpad), flops, desc)
        return y

    @staticmethod
    def backward(ctx, dy):
        src, w = ctx.saved_tensors
        mode, geo, flops, desc = ctx.cfg
        if not ctx.needs_input_grad[2]:
            return None, None, None, None, None, None
        cout = w.shape[0]
        dy = ops.as_cl(dy)
        st = A.stream_ptr()
        vdz = ops.view(dy)
        dw = torch.empty_like(w)
        if mode == "band":
            n, hp, wo, wp, P, kh, kw, ctot = geo
            band = _band_view(src, n, hp, wo, wp, P)
            dwp = ops.zeros_f32(cout * kh * 8 * P, dy.device)
            d = ops._desc(kh, 1, 2, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False, stride_w=1)
            ops._timed("wgrad", flops, lambda: A.call("segsde_conv2d_wgrad_tc", C.byref(band), None, C.byref(vdz), A.ptr(dwp),
                                                      None, C.byref(d), st), desc)
            ops.log_route("wgrad", True)
            A.call("segsde_stem_pack_w", A.ptr(dw), A.ptr(dwp), C.c_int(cout), C.c_int(kh), C.c_int(kw), C.c_int(ctot),
                   C.c_int(P), C.c_int(1), st)
            return None, None, dw, None, None, None
        k, kpad = geo
        dwp = torch.zeros(cout * kpad, device=dy.device, dtype=torch.float32)
        d = ops._desc(1, 1, 1, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False)
        v1 = ops.view(src)

        def launch_w():
            if A.try_call("segsde_conv2d_wgrad_tc", C.byref(v1), None, C.byref(vdz), A.ptr(dwp), None, C.byref(d), st):
                return ops.log_route("wgrad", True)
            A.call("segsde_conv2d_wgrad", C.byref(v1), None, C.byref(vdz), A.ptr(dwp), None, C.byref(d), st)
            ops.log_route("wgrad", False)
        ops._timed("wgrad", flops, launch_w, desc)
        A.call("segsde_copy_rows", A.ptr(dwp), C.c_int(kpad), A.ptr(dw), C.c_int(k), C.c_int(cout), C.c_int(k), st)
        return None, None, dw, None, None, None


class _HeadConvFn(torch.autograd.Function):
    """C -> 1, 3x3, pad 1 (the sigmoid disparity heads, depth_decoder.py:69-70,107-112):
    z = 1x1conv(x) to 9 tap planes + a 9-tap stencil; backward = adjoint stencil of dy (gcol) + two 1x1 GEMMs
    (dgrad: gcol x Wz, wgrad: x^T x gcol).  Default: N padded to 32 planes on the tcgen05 kernels;
    SEGSDE_HEAD_FEWCOUT=1: 12 planes (9 used) on the few-output-channel fp32 kernels of conv_fewcout.cu (measured
    slower, kept as the fp32 form of the same decomposition)."""

    @staticmethod
    def forward(ctx, x, weight, bias, pad_mode, act):
        A.require_cuda(x, weight)
        x = ops.as_cl(x)
        w = ops.ohwi(weight.detach())            # [1][3][3][C] == [9][C]
        n, c, h, wd = x.shape
        dev, st = x.device, A.stream_ptr()
        y = ops.cl_empty(n, 1, h, wd, dev)
        reflect = int(pad_mode == A.PAD_REFLECT)
        b = bias.detach() if bias is not None else None
        flops = 2.0 * n * h * wd * 9 * c
        few = HEAD_FEWCOUT
        planes = 12 if few else 32
        wz = ops.zeros_f32(planes * c, dev)
        A.call("segsde_copy_rows", A.ptr(w), C.c_int(c), A.ptr(wz), C.c_int(c), C.c_int(9), C.c_int(c), st)
        z = ops.cl_empty(n, planes, h, wd, dev)
        d1 = ops._desc(1, 1, 1, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False)
        desc = "head %d->1 k3 out %dx%d (1x1 GEMM%s + stencil)" % (c, h, wd, " fp32" if few else "")
        if few:
            vx, vz = ops.view(x), ops.view(z)

            def launch():
                A.call("segsde_conv2d_fwd", C.byref(vx), None, A.ptr(wz), None, C.byref(vz), C.byref(d1), st)
                ops.log_route("fprop", False, "cc:fewcout")
            ops._timed("fprop", flops, launch, desc)
        else:
            _fwd(x, None, wz, None, z, d1, "fprop", flops, desc)
        A.call("segsde_head_stencil_fwd", C.byref(ops.view(z)), C.byref(ops.view(y)), A.ptr(b), C.c_int(act),
               C.c_int(reflect), C.c_int(1), st)
        ctx.save_for_backward(x, w, wz, y if act != A.ACT_NONE else None)
        ctx.cfg = (reflect, act, bias is not None, desc, few, planes)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, wz, y = ctx.saved_tensors
        reflect, act, has_bias, desc, few, planes = ctx.cfg
        n, c, h, wd = x.shape
        dev, st = x.device, A.stream_ptr()
        dy = ops.as_cl(dy)
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        db = ops.zeros_f32(1, dev) if (has_bias and need_b) else None
        if act != A.ACT_NONE or db is not None:
            dz = ops.cl_empty(*dy.shape, dev) if act != A.ACT_NONE else None
            A.call("segsde_act_bwd_bias", ops._ref(ops.view(y)) if y is not None else None, C.byref(ops.view(dy)),
                   ops._ref(ops.view(dz)) if dz is not None else None, C.c_int(act), A.ptr(db), st)
            dz = dy if dz is None else dz
        else:
            dz = dy
        dx = dw = None
        if need_x or need_w:
            gcol = ops.cl_empty(n, planes, h, wd, dev)
            A.call("segsde_head_gcol", C.byref(ops.view(dz)), C.byref(ops.view(gcol)), C.c_int(reflect), C.c_int(1), st)
            d1 = ops._desc(1, 1, 1, 0, 1, A.PAD_ZERO, False, A.ACT_NONE, False)
            fl = 2.0 * n * h * wd * 9 * c
            if need_x:
                dx = ops.cl_empty(n, c, h, wd, dev)
                if few:
                    vg, vdx = ops.view(gcol), ops.view(dx)

                    def launch_x():
                        A.call("segsde_conv2d_dgrad", C.byref(vg), A.ptr(wz), C.byref(vdx), None, C.byref(d1), st)
                        ops.log_route("dgrad", False, "cc:fewcout")
                    ops._timed("dgrad", fl, launch_x, desc)
                else:
                    wzt = torch.empty(c * planes, device=dev, dtype=torch.float32)
                    A.call("segsde_weight_transpose_flip", A.ptr(wz), A.ptr(wzt), C.c_int(planes), C.c_int(1), C.c_int(1),
                           C.c_int(c), C.c_int(0), C.c_int(c), st)
                    _fwd(gcol, None, wzt, None, dx, d1, "dgrad", fl, desc)
            if need_w:
                dwz = ops.zeros_f32(planes * c, dev)
                v1, vg = ops.view(x), ops.view(gcol)

                def launch_w():
                    if not few and A.try_call("segsde_conv2d_wgrad_tc", C.byref(v1), None, C.byref(vg), A.ptr(dwz), None,
                                              C.byref(d1), st):
                        return ops.log_route("wgrad", True)
                    A.call("segsde_conv2d_wgrad", C.byref(v1), None, C.byref(vg), A.ptr(dwz), None, C.byref(d1), st)
                    ops.log_route("wgrad", False, "cc:fewcout" if few else None)
                ops._timed("wgrad", fl, launch_w, desc)
                dw = torch.empty_like(w)
                A.call("segsde_copy_rows", A.ptr(dwz), C.c_int(c), A.ptr(dw), C.c_int(c), C.c_int(9), C.c_int(c), st)
        return dx, dw, db, None, None


def conv2d(x1, weight, bias=None, x2=None, stride=1, pad=0, dil=1, pad_mode=A.PAD_ZERO, up1=False,
           act=A.ACT_NONE, nchw_norm_in=False, bn_stats=None):
    """y = act(conv(cat(up?(x1), x2)) + bias) — see segsde_conv2d_fwd / segsde_conv2d_fwd_tc.
    bn_stats: optional zero-filled float64 tensor [3*Cout]; on return it holds the BatchNorm batch statistics of y
    (sum, sum of squares, shift) — from the convolution epilogue when possible, else from segsde_bn_stats."""
    if bn_stats is not None:
        box = [bn_stats, False]
        if (nchw_norm_in and bias is None and act == A.ACT_NONE and dil == 1 and weight.shape[0] % 64 == 0 and _tc_enabled()):
            y = _StemConvFn.apply(x1, x2, weight, stride, pad, box)
        else:
            y = _Conv2dFn.apply(x1, x2, weight, bias, stride, pad, dil, pad_mode, up1, act, nchw_norm_in, box)
        if not box[1]:
            A.call("segsde_bn_stats", C.byref(ops.view(y.detach())), A.ptr(bn_stats), A.stream_ptr())
        return y
    if (PHASE_UPCONV and up1 and x2 is None and tuple(weight.shape[2:]) == (3, 3) and stride == 1 and pad == 1 and dil == 1
            and pad_mode == A.PAD_REFLECT and not nchw_norm_in and bn_stats is None and weight.shape[1] % 64 == 0
            and weight.shape[0] % 64 == 0 and x1.shape[-1] % 32 == 0 and x1.shape[-2] >= 2 and _tc_enabled()):
        return _UpConv3x3Fn.apply(x1, weight, bias, act)
    if (weight.shape[0] == 1 and tuple(weight.shape[2:]) == (3, 3) and x2 is None and not up1 and stride == 1
            and pad == 1 and dil == 1 and not nchw_norm_in and weight.shape[1] % 64 == 0 and x1.shape[-1] % 32 == 0
            and _tc_enabled()):
        return _HeadConvFn.apply(x1, weight, bias, pad_mode, act)
    if (nchw_norm_in and bias is None and act == A.ACT_NONE and dil == 1 and weight.shape[0] % 64 == 0
            and _tc_enabled()):
        return _StemConvFn.apply(x1, x2, weight, stride, pad)
    return _Conv2dFn.apply(x1, x2, weight, bias, stride, pad, dil, pad_mode, up1, act, nchw_norm_in)
