"""MonodepthLoss — drop-in for the reference class (loss/monodepth_loss.py:16-192).

Same constructor kwargs, same methods (`generate_images_pred`, `compute_losses`,
`generate_depth_test_pred`) and the same `{"loss", "loss/<s>"}` result dict, but the whole
per-scale chain (bilinear disparity upsample -> depth -> backproject -> project -> grid_sample ->
SSIM+L1 -> identity auto-mask -> min -> mean, plus the edge-aware smoothness term) runs as fused
sm_100a kernels behind the C-ABI — ONE launch covers all scales — and its gradient w.r.t. the
disparities and the pose matrices is produced in the same pass (see csrc/reproj.cu).
"""
import ctypes as C

import torch

from .. import _cabi as A


class _MonoLossFn(torch.autograd.Function):
    """(disp_0..S-1, T_0..F-1) -> tensor [S+1] = (loss/0, ..., loss/S-1, loss)."""

    @staticmethod
    def forward(ctx, owner, tgt, srcs, colors, K, inv_K, noises, sel_out, n_disp, *dyn):
        disps, Ts = dyn[:n_disp], dyn[n_disp:]
        A.require_cuda(tgt, K, inv_K, *disps, *Ts)
        S, F = len(disps), len(srcs)
        B, _, H, W = tgt.shape
        dev = tgt.device
        st = A.stream_ptr()
        need_grad = any(ctx.needs_input_grad[9:])
        if S > A.REPROJ_MAX_SCALES:
            raise A.SegsdeError("MonodepthLoss: at most %d scales per fused launch (got %d)" % (A.REPROJ_MAX_SCALES, S))
        tiles = A.lib().segsde_reproj_tiles(C.c_int(H), C.c_int(W))
        reproj = torch.empty(S, device=dev, dtype=torch.float32)
        sacc = torch.zeros(S, 2 + B, device=dev, dtype=torch.float32)
        out = torch.empty(S + 1, device=dev, dtype=torch.float32)
        partial = torch.empty(S, B * tiles, device=dev, dtype=torch.float32)
        gdisps, gTs, gT_partial = [], None, None
        if need_grad:
            gTs = torch.empty(S, F, B, 16, device=dev, dtype=torch.float32)
            gT_partial = torch.empty(S, F, B, tiles, 12, device=dev, dtype=torch.float32)
        flags = (A.REPROJ_NO_SSIM if owner.no_ssim else 0) | (A.REPROJ_AVG if owner.avg_reprojection else 0) \
            | (A.REPROJ_NO_AUTOMASK if owner.disable_automasking else 0)
        Kc, iKc = K.contiguous(), inv_K.contiguous()
        Tc = [t.detach().contiguous() for t in Ts]
        owner._step += 1
        ds = []
        for s in range(S):
            d = disps[s].detach()
            ds.append(d if d.is_contiguous() else d.contiguous())
        # ---- one launch: every scale's warp + SSIM/L1 + auto-mask + min + mean (+ gradients) -------------------
        a = A.ReprojArgs()
        a.tgt = tgt.data_ptr()
        for f in range(F):
            a.src[f] = srcs[f].data_ptr()
            a.T[f] = Tc[f].data_ptr()
        a.K, a.inv_K = Kc.data_ptr(), iKc.data_ptr()
        a.B, a.H, a.W, a.F, a.S = B, H, W, F, S
        a.seed, a.offset = owner.seed, owner._step * 8
        a.min_depth, a.max_depth, a.flags = owner.min_depth, owner.max_depth, flags
        a.loss_partial = partial.data_ptr()
        a.gT_partial = gT_partial.data_ptr() if need_grad else None
        for s in range(S):
            a.disp[s], a.hs[s], a.ws[s] = ds[s].data_ptr(), ds[s].shape[-2], ds[s].shape[-1]
            a.noise[s] = noises[s].data_ptr() if noises is not None else None
            a.ident_sel[s] = sel_out[s].data_ptr() if sel_out is not None else None
            if need_grad:
                g = torch.zeros_like(ds[s])
                gdisps.append(g)
                a.gdisp[s] = g.data_ptr()
        A.call("segsde_reproj_fused", C.byref(a), st)
        A.call("segsde_reproj_finalize", A.ptr(partial), C.c_int(B * tiles), C.c_int64(B * H * W), C.c_int(S),
               A.ptr(reproj), A.ptr(gT_partial), A.ptr(Kc), C.c_int(B), C.c_int(tiles), C.c_int(F), A.ptr(gTs), st)
        for s in range(S):
            d = ds[s]
            hs, ws = d.shape[-2:]
            # smoothness on the scale's own resolution (monodepth_loss.py:182-186)
            col = colors[s]
            mean = torch.empty(B, device=dev, dtype=torch.float32)
            A.call("segsde_smooth_mean", A.ptr(d), C.c_int(B), C.c_int(hs), C.c_int(ws), A.ptr(mean), st)
            ghat = torch.empty_like(d) if need_grad else None
            A.call("segsde_smooth_fused", A.ptr(d), A.ptr(col), A.ptr(mean), C.c_int(B), C.c_int(hs),
                   C.c_int(ws), A.ptr(sacc[s]), A.ptr(ghat), st)
            if need_grad:
                A.call("segsde_smooth_grad_finalize", A.ptr(ghat), A.ptr(mean), A.ptr(sacc[s]), C.c_int(B),
                       C.c_int(hs), C.c_int(ws), C.c_float(owner.disparity_smoothness / (2 ** s)),
                       A.ptr(gdisps[s]), st)
        wsm = (C.c_float * S)(*[owner.disparity_smoothness / (2 ** s) for s in range(S)])
        A.call("segsde_mono_combine", A.ptr(reproj), A.ptr(sacc), C.c_int(S), C.c_int(2 + B), wsm,
               A.ptr(out), st)
        ctx.S, ctx.F = S, F
        ctx.gdisps, ctx.gTs = gdisps, gTs
        ctx.shapes = [d.shape for d in disps]
        return out

    @staticmethod
    def backward(ctx, gout):
        S, F = ctx.S, ctx.F
        if ctx.gTs is None:
            return (None,) * (9 + S + F)
        gout = gout.contiguous()
        st = A.stream_ptr()
        g_total = C.c_void_p(gout.data_ptr() + 4 * S)
        grads = []
        for s in range(S):
            if ctx.needs_input_grad[9 + s]:
                g = torch.empty_like(ctx.gdisps[s])
                A.call("segsde_scale_by_dev", A.ptr(ctx.gdisps[s]), g_total, C.c_float(1.0 / S),
                       C.c_void_p(gout.data_ptr() + 4 * s), C.c_float(1.0), A.ptr(g), C.c_int(0),
                       C.c_int64(g.numel()), st)
                grads.append(g.view(ctx.shapes[s]))
            else:
                grads.append(None)
        for f in range(F):
            if ctx.needs_input_grad[9 + S + f]:
                B = ctx.gTs.shape[2]
                g = torch.empty(B, 4, 4, device=gout.device, dtype=torch.float32)
                for s in range(S):
                    A.call("segsde_scale_by_dev", A.ptr(ctx.gTs[s, f]), g_total, C.c_float(1.0 / S),
                           C.c_void_p(gout.data_ptr() + 4 * s), C.c_float(1.0), A.ptr(g),
                           C.c_int(1 if s > 0 else 0), C.c_int64(B * 16), st)
                grads.append(g)
            else:
                grads.append(None)
        return (None,) * 9 + tuple(grads)


class MonodepthLoss:
    """Same signature as the reference (loss/monodepth_loss.py:17-19)."""

    def __init__(self, num_scales, frame_ids, height, width, batch_size, min_depth, max_depth,
                 test_min_depth, test_max_depth, disparity_smoothness,
                 no_ssim, avg_reprojection, disable_automasking, crop_h=None, crop_w=None, is_train=True,
                 materialize_outputs=False, noise="philox", seed=None):
        self.num_scales = num_scales
        self.scales = list(range(self.num_scales))
        self.height = height if crop_h is None or not is_train else crop_h
        self.width = width if crop_w is None or not is_train else crop_w
        self.batch_size = batch_size
        self.frame_ids = frame_ids
        self.min_depth = min_depth
        self.max_depth = max_depth
        self.test_min_depth = test_min_depth
        self.test_max_depth = test_max_depth
        self.disparity_smoothness = disparity_smoothness
        self.no_ssim = no_ssim
        self.avg_reprojection = avg_reprojection
        self.disable_automasking = disable_automasking
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.depth_metric_names = ["abs_rel", "sq_rel", "rms", "log_rms", "a1", "a2", "a3"]
        # side outputs of the reference (("depth",0,s), ("sample",f,s), ("color",f,s), identity_selection)
        # are only written when asked for: nothing on the training path reads them.
        self.materialize_outputs = materialize_outputs
        # "philox": in-kernel counter-based noise; "torch": the reference's own CPU randn stream
        # (monodepth_loss.py:163-164) for seeded parity runs; or set .replay_noise to a list of tensors
        self.noise = noise
        self.replay_noise = None
        self.seed = A.default_seed(0x5E65DE) if seed is None else seed
        self._step = 0
        if not self.no_ssim:
            from ..models.monodepth_layers import SSIM
            self.ssim = SSIM()
        from ..models.monodepth_layers import BackprojectDepth, Project3D
        self.backproject_depth = {}
        self.project_3d = {}
        for scale in self.scales:
            h = self.height // (2 ** scale)
            w = self.width // (2 ** scale)
            self.backproject_depth[scale] = BackprojectDepth(self.batch_size, h, w)
            self.project_3d[scale] = Project3D(self.batch_size, h, w)

    # ------------------------------------------------------------------------------------------
    def _src_frames(self, inputs, outputs):
        srcs, Ts = [], []
        for frame_id in self.frame_ids[1:]:
            srcs.append(inputs[("color", frame_id, 0)])
            Ts.append(inputs["stereo_T"] if frame_id == "s" else outputs[("cam_T_cam", 0, frame_id)])
        return srcs, Ts

    def generate_depth_test_pred(self, outputs):
        assert outputs[("disp", 0)].shape[-2:] == (self.height, self.width), outputs[("disp", 0)].shape[-2:]
        st = A.stream_ptr()
        for scale in self.scales:
            disp = outputs[("disp", scale)].detach().contiguous()
            A.require_cuda(disp)
            B = disp.shape[0]
            depth = torch.empty(B, 1, self.height, self.width, device=disp.device, dtype=torch.float32)
            A.call("segsde_disp_to_depth_up", A.ptr(disp), C.c_int(B), C.c_int(disp.shape[-2]),
                   C.c_int(disp.shape[-1]), C.c_int(self.height), C.c_int(self.width),
                   C.c_float(self.test_min_depth), C.c_float(self.test_max_depth), A.ptr(depth), st)
            outputs[("depth", 0, scale)] = depth

    def generate_images_pred(self, inputs, outputs):
        """Reference: monodepth_loss.py:64-102. The warp itself is fused into compute_losses; here only
        the shape contract is checked and, on request, the side outputs are materialised."""
        assert outputs[("disp", 0)].shape[-2:] == (
            self.height, self.width), f'{outputs[("disp", 0)].shape[-2:]} should be {(self.height, self.width)} '
        if not self.materialize_outputs:
            return
        st = A.stream_ptr()
        srcs, Ts = self._src_frames(inputs, outputs)
        K, iK = inputs[("K", 0)].contiguous(), inputs[("inv_K", 0)].contiguous()
        for scale in self.scales:
            disp = outputs[("disp", scale)].detach().contiguous()
            B = disp.shape[0]
            dev = disp.device
            depth = torch.empty(B, 1, self.height, self.width, device=dev, dtype=torch.float32)
            for i, frame_id in enumerate(self.frame_ids[1:]):
                sample = torch.empty(B, self.height, self.width, 2, device=dev, dtype=torch.float32)
                color = torch.empty(B, 3, self.height, self.width, device=dev, dtype=torch.float32)
                src = srcs[i].contiguous()
                A.call("segsde_reproj_materialize", A.ptr(src), A.ptr(disp), A.ptr(K), A.ptr(iK),
                       A.ptr(Ts[i].detach().contiguous()), C.c_int(B), C.c_int(self.height),
                       C.c_int(self.width), C.c_int(disp.shape[-2]), C.c_int(disp.shape[-1]),
                       C.c_float(self.min_depth), C.c_float(self.max_depth), A.ptr(depth), A.ptr(sample),
                       A.ptr(color), st)
                outputs[("sample", frame_id, scale)] = sample
                outputs[("color", frame_id, scale)] = color
                if not self.disable_automasking:
                    outputs[("color_identity", frame_id, scale)] = inputs[("color", frame_id, 0)]
            outputs[("depth", 0, scale)] = depth

    def _noise(self, B, F, dev):
        if self.disable_automasking:
            return None
        nf = 1 if self.avg_reprojection else F
        if self.replay_noise is not None:
            return [n.to(dev).contiguous() for n in self.replay_noise]
        if self.noise == "torch":   # exactly the reference's RNG consumption (one CPU randn per scale)
            return [(torch.randn(B, nf, self.height, self.width).to(dev) * 0.00001) for _ in self.scales]
        return None

    def compute_losses(self, inputs, outputs):
        """Reference: monodepth_loss.py:118-192."""
        tgt = inputs[("color", 0, 0)].contiguous()
        srcs, Ts = self._src_frames(inputs, outputs)
        srcs = [s.contiguous() for s in srcs]
        disps = [outputs[("disp", s)] for s in self.scales]
        colors = [inputs[("color", 0, s)].contiguous() for s in self.scales]
        B = tgt.shape[0]
        noises = self._noise(B, len(srcs), tgt.device)
        sel = None
        if self.materialize_outputs and not self.disable_automasking:
            sel = [torch.empty(B, self.height, self.width, device=tgt.device) for _ in self.scales]
        Ts = [t.to(torch.float32) for t in Ts]
        out = _MonoLossFn.apply(self, tgt, srcs, colors, inputs[("K", 0)], inputs[("inv_K", 0)], noises, sel,
                                len(disps), *[d.to(torch.float32) for d in disps], *Ts)
        if sel is not None:
            for s in self.scales:
                outputs["identity_selection/{}".format(s)] = sel[s]
        losses = {"loss/{}".format(s): out[s] for s in self.scales}
        losses["loss"] = out[len(self.scales)]
        return losses
