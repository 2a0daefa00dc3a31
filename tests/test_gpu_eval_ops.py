"""GPU: validation-path, label-selection and input-pipeline ops against their definitions.

 * confusion matrix / scores: the reference's `runningScore` arithmetic (evaluation/metrics.py:12-57, restated with numpy
   bincount exactly as there) on the same labels / predictions — integer counts must match exactly.
 * BatchNorm folding: eval-mode network forward with folding on vs off and vs the CPU oracle.
 * adaptive pooling / cdist / iterative farthest point: torch's ops and a transcription of label_selection.py:617-640.
 * Gaussian blur / colour jitter / area pyramid: PyTorch restatements of the kornia-0.4 definitions (kornia itself is
   absent here — "parity unpinned" for those two) and F.interpolate(mode="area").
 * GradScaler: torch.cuda.amp.GradScaler on the same gradient sequence incl. an overflow step.
"""
import contextlib
import io
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import segsde_oracle as O
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _ref_hist(gt, pred, n):           # evaluation/metrics.py:12-17
    mask = (gt >= 0) & (gt < n)
    return np.bincount(n * gt[mask].astype(int) + pred[mask], minlength=n ** 2).reshape(n, n)


def test_running_score_matches_reference_arithmetic():
    from improving_segmentation_with_selfsupervised_depth_b200.evaluation import RunningScore
    g = torch.Generator().manual_seed(0)
    n = 19
    rs, hist = RunningScore(n), np.zeros((n, n))
    for fmt in (torch.contiguous_format, torch.channels_last):
        logits = torch.randn(3, n, 40, 72, generator=g)
        lbl = torch.randint(0, n, (3, 40, 72), generator=g)
        lbl[torch.rand(3, 40, 72, generator=g) < 0.1] = 250
        rs.update_from_logits(lbl.cuda(), logits.cuda().contiguous(memory_format=fmt))
        pred = logits.max(1)[1].numpy()
        for lt, lp in zip(lbl.numpy(), pred):
            hist += _ref_hist(lt.flatten(), lp.flatten(), n)
    assert np.array_equal(rs.confusion_matrix, hist)
    # reference signature with host arrays of predictions
    pred2 = torch.randint(0, n, (2, 16, 16), generator=g).numpy()
    gt2 = torch.randint(0, n, (2, 16, 16), generator=g).numpy()
    rs.update(gt2, pred2)
    for lt, lp in zip(gt2, pred2):
        hist += _ref_hist(lt.flatten(), lp.flatten(), n)
    assert np.array_equal(rs.confusion_matrix, hist)
    score, cls_iu = rs.get_scores()
    iu = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
    assert abs(score["Mean IoU : \t"] - np.nanmean(iu)) < 1e-12
    assert abs(score["Overall Acc: \t"] - np.diag(hist).sum() / hist.sum()) < 1e-12
    rs.reset()
    assert rs.confusion_matrix.sum() == 0


def test_eval_forward_with_folded_batchnorm(contracts):
    """Inference with BatchNorm folded into the convolutions == unfolded eval forward == CPU oracle (fp32 route)."""
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    from improving_segmentation_with_selfsupervised_depth_b200.models import layers
    from improving_segmentation_with_selfsupervised_depth_b200.synthetic import mono_config
    ops.USE_TC = False
    models, _ = P.install_dropin()
    B, H, W = 2, 64, 96
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.get_model(mono_config("resnet50", H, W), 19)
    sd = O.synthetic_state_dict(model.state_dict(), seed=4)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    inputs = O.synthetic_inputs(B, H, W, seed=2)
    gin = {k: v.cuda() for k, v in inputs.items()}
    outs, launches = {}, {}
    for fold in (True, False):
        layers.FOLD_EVAL_BN = fold
        n0 = A.launch_count()
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            outs[fold] = model(gin)
        launches[fold] = A.launch_count() - n0
    layers.FOLD_EVAL_BN = True
    with torch.no_grad():
        ref = O.model_forward(sd, inputs, {"num_layers": 50, "rswd": [False, False, True], "frame_ids": [0, -1, 1]}, O.BNMode(False))
    for s in range(4):
        assert rel_err(outs[True][("disp", s)], ref[("disp", s)]) < 4e-4, s          # folding re-associates gamma / sqrt(var)
        assert rel_err(outs[False][("disp", s)], ref[("disp", s)]) < 2e-4, s
        assert rel_err(outs[True][("disp", s)], outs[False][("disp", s)]) < 4e-4, s
    assert launches[True] < launches[False]          # the BatchNorm passes are gone
    assert rel_err(outs[True][("cam_T_cam", 0, 1)], ref[("cam_T_cam", 0, 1)]) < 1e-4


def test_label_selection_ops():
    from improving_segmentation_with_selfsupervised_depth_b200 import label_ops as L
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 5, 37, 53, generator=g)
    for mode, fn in (("avg", F.adaptive_avg_pool2d), ("max", F.adaptive_max_pool2d)):
        assert rel_err(L.adaptive_pool2d(x.cuda(), (4, 8), mode), fn(x, (4, 8))) < 2e-6
        assert rel_err(L.adaptive_pool2d(x[0].cuda(), (6, 12), mode), fn(x[0], (6, 12))) < 2e-6
    feats = [torch.randn(1, 7, 4, 8, generator=g) for _ in range(45)]
    for p in (1, 2, 3):
        d = L.calc_feature_distance([f.cuda() for f in feats], p=p)
        f = torch.cat(feats).flatten(start_dim=1)
        ref = torch.cdist(f, f, p=p)
        ref.fill_diagonal_(0)
        assert rel_err(d, ref) < 1e-5, p
    dn = L.calc_feature_distance([f.cuda() for f in feats], p=2, normalize_features=True)
    fc = torch.cat(feats)
    std, mean = torch.std_mean(fc, dim=[0, 2, 3], keepdim=True)
    fn_ = ((fc - mean) / std).flatten(start_dim=1)
    refn = torch.cdist(fn_, fn_, p=2)
    refn.fill_diagonal_(0)
    assert rel_err(dn, refn) < 1e-5

    # iterative_farthest_point: transcription of label_selection.py:617-640 on the same matrix
    def reference_ifp(current, fd, n_new, preselected=None):
        dist = fd["distances"].clone()
        cur = [fd["img_idx_to_dist_i"][s] for s in current]
        if preselected is not None:
            pre = [fd["img_idx_to_dist_i"][s] for s in preselected]
            ign = [i for i in range(dist.shape[0]) if i not in pre]
            dist[:, ign] = 0
        new, dists = [], []
        for _ in range(n_new):
            m = torch.min(dist[cur, :], dim=0)
            far = torch.max(m.values, dim=0)
            ns = far.indices.item()
            if ns in cur:
                break
            cur.append(ns)
            new.append(ns)
            dists.append(far.values)
        return [fd["dist_i_to_img_idx"][s] for s in new], dists
    ids = [100 + 3 * i for i in range(45)]
    fd_cpu = {"distances": ref.clone(), "dist_i_to_img_idx": dict(enumerate(ids)), "img_idx_to_dist_i": {v: k for k, v in enumerate(ids)}}
    fd_gpu = dict(fd_cpu, distances=ref.cuda())
    for pre in (None, ids[::2]):
        a, da = L.iterative_farthest_point([ids[0], ids[4]], fd_gpu, 12, pre)
        b, db = reference_ifp([ids[0], ids[4]], fd_cpu, 12, pre)
        assert a == b, (a, b)
        assert all(abs(float(x) - float(y)) < 1e-6 for x, y in zip(da, db))
    a, _ = L.iterative_farthest_point([ids[0]], fd_gpu, 100, ids[:5])       # stops once the candidates are exhausted
    b, _ = reference_ifp([ids[0]], fd_cpu, 100, ids[:5])
    assert a == b and len(a) < 100


def test_augment_ops():
    from improving_segmentation_with_selfsupervised_depth_b200 import augment as G
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 64, 96, generator=g)
    # Gaussian blur: reflect pad + depthwise conv with the outer product of the normalised 1-D Gaussians
    for (ky, kx), sigma in (((7, 11), 0.8), ((51, 33), 1.15), ((1, 5), 0.3)):
        y = G.gaussian_blur2d(x.cuda(), (ky, kx), (sigma, sigma))
        ty, tx = G.gaussian_taps(ky, sigma, "cpu").double(), G.gaussian_taps(kx, sigma, "cpu").double()
        k2 = torch.outer(ty, tx)[None, None].repeat(3, 1, 1, 1)
        ref = F.conv2d(F.pad(x.double(), (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect"), k2, groups=3)
        assert rel_err(y, ref) < 2e-6, (ky, kx)
    big = torch.rand(1, 3, 512, 1024, generator=g)
    np.random.seed(3)
    yb, _ = G.gaussian_blur(0.9, data=big.cuda())
    np.random.seed(3)
    sigma = np.random.uniform(0.15, 1.15)
    ty, tx = G.gaussian_taps(51, sigma, "cpu").double(), G.gaussian_taps(103, sigma, "cpu").double()
    ref = F.conv2d(F.pad(big.double(), (51, 51, 25, 25), mode="reflect"), torch.outer(ty, tx)[None, None].repeat(3, 1, 1, 1), groups=3)
    assert rel_err(yb, ref) < 2e-6
    assert G.gaussian_blur(0.3, data=big)[0] is big          # below the threshold: untouched (transformsgpu.py:24)
    # area pyramid == F.interpolate(mode="area")
    for s, y in zip((1, 2, 3), G.area_pyramid(x.cuda())):
        assert rel_err(y, F.interpolate(x, size=(64 >> s, 96 >> s), mode="area")) < 2e-6, s

    # colour jitter primitives against a torch restatement of the same definitions
    def hsv(img):
        r, gch, b = img[:, 0], img[:, 1], img[:, 2]
        mx, mn = img.max(1)[0], img.min(1)[0]
        d = mx - mn
        dd = torch.where(d == 0, torch.ones_like(d), d)
        h = torch.where(mx == r, (gch - b) / dd, torch.where(mx == gch, 2 + (b - r) / dd, 4 + (r - gch) / dd)) / 6
        h = (h - torch.floor(h)) * 2 * math.pi
        return h, d / (mx + 1e-6), mx

    def rgb(h, s, v):
        hh = h / (2 * math.pi) * 6
        hi = torch.floor(hh)
        f = hh - hi
        i = hi.long() % 6
        p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
        sel = lambda a: torch.stack(a, 0).gather(0, i[None])[0]        # noqa: E731
        return torch.stack([sel([v, q, p, p, t, v]), sel([t, v, v, q, p, p]), sel([p, p, t, v, v, q])], 1)
    xd = x.double()
    y = G.apply_color_jitter(x.cuda(), brightness=0.1, order=(0, 1, 2, 3))
    assert rel_err(y, (xd + 0.1).clamp(0, 1)) < 5e-6
    y = G.apply_color_jitter(x.cuda(), contrast=1.2)
    assert rel_err(y, (xd * 1.2).clamp(0, 1)) < 5e-6
    h, s, v = hsv(xd)
    y = G.apply_color_jitter(x.cuda(), saturation=0.8)
    assert rel_err(y, rgb(h, (s * 0.8).clamp(0, 1), v)) < 2e-5
    y = G.apply_color_jitter(x.cuda(), hue=0.5)
    hh = torch.remainder(h + 0.5, 2 * math.pi)
    assert (y.double().cpu() - rgb(hh, s, v)).abs().median().item() < 1e-5      # isolated sector-boundary pixels may differ
    torch.manual_seed(5)
    yj, _ = G.color_jitter(0.9, data=x.cuda())
    assert yj.shape == x.shape and float((yj.cpu() - x).abs().max()) > 1e-3
    assert G.color_jitter(0.1, data=x)[0] is x


def test_grad_scaler_matches_torch():
    from improving_segmentation_with_selfsupervised_depth_b200 import optim as Opt
    g = torch.Generator().manual_seed(6)
    shapes = [(64, 32, 3, 3), (128,), (19, 64, 1, 1), (300000,)]
    p0 = [torch.randn(s, generator=g) for s in shapes]
    seq = []
    for step in range(6):
        gr = [torch.randn(s, generator=g) * 1e-3 for s in shapes]
        if step == 2:
            gr[1][5] = float("inf")
        if step == 4:
            gr[3][77] = float("nan")
        seq.append(gr)

    def run(own):
        ps = [torch.nn.Parameter(t.clone().cuda()) for t in p0]
        if own:
            opt, sc = Opt.SGD(ps, lr=0.1, momentum=0.9), Opt.GradScaler(init_scale=1024.0, growth_interval=2)
        else:
            opt, sc = torch.optim.SGD(ps, lr=0.1, momentum=0.9), torch.cuda.amp.GradScaler(init_scale=1024.0, growth_interval=2)
        scales = []
        for gr in seq:
            sc.scale(torch.ones(1, device="cuda"))           # torch's scaler initialises its state lazily in scale()
            for p, gg in zip(ps, gr):
                p.grad = (gg.cuda() * sc.get_scale()).clone()
            sc.unscale_(opt)
            sc.step(opt)
            sc.update()
            scales.append(sc.get_scale())
        return [p.detach().cpu() for p in ps], scales
    mine, s1 = run(True)
    ref, s2 = run(False)
    assert s1 == s2, (s1, s2)
    for a, b in zip(mine, ref):
        assert rel_err(a, b) < 2e-6
