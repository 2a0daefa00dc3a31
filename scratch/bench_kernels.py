"""Standalone timing of the hot kernels at the bench shapes (B=12, 512x1024): CUDA events, 3 warm-up + 5 timed."""
import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
from improving_segmentation_with_selfsupervised_depth_b200.synthetic import MONO_LOSS_KW, synthetic_inputs
models, loss = P.install_dropin()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
B, H, W = 12, 512, 1024
dev = torch.device('cuda')

def timeit(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

if which in ("all", "reproj"):
    inputs = {k: v.to(dev) for k, v in synthetic_inputs(B, H, W).items()}
    ml = loss.MonodepthLoss(height=H, width=W, batch_size=B, **MONO_LOSS_KW)
    disps = [torch.rand(B, 1, H >> s, W >> s, device=dev) * 0.6 + 0.2 for s in range(4)]
    T = {f: torch.eye(4, device=dev).repeat(B, 1, 1) for f in (-1, 1)}
    T[-1][:, 0, 3] = 0.05; T[1][:, 0, 3] = -0.05
    for req in (False, True):
        def run():
            out = {("disp", s): disps[s].clone().requires_grad_(req) for s in range(4)}
            out.update({("cam_T_cam", 0, f): T[f] for f in (-1, 1)})
            ml.generate_images_pred(inputs, out)
            return ml.compute_losses(inputs, out)["loss"]
        A.PROFILE = None
        run(); run()
        A.PROFILE = []
        for _ in range(3): run()
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for n, a, b in A.PROFILE if n == "segsde_reproj_fused"]
        A.PROFILE = None
        per = sum(ts) / len(ts)
        byt = B * (4 * 9 * H * W + 4 * H * W)     # scale-0 algorithmic bytes per launch, coarser scales read less
        print("reproj_fused grad=%s: %.3f ms/launch (4 scales avg), scale-0 algorithmic %.1f MB -> %.0f GB/s" % (req, per, byt / 1e6, byt / per / 1e6))

def conv_case(c1, c2, cout, k, h, w, reflect, up1, name):
    x1 = torch.randn(B, c1, h, w, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    x2 = torch.randn(B, c2, h * (2 if up1 else 1), w * (2 if up1 else 1), device=dev).contiguous(memory_format=torch.channels_last).requires_grad_() if c2 else None
    wt = (torch.randn(cout, c1 + c2, k, k, device=dev) * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_()
    bias = torch.zeros(cout, device=dev, requires_grad=True)
    ops.PROFILE = []; ops.PROFILE_DESC = []
    for _ in range(4):
        y = ops.conv2d(x1, wt, bias, x2=x2, pad=k // 2, pad_mode=A.PAD_REFLECT if reflect else A.PAD_ZERO, up1=up1, act=A.ACT_ELU)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    rows = {}
    for (kind, fl, a, b) in ops.PROFILE[len(ops.PROFILE) // 4:]:
        rows.setdefault(kind, []).append((a.elapsed_time(b), fl))
    ops.PROFILE = None; ops.PROFILE_DESC = None
    print(name, " ".join("%s %.2f ms %.0f TF/s" % (k_, sum(t for t, _ in v) / len(v), sum(f for _, f in v) / sum(t for t, _ in v) / 1e9) for k_, v in rows.items()))

if which in ("all", "conv"):
    conv_case(64, 0, 64, 3, 256, 512, True, True, "upconv(0,1) 64->64 @512x1024 up")
    conv_case(128, 64, 128, 3, 128, 256, True, True, "upconv(1,1) 128+64->128 @256x512 up")
    conv_case(128, 0, 64, 3, 256, 512, True, False, "upconv(0,0) 128->64 @256x512")
    conv_case(256, 512, 256, 3, 32, 64, True, True, "upconv(3,1) 256+512->256 @64x128 up")
    conv_case(256, 0, 64, 1, 128, 256, False, False, "layer1 1x1 256->64 @128x256")
    conv_case(64, 0, 64, 3, 128, 256, False, False, "layer1 3x3 64->64 @128x256")
    conv_case(2048, 0, 256, 3, 32, 64, False, False, "aspp 3x3 2048->256 @32x64")
