"""Where does the end-to-end step time go: H2D bandwidth, CPU enqueue time per step, GPU time per step."""
import contextlib, io, sys, time
sys.path.insert(0, '/root/repo')
import torch
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200.synthetic import MONO_LOSS_KW, mono_config, synthetic_inputs
B, H, W = 12, 512, 1024
models, loss = P.install_dropin()
dev = torch.device('cuda')
with contextlib.redirect_stdout(io.StringIO()):
    model = models.get_model(mono_config('resnet50', H, W), 19).to(dev).train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-4, fused=True)
ml = loss.MonodepthLoss(height=H, width=W, batch_size=B, **MONO_LOSS_KW)
host = {k: v.pin_memory() for k, v in synthetic_inputs(B, H, W).items()}
print("pinned:", all(v.is_pinned() for v in host.values()), "bytes", sum(v.numel() * v.element_size() for v in host.values()))
inputs = {k: v.to(dev) for k, v in host.items()}
def step(inp):
    opt.zero_grad(set_to_none=True)
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inp)
    ml.generate_images_pred(inp, out)
    l = ml.compute_losses(inp, out)['loss']
    l.backward()
    opt.step()
    return l
for _ in range(3): step(inputs)
torch.cuda.synchronize()
# H2D bandwidth
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e0.record(); bufs = {k: v.to(dev, non_blocking=True) for k, v in host.items()}; e1.record(); torch.cuda.synchronize()
    print("H2D ms", e0.elapsed_time(e1))
# CPU enqueue vs GPU time
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    l = step(inputs)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("cpu enqueue %.1f ms, gpu %.1f ms, wall %.1f ms" % ((t1 - t0) * 1e3, e0.elapsed_time(e1), (t2 - t0) * 1e3))
# finer: forward / loss / backward / opt enqueue times
torch.cuda.synchronize(); t = [time.perf_counter()]
opt.zero_grad(set_to_none=True)
with contextlib.redirect_stdout(io.StringIO()):
    out = model(inputs)
t.append(time.perf_counter())
ml.generate_images_pred(inputs, out); l = ml.compute_losses(inputs, out)['loss']; t.append(time.perf_counter())
l.backward(); t.append(time.perf_counter())
opt.step(); t.append(time.perf_counter())
print("enqueue ms: fwd %.1f loss %.1f bwd %.1f opt %.1f" % tuple((b - a) * 1e3 for a, b in zip(t[:-1], t[1:])))
