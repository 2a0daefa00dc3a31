"""One profiled training step of the bench configuration: per-entry-point and per-conv breakdown."""
import contextlib, io, sys, collections
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
from improving_segmentation_with_selfsupervised_depth_b200.synthetic import MONO_LOSS_KW, mono_config, synthetic_inputs
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 12, 512, 1024
models, loss = P.install_dropin()
dev = torch.device('cuda')
with contextlib.redirect_stdout(io.StringIO()):
    model = models.get_model(mono_config('resnet50', H, W), 19).to(dev).train()
params = [p for p in model.parameters() if p.requires_grad]
from improving_segmentation_with_selfsupervised_depth_b200 import optim as _so
opt = _so.Adam(params, lr=1e-4)
ml = loss.MonodepthLoss(height=H, width=W, batch_size=B, **MONO_LOSS_KW)
inputs = {k: v.to(dev) for k, v in synthetic_inputs(B, H, W).items()}
def step():
    opt.zero_grad(set_to_none=True)
    with contextlib.redirect_stdout(io.StringIO()):
        out = model(inputs)
    ml.generate_images_pred(inputs, out)
    l = ml.compute_losses(inputs, out)['loss']
    l.backward()
    opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print('unprofiled step ms', e0.elapsed_time(e1))
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize()
print('host issue time of one step (queue empty, no sync inside): %.1f ms' % ((t1 - t0) * 1e3))
A.PROFILE = []; ops.PROFILE = []; ops.PROFILE_DESC = []; ops.ROUTES = []
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print('profiled step ms', e0.elapsed_time(e1))
agg = collections.defaultdict(lambda: [0.0, 0])
for name, a, b in A.PROFILE:
    agg[name][0] += a.elapsed_time(b); agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print('--- by entry point (ms, calls) total %.1f' % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-34s %8.2f %5d' % (k, v[0], v[1]))
print('--- convs (kind, desc, ms, TF/s)')
rows = []
routes = [r for _, r in ops.ROUTES] if len(ops.ROUTES) == len(ops.PROFILE) else [''] * len(ops.PROFILE)
for (kind, fl, a, b), d, rt in zip(ops.PROFILE, ops.PROFILE_DESC, routes):
    ms = a.elapsed_time(b); rows.append((ms, kind, d, fl / (ms * 1e-3 + 1e-12) / 1e12, rt))
byroute = collections.defaultdict(lambda: [0.0, 0.0, 0])
for ms, kind, d, tf, rt in rows:
    byroute[(kind, rt)][0] += ms; byroute[(kind, rt)][1] += tf * ms; byroute[(kind, rt)][2] += 1
print('--- by (kind, route): ms, launches, TF/s')
for k, v in sorted(byroute.items(), key=lambda kv: -kv[1][0]):
    print('%-8s %-14s %7.2f ms %4d  %7.1f TF/s' % (k[0], k[1], v[0], v[2], v[1] / (v[0] + 1e-9)))
for ms, kind, d, tf, rt in sorted(rows, key=lambda r: -r[0])[:80]:
    print('%7.2f ms %-6s %-12s %-44s %7.1f TF/s' % (ms, kind, rt, d, tf))
