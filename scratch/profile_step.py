"""One profiled training step of a bench configuration: per-entry-point and per-conv (route, layer) breakdown.
    python scratch/profile_step.py [dec5|dec6|joint|depthmix] [batch]"""
import collections
import contextlib
import io
import os
import sys
import time
import types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
name = sys.argv[1] if len(sys.argv) > 1 else "dec5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_BATCH[name]
H, W = 512, 1024
dev = torch.device('cuda')
args = types.SimpleNamespace(config=name, batch=B, height=H, width=W)
model, params, task_step = bench.build_task(args, dev, 1, 0)
inputs = {k: v.to(dev) for k, v in bench.synthetic(B, H, W, 1234, name in ("joint", "depthmix")).items()}
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        task_step(inputs)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print('%s B=%d unprofiled step ms' % (name, B), e0.elapsed_time(e1))
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize()
print('host issue time of one step (queue empty, no sync inside): %.1f ms' % ((t1 - t0) * 1e3))
A.PROFILE = []; ops.PROFILE = []; ops.PROFILE_DESC = []; ops.ROUTES = []
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print('profiled step ms', e0.elapsed_time(e1))
agg = collections.defaultdict(lambda: [0.0, 0])
for nm, a, b in A.PROFILE:
    agg[nm][0] += a.elapsed_time(b); agg[nm][1] += 1
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
layers = collections.defaultdict(lambda: [0.0, 0.0, 0])
for ms, kind, d, tf, rt in rows:
    layers[(kind, rt, d)][0] += ms; layers[(kind, rt, d)][1] += tf * ms; layers[(kind, rt, d)][2] += 1
print('--- layers (same kind/route/shape merged): ms, launches, TF/s')
for k, v in sorted(layers.items(), key=lambda kv: -kv[1][0])[:90]:
    print('%7.2f ms x%-3d %-6s %-12s %-46s %7.1f TF/s' % (v[0], v[2], k[0], k[1], k[2], v[1] / (v[0] + 1e-9)))
