import sys, os
sys.path.insert(0, '/root/repo')
import torch
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
dev = torch.device('cuda'); B = 12
def run(c1, cout, h, w, act, bias):
    x = torch.randn(B, c1, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, c1, 3, 3, device=dev) * 0.05).contiguous(memory_format=torch.channels_last)
    bb = torch.zeros(cout, device=dev) if bias else None
    for _ in range(3): y = ops.conv2d(x, wt, bb, pad=1, act=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y = ops.conv2d(x, wt, bb, pad=1, act=act)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * h * w * cout * 9 * c1
    print("dbg=%s %d->%d %dx%d act=%d bias=%d: %.3f ms %.0f TF/s" % (os.environ.get("SEGSDE_TC_DBG", "0"), c1, cout, h, w, act, bias, ms, fl / ms / 1e9))
run(64, 64, 512, 1024, 2, True)
run(64, 64, 512, 1024, 0, False)
run(128, 128, 256, 512, 2, True)
