"""Forward-only timing of encoder-type convolutions (B=12), with and without BatchNorm statistics in the epilogue."""
import sys, os
sys.path.insert(0, '/root/repo')
import torch
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
B = 12
dev = torch.device('cuda')
ops.USE_TC = True


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cases = [  # cin, cout, k, stride, dil, h_in, w_in, stats
    (64, 256, 1, 1, 1, 128, 256, True), (64, 256, 1, 1, 1, 128, 256, False), (256, 64, 1, 1, 1, 128, 256, True),
    (64, 64, 3, 1, 1, 128, 256, True), (64, 64, 1, 1, 1, 128, 256, True),
    (256, 128, 1, 1, 1, 128, 256, True), (128, 128, 3, 2, 1, 128, 256, True), (128, 512, 1, 1, 1, 64, 128, True),
    (512, 128, 1, 1, 1, 64, 128, True), (128, 128, 3, 1, 1, 64, 128, True),
    (256, 256, 3, 1, 1, 32, 64, True), (256, 1024, 1, 1, 1, 32, 64, True), (1024, 256, 1, 1, 1, 32, 64, True),
    (512, 2048, 1, 1, 1, 32, 64, True), (2048, 512, 1, 1, 1, 32, 64, True), (512, 512, 3, 1, 2, 32, 64, True),
    (2048, 256, 3, 1, 12, 32, 64, True), (2048, 256, 1, 1, 1, 32, 64, True),
]
tot = 0.0
with torch.no_grad():
    for cin, cout, k, s, d, h, w, st in cases:
        x = torch.randn(B, cin, h, w, device=dev).contiguous(memory_format=torch.channels_last)
        wt = (torch.randn(cout, cin, k, k, device=dev) * 0.05).contiguous(memory_format=torch.channels_last)
        sums = torch.zeros(3 * cout, device=dev, dtype=torch.float64) if st else None
        pad = d * (k // 2)
        ms = timeit(lambda: ops.conv2d(x, wt, None, stride=s, pad=pad, dil=d, bn_stats=sums))
        ho, wo = (h + 2 * pad - d * (k - 1) - 1) // s + 1, (w + 2 * pad - d * (k - 1) - 1) // s + 1
        fl = 2.0 * B * ho * wo * cout * cin * k * k
        by = 4.0 * B * (h * w * cin + ho * wo * cout)
        tot += ms
        print("%4d->%4d k%d s%d d%-2d in %3dx%3d stats=%d : %.3f ms  %6.1f TF/s  %6.0f GB/s" % (cin, cout, k, s, d, h, w, st, ms, fl / ms / 1e9, by / ms / 1e6))
print("total %.3f ms" % tot)
