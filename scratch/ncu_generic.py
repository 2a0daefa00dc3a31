"""A few launches of the generic tcgen05 kernel on encoder / ASPP shapes (B=12) for an ncu --set full capture."""
import sys
sys.path.insert(0, '/root/repo')
import torch
import improving_segmentation_with_selfsupervised_depth_b200 as P
from improving_segmentation_with_selfsupervised_depth_b200 import ops
B, dev = 12, torch.device('cuda')
ops.USE_TC = True
cases = [(512, 512, 3, 2, 32, 64), (2048, 256, 3, 12, 32, 64), (64, 256, 1, 1, 128, 256), (256, 64, 1, 1, 128, 256),
         (512, 2048, 1, 1, 32, 64), (256, 256, 3, 1, 32, 64)]
with torch.no_grad():
    for cin, cout, k, d, h, w in cases:
        x = torch.randn(B, cin, h, w, device=dev).contiguous(memory_format=torch.channels_last)
        wt = (torch.randn(cout, cin, k, k, device=dev) * 0.05).contiguous(memory_format=torch.channels_last)
        sums = torch.zeros(3 * cout, device=dev, dtype=torch.float64)
        for _ in range(2):
            ops.conv2d(x, wt, None, stride=1, pad=d * (k // 2), dil=d, bn_stats=sums)
torch.cuda.synchronize()
print("done")
