import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A, ops
torch.manual_seed(0)
def run(N, H, W, c1, co, k, pad, fill=None):
    x = torch.randn(N, c1, H, W, device='cuda') if fill is None else torch.full((N, c1, H, W), fill, device='cuda')
    x = x.contiguous(memory_format=torch.channels_last)
    Ho, Wo = H + 2*pad - (k-1), W + 2*pad - (k-1)
    dz = torch.randn(N, co, Ho, Wo, device='cuda') if fill is None else torch.ones(N, co, Ho, Wo, device='cuda')
    dz = dz.contiguous(memory_format=torch.channels_last)
    dw = torch.zeros(co, c1, k, k, device='cuda').contiguous(memory_format=torch.channels_last)
    dw2 = torch.zeros_like(dw)
    d = ops._desc(k, k, 1, pad, 1, 0, False, 0, False)
    st = A.stream_ptr()
    rc = A.lib().segsde_conv2d_wgrad_tc(C.byref(ops.view(x)), None, C.byref(ops.view(dz)), A.ptr(dw), None, C.byref(d), st)
    A.call("segsde_conv2d_wgrad", C.byref(ops.view(x)), None, C.byref(ops.view(dz)), A.ptr(dw2), None, C.byref(d), st)
    torch.cuda.synchronize()
    print('case', (N,H,W,c1,co,k,pad,fill), 'rc', rc, 'tc absmax', dw.abs().max().item(), 'ref absmax', dw2.abs().max().item(),
          'err', ((dw-dw2).abs().max()/dw2.abs().max()).item())
    print(' tc[0,:6]', dw.permute(0,2,3,1).reshape(co,-1)[0,:6].tolist())
    print(' rf[0,:6]', dw2.permute(0,2,3,1).reshape(co,-1)[0,:6].tolist())
run(1, 1, 32, 32, 32, 1, 0, fill=1.0)
run(1, 1, 32, 32, 32, 1, 0)
run(2, 16, 32, 64, 64, 1, 0)
run(2, 16, 32, 128, 128, 1, 0)
run(2, 16, 32, 64, 128, 3, 1)
