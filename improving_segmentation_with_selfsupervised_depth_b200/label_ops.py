"""Label-selection scoring ops (label_selection.py:347-650) on the sm_100a kernels: adaptive
pooling of the depth features, the pairwise feature-distance matrix (`torch.cdist`) and the iterative-farthest-point
selection as ONE launch (the reference runs two reductions and an `.item()` per selected sample)."""
import ctypes as C

import torch

from . import _cabi as A


def adaptive_pool2d(x, output_size, mode="avg"):
    """F.adaptive_avg_pool2d / F.adaptive_max_pool2d (label_selection.py:398-403) for 3-D (C,H,W) or 4-D input."""
    A.require_cuda(x)
    if mode not in ("avg", "max"):
        raise NotImplementedError(mode)
    squeeze = x.dim() == 3
    x4 = (x.unsqueeze(0) if squeeze else x).detach().float().contiguous()
    n, c, h, w = x4.shape
    oh, ow = int(output_size[0]), int(output_size[1])
    y = torch.empty(n, c, oh, ow, device=x.device, dtype=torch.float32)
    A.call("segsde_adaptive_pool", A.ptr(x4), C.c_int(n * c), C.c_int(h), C.c_int(w), C.c_int(oh), C.c_int(ow),
           C.c_int(int(mode == "max")), A.ptr(y), A.stream_ptr())
    return y.squeeze(0) if squeeze else y


def calc_feature_distance(features, bias=None, bias_weight=0.0, p=2, normalize_features=False):
    """`_calc_feature_distance` (label_selection.py:573-614, the non-patch-wise branch every shipped config uses):
    features: list of 1 x C x H x W tensors (or one N x C x H x W tensor) -> N x N distance matrix, zero diagonal."""
    f = torch.cat(list(features)) if isinstance(features, (list, tuple)) else features
    A.require_cuda(f)
    f = f.detach().float()
    n = f.shape[0]
    if normalize_features:      # host-side statistics like the reference (torch.std_mean over N, H, W)
        std, mean = torch.std_mean(f, dim=[0, 2, 3], keepdim=True)
        f = (f - mean) / std
    f = f.reshape(n, -1).contiguous()
    out = torch.empty(n, n, device=f.device, dtype=torch.float32)
    A.call("segsde_pairwise_distance", A.ptr(f), C.c_int(n), C.c_int64(f.shape[1]), C.c_float(float(p)), A.ptr(out),
           A.stream_ptr())
    if bias_weight > 0:
        if bias is None or len(bias) != n:
            raise ValueError("calc_feature_distance: one bias per sample expected")
        out += torch.as_tensor([float(b) for b in bias], device=out.device, dtype=torch.float32)
    out.fill_diagonal_(0)
    return out


def iterative_farthest_point(current_samples, feature_distances, n_new, preselected_samples=None):
    """label_selection.py:617-640 with the same arguments and return value (new sample ids, their distances)."""
    dist = feature_distances["distances"]
    to_img, to_i = feature_distances["dist_i_to_img_idx"], feature_distances["img_idx_to_dist_i"]
    A.require_cuda(dist)
    dist = dist.detach().float().clone()
    n = dist.shape[0]
    if preselected_samples is not None:
        keep = torch.zeros(n, dtype=torch.bool, device=dist.device)
        keep[[to_i[s] for s in preselected_samples]] = True
        dist[:, ~keep] = 0
    cur = torch.zeros(n, dtype=torch.int32, device=dist.device)
    cur[[to_i[s] for s in current_samples]] = 1
    new_idx = torch.empty(max(n_new, 1), dtype=torch.int64, device=dist.device)
    new_dist = torch.empty(max(n_new, 1), dtype=torch.float32, device=dist.device)
    count = torch.zeros(1, dtype=torch.int32, device=dist.device)
    scratch = torch.empty(n, dtype=torch.float32, device=dist.device)
    A.call("segsde_farthest_point", A.ptr(dist.contiguous()), C.c_int(n), A.ptr(cur), C.c_int(n_new), A.ptr(new_idx),
           A.ptr(new_dist), A.ptr(count), A.ptr(scratch), A.stream_ptr())
    k = int(count)
    idx = new_idx[:k].tolist()
    return [to_img[i] for i in idx], [new_dist[i] for i in range(k)]
