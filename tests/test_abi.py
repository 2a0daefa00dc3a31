"""CPU: the C-ABI library builds/loads and exports every symbol the header declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "segsde_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(segsde_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 40
    for must in ("segsde_reproj_fused", "segsde_conv2d_fwd", "segsde_conv2d_fwd_tc", "segsde_bn_stats", "segsde_ce_fwd"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        from improving_segmentation_with_selfsupervised_depth_b200.csrc.build import build
        build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.segsde_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.segsde_version()


def test_argument_errors_do_not_need_a_gpu():
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    lib = A.lib()
    assert lib.segsde_reproj_fused(None, None) == -1
    assert lib.segsde_conv2d_fwd(None, None, None, None, None, None, None) == -1
    assert b"invalid argument" in lib.segsde_error_string(-1)
    assert lib.segsde_reproj_tiles(512, 1024) == 37 * 16      # 28-column strips x 32-row bands


def test_product_path_refuses_cpu_tensors():
    import pytest
    import torch
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    with pytest.raises(A.SegsdeError):
        ops.conv2d(torch.zeros(1, 4, 8, 8), torch.zeros(4, 4, 3, 3))


def _struct_fields(name):
    """Field names of a `typedef struct { ... } name;` in the header, in order."""
    src = open(os.path.join(ROOT, "include", "segsde_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = {m.group(2): m.group(1) for m in re.finditer(r"typedef struct \{([^{}]*)\}\s*(\w+);", src)}[name]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.sub(r"\[.*?\]", "", part.strip().split()[-1].lstrip("*")))
    return fields


def test_ctypes_structs_mirror_the_header():
    """The ctypes mirrors in _cabi.py carry the header's fields in the header's order (a drifted struct would
    silently shift every later argument of a kernel call)."""
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    for cname, cls in (("segsde_nhwc_t", A.NHWC), ("segsde_conv_desc_t", A.ConvDesc), ("segsde_reproj_args_t", A.ReprojArgs)):
        assert [f[0] for f in cls._fields_] == _struct_fields(cname), cname
    assert ctypes.sizeof(A.ConvDesc) == 4 * len(A.ConvDesc._fields_)


def test_zero_pools_hand_out_disjoint_zeroed_slices():
    """ops.zeros_f32 / zeros_f64 (gradient and statistics accumulators): aligned, disjoint, zero, right layout."""
    import torch
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    dev = torch.device("cpu")
    a, b = ops.zeros_f32(100, dev), ops.zeros_f32(7, dev)
    assert a.numel() == 100 and b.numel() == 7 and float(a.abs().sum()) == 0.0
    assert (b.data_ptr() - a.data_ptr()) % 256 == 0 and b.data_ptr() >= a.data_ptr() + 400
    a += 1.0
    assert float(b.abs().sum()) == 0.0
    big = ops.zeros_f32(ops._ZPOOL32_CHUNK, dev)          # larger than a chunk share: its own allocation
    assert big.numel() == ops._ZPOOL32_CHUNK and float(big[:1000].abs().sum()) == 0.0
    w = torch.empty(8, 4, 3, 3).contiguous(memory_format=torch.channels_last)
    g = ops.zeros_like_w(w)
    assert g.shape == w.shape and g.stride() == w.stride() and float(g.abs().sum()) == 0.0
    d = ops.zeros_f64(5, dev)
    assert d.dtype == torch.float64 and d.numel() == 5 and d.data_ptr() % 16 == 0
    # a chunk boundary: fill the current chunk, the next request must come from a fresh zeroed chunk
    rest = ops._ZPOOL32[dev][1]
    filler = ops.zeros_f32(max(1, (ops._ZPOOL32_CHUNK - rest) // 4 - 64), dev)
    filler += 3.0
    for _ in range(4):
        t = ops.zeros_f32(ops._ZPOOL32_CHUNK // 4 - 64, dev)
        assert float(t[:4096].abs().sum()) == 0.0
        t += 1.0
