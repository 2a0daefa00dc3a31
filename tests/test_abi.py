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
    assert lib.segsde_reproj_tiles(512, 1024) == 32 * 64


def test_product_path_refuses_cpu_tensors():
    import pytest
    import torch
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    with pytest.raises(A.SegsdeError):
        ops.conv2d(torch.zeros(1, 4, 8, 8), torch.zeros(4, 4, 3, 3))
