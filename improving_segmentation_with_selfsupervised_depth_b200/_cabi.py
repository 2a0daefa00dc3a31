"""ctypes binding of libsegsde_b200.so (the C-ABI declared in include/segsde_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, a Python
exception is raised (reference convention: plain exceptions, e.g. models/__init__.py:23).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsegsde_b200.so")


class SegsdeError(RuntimeError):
    pass


class NHWC(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


class ConvDesc(C.Structure):
    _fields_ = [("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("dil", C.c_int32), ("pad_mode", C.c_int32), ("up1", C.c_int32), ("act", C.c_int32),
                ("nchw_norm_in", C.c_int32), ("stride_w", C.c_int32)]


class ReprojArgs(C.Structure):
    _fields_ = [("tgt", C.c_void_p), ("src", C.c_void_p * 2), ("K", C.c_void_p), ("inv_K", C.c_void_p),
                ("T", C.c_void_p * 2), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("F", C.c_int32),
                ("S", C.c_int32), ("disp", C.c_void_p * 4), ("hs", C.c_int32 * 4), ("ws", C.c_int32 * 4),
                ("noise", C.c_void_p * 4), ("seed", C.c_uint64), ("offset", C.c_uint64),
                ("min_depth", C.c_float), ("max_depth", C.c_float), ("flags", C.c_int32),
                ("loss_partial", C.c_void_p), ("ident_sel", C.c_void_p * 4), ("gdisp", C.c_void_p * 4),
                ("gT_partial", C.c_void_p)]


ACT_NONE, ACT_RELU, ACT_ELU, ACT_SIGMOID = 0, 1, 2, 3
PAD_ZERO, PAD_REFLECT = 0, 1
REPROJ_NO_SSIM, REPROJ_AVG, REPROJ_NO_AUTOMASK = 1, 2, 4
REPROJ_MAX_SCALES = 4
E_UNSUPPORTED = -3

_lib = None


def lib():
    """Loads the shared library once. Raises SegsdeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SegsdeError(
                "libsegsde_b200.so not found at %s — run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.segsde_version.restype = C.c_char_p
        _lib.segsde_error_string.restype = C.c_char_p
        _lib.segsde_launch_count.restype = C.c_int64
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().segsde_error_string(C.c_int(code)).decode()
        raise SegsdeError("%s failed (%d): %s" % (what, code, msg))


# when a list, every entry-point call is bracketed by CUDA events: (name, ev0, ev1) — profiling scripts only
PROFILE = None
PROFILE_NAMES = None      # optional set: only these entry points are bracketed


def _invoke(name, args):
    fn = getattr(lib(), name)
    if PROFILE is None or (PROFILE_NAMES is not None and name not in PROFILE_NAMES):
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    code = fn(*args)
    e1.record()
    PROFILE.append((name, e0, e1))
    return code


def call(name, *args):
    """Calls an int-returning entry point and raises on a non-zero code."""
    code = _invoke(name, args)
    if code != 0:
        check(code, name)


def try_call(name, *args):
    """Like call() but returns False on SEGSDE_E_UNSUPPORTED (shape outside a kernel family)."""
    code = _invoke(name, args)
    if code == E_UNSUPPORTED:
        return False
    if code != 0:
        check(code, name)
    return True


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def require_cuda(*ts):
    """Every kernel launches on the CURRENT device and stream: tensors must live there (a model on another GPU of
    the same process must be driven under `torch.cuda.device(...)`)."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise SegsdeError("segsde_b200 ops run on CUDA tensors only (got a %s tensor); there is "
                              "no CPU fallback" % t.device)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise SegsdeError("tensor on %s but the current CUDA device is %d: wrap the call in "
                              "torch.cuda.device(tensor.device)" % (t.device, cur))


def default_seed(salt):
    """Seed of the in-kernel Philox streams (dropout masks, auto-mask tie-break noise) when the caller gives none:
    derived from torch's global seed (`torch.manual_seed`) and the data-parallel rank, so that seeded runs are
    reproducible and ranks / runs with different seeds draw different masks."""
    rank = int(os.environ.get("RANK", "0"))
    return ((int(torch.initial_seed()) ^ int(salt)) * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF


def launch_count():
    return int(lib().segsde_launch_count())
