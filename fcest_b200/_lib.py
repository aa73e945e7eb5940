"""ctypes binding of ``csrc/libfcest_b200.so`` (the C ABI declared in ``include/fcest_b200.h``).

There is no CPU fallback: if the shared library is missing or a CUDA device is unavailable every
entry point raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``make -C fcest_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfcest_b200.so")

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_ull = C.c_ulonglong
_d = C.c_double

# name -> (restype, argtypes).  Kept in lock-step with include/fcest_b200.h (tests/test_abi.py checks).
SIGNATURES = {
    "fcest_wishart_loglik_smem_bytes": (_ll, [_i, _i]),
    "fcest_wishart_loglik_scratch_bytes": (_ll, [_i, _i]),
    "fcest_wishart_loglik": (_i, [_p, _p, _ll, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _ll, _p, _p,
                                  _p, _p, _p, _p, _p]),
    "fcest_dgemm_workspace_bytes": (_ll, [_i, _i, _i, _ll, _i]),
    "fcest_dgemm": (_i, [_i, _i, _i, _i, _i, _d, _p, _ll, _ll, _p, _ll, _ll, _d, _p, _ll, _ll, _p, _p, _i,
                         _p, _ll, _p]),
    "fcest_dgemm_ozaki_workspace_bytes": (_ll, [_i, _i, _i, _i]),
    "fcest_dgemm_ozaki": (_i, [_i, _i, _i, _i, _i, _d, _p, _ll, _p, _ll, _p, _ll, _p, _i, _p, _ll, _p]),
    "fcest_dgemm_ozaki_debug": (None, [_p]),
    "fcest_dgemm_ozaki_profile": (_i, [_i]),
    "fcest_dgemm_ozaki_profile_read": (_i, [_p, _i]),
    "fcest_int8_mma_peak": (_ll, [_i, _p]),
    "fcest_pack_outer": (_i, [_p, _ll, _i, _i, _p, _ll, _p, _p, _i, _p]),
    "fcest_tri_syrk_pack": (_i, [_p, _i, _i, _p, _ll, _p, _p]),
    "fcest_tri_grad": (_i, [_p, _ll, _p, _i, _i, _p, _d, _p]),
    "fcest_tri_grad_adam": (_i, [_p, _ll, _p, _i, _i, _d, _p, _p, _p, _d, _d, _d, _p]),
    "fcest_sym_matvec": (_i, [_p, _ll, _p, _ll, _i, _i, _p, _ll, _i, _i, _p, _ll, _p, _i, _p]),
    "fcest_matern52": (_i, [_p, _i, _p, _i, _i, _p, _p, _d, _p, _ll, _p]),
    "fcest_matern52_bwd": (_i, [_p, _i, _p, _i, _i, _p, _p, _p, _ll, _i, _p, _p, _p, _p, _i, _p]),
    "fcest_chol_single": (_i, [_p, _i, _ll, _p, _p]),
    "fcest_tri_inv_single": (_i, [_p, _i, _ll, _p, _ll, _p]),
    "fcest_chol_inv_single": (_i, [_p, _i, _ll, _p, _ll, _p, _p]),
    "fcest_tril": (_i, [_p, _i, _ll, _i, _p]),
    "fcest_transpose": (_i, [_p, _i, _i, _ll, _p, _ll, _p]),
    "fcest_axpby": (_i, [_ll, _d, _p, _d, _p, _p]),
    "fcest_sum": (_i, [_p, _ll, _ll, _d, _p, _i, _p]),
    "fcest_sumsq": (_i, [_p, _i, _i, _ll, _d, _p, _i, _p, _p]),
    "fcest_softplus_fwd": (_i, [_p, _p, _i, _p]),
    "fcest_softplus_bwd": (_i, [_p, _p, _p, _i, _p]),
    "fcest_adam_step": (_i, [_p, _p, _p, _p, _ll, _d, _ll, _d, _d, _d, _d, _p]),
    "fcest_adam_step_dev": (_i, [_p, _p, _p, _p, _ll, _d, _p, _d, _d, _d, _p]),
    "fcest_randn": (_i, [_p, _ll, _ull, _ull, _p]),
    "fcest_cov_moments_scratch_bytes": (_ll, [_i, _i]),
    "fcest_cov_moments": (_i, [_p, _p, _ll, _p, _p, _i, _p, _i, _i, _i, _i, _i, _ll, _p, _p, _p, _i, _p, _p,
                               _p, _p]),
    "fcest_batched_chol_info": (_i, [_p, _i, _i, _d, _p, _p, _p]),
    "fcest_window_cov": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fcest_window_cv": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _d, _p, _ll, _p]),
    "fcest_filtfilt_scratch_bytes": (_ll, [_i, _ll, _i]),
    "fcest_filtfilt": (_i, [_p, _i, _ll, _p, _p, _p, _i, _p, _p, _p]),
    "fcest_tvfc_summary": (_i, [_p, _i, _i, _i, _p, _p]),
    "fcest_launch_count": (_ll, []),
    "fcest_launch_count_add": (_ll, [_ll]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen the C-ABI library and attach signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fcest_b200: CUDA library not found at {LIB_PATH}. There is no CPU fallback; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("fcest_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")


_arg_device = None   # device of the tensors marshalled for the next call()


def ptr(t):
    """Device pointer of a tensor (or NULL).  Also notes the tensor's device: ``call`` launches on THAT device's
    current stream, not on whatever device happens to be current in the calling thread."""
    global _arg_device
    if t is None:
        return None
    assert t.is_cuda and t.dtype in (torch.float64, torch.int32), (t.device, t.dtype)
    _arg_device = t.device   # every operand of one call lives on one device (the kernels take raw pointers)
    return C.c_void_p(t.data_ptr())


def stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_ERR = {-1: "operand alignment (16-byte pointers, even leading dimensions)", -2: "unsupported size",
        -3: "bad argument"}


# Optional per-call device timing (bench.py): {"min_flops": f, "records": [(name, flops, ev0, ev1)]}.
# CUDA events are recorded on the launching stream around the C call; nothing is synchronised here.
PROFILE = None


def call(name: str, *args, flops: float = 0.0):
    """Invoke an ``int``-returning entry point on the current stream and raise on a non-zero status."""
    global _arg_device
    require_cuda()
    fn = getattr(load_library(), name)
    dev, _arg_device = _arg_device, None
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):   # kernels launch in the context of the device that owns the operands
            return _launch(fn, name, args, flops, dev)
    return _launch(fn, name, args, flops, dev)


def _launch(fn, name, args, flops, dev):
    prof = PROFILE
    if prof is not None and flops >= prof["min_flops"]:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.current_stream(dev))
        rc = fn(*args, stream(dev))
        ev1.record(torch.cuda.current_stream(dev))
        prof["records"].append((name, flops, ev0, ev1))
    else:
        rc = fn(*args, stream(dev))
    if rc != 0:
        what = _ERR.get(rc, f"CUDA error {rc}")
        raise RuntimeError(f"{name} failed: {what}")


def launch_count() -> int:
    return int(load_library().fcest_launch_count())
