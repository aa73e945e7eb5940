"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares; the Python binding table agrees with the header; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fcest_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fcest_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ["fcest_wishart_loglik", "fcest_dgemm", "fcest_pack_outer", "fcest_tri_syrk_pack",
                 "fcest_tri_grad", "fcest_sym_matvec", "fcest_matern52", "fcest_matern52_bwd",
                 "fcest_chol_inv_single", "fcest_cov_moments", "fcest_adam_step", "fcest_randn",
                 "fcest_window_cov", "fcest_window_cv"]:
        assert must in names, must


def test_library_exports_every_declared_symbol():
    from fcest_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/fcest_b200.h but not exported"


def test_binding_table_matches_header():
    from fcest_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    _lib.load_library()


def test_no_cpu_fallback_without_gpu():
    import torch
    from fcest_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        ops.call("fcest_axpby", 0, 1.0, None, 0.0, None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fcest_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"


def test_synthetic_inputs_equal_the_checker_inputs():
    """bench.py's CUDA arm builds its inputs with fcest_b200.helpers.synthetic (no oracle import on the product
    side); the CPU arm uses oracle.wishart_oracle.make_inputs.  Same seed -> identical arrays."""
    import numpy as np
    from fcest_b200.helpers.synthetic import make_synthetic
    from oracle import wishart_oracle as wo
    for kw in (dict(N=30, D=4, S=3, M=9, seed=2), dict(N=12, D=3, S=2, M=None, seed=5, a_option="scaled"),
               dict(N=10, D=3, S=2, M=4, seed=1, perturbed=False, a_option="train_diagonal")):
        s = make_synthetic(**kw)
        x, y, eps, p = wo.make_inputs(**kw)
        assert np.array_equal(s["x"], x.numpy()) and np.array_equal(s["y"], y.numpy())
        assert np.array_equal(s["eps"], eps.numpy())
        for k in ("Z", "q_mu", "q_sqrt", "A", "lam"):
            assert np.array_equal(s[k], p[k].numpy()), k


def test_likelihood_scratch_sizes_follow_the_kernel_choice():
    """`fcest_wishart_loglik_scratch_bytes` is a pure size computation (no launch): nothing for the warp-per-matrix kernel
    (D <= 55), sample-sum slots + one G chunk for the shared-memory-tiled kernel (56 .. 207) and for the cluster-tiled
    kernel (208 .. 407).  Without a device the library assumes 148 SMs."""
    from fcest_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    f = lib.fcest_wishart_loglik_scratch_bytes
    f.restype = ctypes.c_longlong
    f.argtypes = [ctypes.c_int, ctypes.c_int]
    assert f(50, 50) == 0 and f(15, 15) == 0
    sms = 148
    for D in (64, 100, 200):                      # tiles kernel: [2 R doubles per CTA][8 samples x sms time steps of G]
        ldgs = (D + 3) // 4 * 4
        assert f(D, D) == 8 * (2 * D * D * sms + 8 * sms * D * ldgs), D
    for D in (232, 256, 400):                     # cluster-tiled kernel: slots for sms / 2 clusters; the L2-scratch kernel's
        ldgs = (D + 3) // 4 * 4                   # buffer (kept for FCEST_LIK_IMPL=cta) is the larger of the two
        need = 8 * (2 * D * D * (sms // 2) + 8 * (sms // 2) * D * ldgs)
        assert f(D, D) >= need, D
    assert f(64, 64) < f(100, 100) < f(200, 200)
