"""Likelihood-only parity: fcest_wishart_loglik (C ABI) against the CPU oracle's closed-form and autograd
gradients, across the kernel variants (warp-per-matrix D <= 55 with every row-block count, the
CTA-per-time-step shared-memory kernel, ragged sample counts, vector A modes).

Tolerances: ve 1e-10 relative-to-max (it feeds the 1e-8 ELBO bar), gradients 1e-8 relative-to-max.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import wishart_oracle as wo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(N, D, S, seed, vector_a=False):
    rng = np.random.default_rng(seed)
    R = D * D
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
    fmean = t(0.3 * rng.standard_normal((N, R)))
    fvar = t(0.05 + rng.random((N, R)))
    y = t(rng.standard_normal((N, D)))
    A = t(1.0 + 0.1 * rng.standard_normal(D)) if vector_a else t(np.eye(D) + 0.2 * rng.standard_normal((D, D)))
    lam = t(wo.inv_softplus(0.01) + 0.3 * rng.standard_normal(D))
    eps = t(rng.standard_normal((S, N, R)))
    return fmean, fvar, y, A, lam, eps


def _rel(a, b):
    a = a.detach().cpu().numpy().reshape(-1)
    b = b.detach().cpu().numpy().reshape(-1)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _check(N, D, S, seed=0, vector_a=False):
    from fcest_b200 import ops
    fmean, fvar, y, A, lam, eps = _inputs(N, D, S, seed, vector_a)
    ve_ref = wo.variational_expectations(fmean, fvar, y, A, lam, eps)
    mu_bar, v_bar, Abar, lam_bar = wo.likelihood_closed_form_grad(fmean, fvar, y, A, lam, eps)
    fm_d, fv_d = ops.as_pitched(fmean.cuda()), ops.as_pitched(fvar.cuda())
    out = ops.wishart_loglik(fm_d, fv_d, eps.cuda(), y.cuda(), A.cuda(), lam.cuda(), need_grad=True)
    assert int(out["info"].max().item()) == 0
    assert _rel(out["ve"], ve_ref) < 1e-10, ("ve", D, S)
    assert _rel(out["g_fmean"], mu_bar) < 1e-8, ("g_fmean", D, S)
    assert _rel(out["g_fvar"], v_bar) < 1e-8, ("g_fvar", D, S)
    assert _rel(out["gA"], Abar) < 1e-8, ("gA", D, S)
    assert _rel(out["glam"], lam_bar) < 1e-8, ("glam", D, S)
    fwd = ops.wishart_loglik(fm_d, fv_d, eps.cuda(), y.cuda(), A.cuda(), lam.cuda(), need_grad=False)
    assert _rel(fwd["ve"], ve_ref) < 1e-10


# D -> row blocks of the warp kernel: 1..7 -> 1, 8..15 -> 2, ..., 48..55 -> 7; D >= 56 -> tiles kernel
@pytest.mark.parametrize("D", [1, 2, 3, 7, 8, 9, 15, 16, 23, 24, 31, 32, 39, 40, 47, 48, 50, 55])
def test_loglik_all_row_block_counts(cuda, D):
    _check(N=5 if D > 30 else 9, D=D, S=3, seed=D)


# shared-memory-tiled CTA-per-matrix kernel (56 <= D <= 207): every capacity (10, 14, 18, 22, 26 row blocks),
# D a multiple of 8 (bordered row opens a new block), odd nu (scalar loads); D = 208 falls to the L2-scratch kernel
@pytest.mark.parametrize("D", [56, 63, 64, 79, 80, 103, 111, 112, 143, 144, 175, 176, 200, 207, 208])
def test_loglik_tiles_kernel(cuda, D):
    _check(N=3, D=D, S=2, seed=D)


def test_loglik_tiles_many_time_steps_and_samples(cuda):
    """More time steps than CTAs and several samples: per-CTA A_bar / lam_bar partials, U1 / U2 accumulation."""
    _check(N=160, D=56, S=3, seed=9)
    _check(N=4, D=72, S=5, seed=10, vector_a=True)


@pytest.mark.parametrize("S", [1, 2, 5, 8, 9, 13, 16])
def test_loglik_sample_chunks(cuda, S):
    """S < 8: fewer warps per CTA; S > 8: several sample chunks accumulated in fixed order."""
    _check(N=7, D=12, S=S, seed=100 + S)
    _check(N=3, D=50, S=S, seed=200 + S)


@pytest.mark.parametrize("D", [4, 20, 50])
def test_loglik_vector_A(cuda, D):
    _check(N=6, D=D, S=4, seed=7, vector_a=True)


def test_loglik_many_time_steps(cuda):
    """More time steps than persistent CTAs: every CTA walks several time steps."""
    _check(N=700, D=10, S=2, seed=3)


def test_loglik_bad_pivot_reported(cuda):
    from fcest_b200 import ops
    N, D, S = 4, 20, 2
    fmean, fvar, y, A, lam, eps = _inputs(N, D, S, 5)
    lam = torch.full((D,), -800.0, dtype=torch.float64)  # softplus -> 0
    fmean = torch.zeros_like(fmean)
    eps = torch.zeros_like(eps)                          # G = 0 -> Sigma = 0: first pivot fails
    out = ops.wishart_loglik(ops.as_pitched(fmean.cuda()), ops.as_pitched(fvar.cuda()), eps.cuda(), y.cuda(),
                             A.cuda(), lam.cuda(), need_grad=False)
    assert (out["info"].cpu().numpy() == 1).all()
    assert torch.isnan(out["ve"]).all()


def test_warp_and_cta_kernels_agree(cuda):
    """FCEST_LIK_IMPL=cta forces the CTA-per-time-step kernel; both variants must agree to rounding."""
    code = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r)\n"
        "from fcest_b200 import ops\n"
        "import importlib.util as iu; sp = iu.spec_from_file_location('tl', %r); tl = iu.module_from_spec(sp); sp.loader.exec_module(tl); _inputs = tl._inputs\n"
        "fm, fv, y, A, lam, eps = _inputs(6, 50, 8, 11)\n"
        "o = ops.wishart_loglik(ops.as_pitched(fm.cuda()), ops.as_pitched(fv.cuda()), eps.cuda(), y.cuda(), A.cuda(), lam.cuda())\n"
        "np.savez(sys.argv[1], **{k: v.cpu().numpy() for k, v in o.items()})\n" % (ROOT, os.path.abspath(__file__)))
    outs = []
    for impl in ("warp", "cta"):
        path = os.path.join(ROOT, "gpurun_out", f"_lik_{impl}.npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        env = dict(os.environ, FCEST_LIK_IMPL=impl)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, cwd=ROOT)
        outs.append(np.load(path))
    for k in ("ve", "g_fmean", "g_fvar", "gA", "glam"):
        a, b = outs[0][k], outs[1][k]
        assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max(), k


# D above the single-SM tile store (208 <= D <= 400; the reference takes any D, likelihoods.py:185): cluster-tiled kernel
# (likelihood_cluster.cu), one size per instantiated row-block capacity (30 / 34 / 38: CTA pairs, 42 / 46 / 51: clusters of 4)
@pytest.mark.parametrize("D", [232, 256, 300, 320, 360, 400])
def test_loglik_large_D(cuda, D):
    _check(N=2, D=D, S=2, seed=D)


def test_loglik_large_D_odd_size_and_vector_A(cuda):
    """Cluster-tiled kernel with an odd D (odd row pitch of the noise: no bulk L2 prefetch, padded G slot rows for the tensor
    map) and with the vector form of A; three time steps on at most two clusters' worth of CTAs."""
    _check(N=3, D=233, S=3, seed=233, vector_a=True)
    _check(N=2, D=305, S=2, seed=305)


def test_large_D_prediction_and_pd_check(cuda):
    """predict moments and the batched PD check at D = 256 and 400 against the oracle."""
    from fcest_b200 import ops
    from fcest_b200.helpers.array_operations import are_all_positive_definite
    for D, n, S in ((256, 2, 3), (400, 1, 2)):
        rng = np.random.default_rng(D)
        t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
        fm, fv = t(0.3 * rng.standard_normal((n, D * D))), t(0.05 + rng.random((n, D * D)))
        A, lam = t(np.eye(D) + 0.1 * rng.standard_normal((D, D))), t(wo.inv_softplus(0.01) + 0.1 * rng.standard_normal(D))
        eps = t(rng.standard_normal((S, n, D, D)))
        for corr, fn in ((0, wo.predict_cov), (1, wo.predict_corr)):
            mean_ref, std_ref = fn(fm, fv, A, lam, eps, D)
            dev = "cuda"
            mean = torch.empty((n, D, D), dtype=torch.float64, device=dev)
            m2 = torch.empty_like(mean); out_mean = torch.empty_like(mean); out_std = torch.empty_like(mean)
            from fcest_b200 import _lib
            nscr = int(_lib.load_library().fcest_cov_moments_scratch_bytes(D, D))
            scratch = torch.empty(max(nscr // 8, 1), dtype=torch.float64, device=dev)
            fmd, fvd = ops.as_pitched(fm.cuda()), ops.as_pitched(fv.cuda())
            eps_d, A_d, lam_d = eps.reshape(S, n, D * D).cuda().contiguous(), A.cuda(), lam.cuda()   # keep alive
            ops.call("fcest_cov_moments", ops.ptr(fmd), ops.ptr(fvd), ops.ld(fmd), ops.ptr(eps_d),
                     ops.ptr(A_d), 0, ops.ptr(lam_d), S, n, D, D, corr, 0, ops.ptr(mean), ops.ptr(m2), None, 1,
                     ops.ptr(out_mean), ops.ptr(out_std), ops.ptr(scratch))
            torch.cuda.synchronize()
            assert _rel(out_mean, mean_ref) < 1e-9 and _rel(out_std, std_ref) < 1e-8, (D, corr)
            if not corr:
                assert are_all_positive_definite(out_mean)
                bad = out_mean.clone()
                bad[0] -= 2.0 * torch.linalg.eigvalsh(bad[0].cpu()).max().item() * torch.eye(D, dtype=torch.float64, device=dev)
                assert not are_all_positive_definite(bad)


def test_kernels_are_run_to_run_deterministic(cuda):
    """No atomics on the data path and fixed-order reductions: five runs of the likelihood (warp and tile kernels),
    of the per-latent triangular kernels (shared work counter) and of the window covariance (two-buffer bulk store)
    give bit-identical outputs -- the race evidence available here (compute-sanitizer is closed on this pool)."""
    from fcest_b200 import ops
    from fcest_b200.models import SlidingWindows
    for D, N, S in ((50, 40, 8), (72, 6, 3), (256, 3, 2)):
        fmean, fvar, y, A, lam, eps = _inputs(N, D, S, 3)
        args = (ops.as_pitched(fmean.cuda()), ops.as_pitched(fvar.cuda()), eps.cuda(), y.cuda(), A.cuda(), lam.cuda())
        ref = ops.wishart_loglik(*args)
        for _ in range(4):
            out = ops.wishart_loglik(*args)
            for k in ("ve", "g_fmean", "g_fvar", "gA", "glam"):
                assert torch.equal(out[k], ref[k]), (D, k)
    g = torch.Generator().manual_seed(0)
    R, M = 300, 200
    ldk = (M * (M + 1) // 2 + 15) // 16 * 16
    q = torch.tril(torch.randn(R, M, M, dtype=torch.float64, generator=g)).cuda().contiguous()
    Gk = torch.randn(R, ldk, dtype=torch.float64, generator=g).cuda()
    outs = []
    for _ in range(5):
        Sk = torch.empty((R, ldk), dtype=torch.float64, device="cuda")
        klp = torch.empty((R, 2), dtype=torch.float64, device="cuda")
        dq = torch.empty_like(q)
        ops.call("fcest_tri_syrk_pack", ops.ptr(q), R, M, ops.ptr(Sk), ldk, ops.ptr(klp))
        ops.call("fcest_tri_grad", ops.ptr(Gk), ldk, ops.ptr(q), R, M, ops.ptr(dq), 1.0)
        outs.append((Sk[:, :M * (M + 1) // 2].clone(), klp, dq))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))
    y = np.random.default_rng(1).standard_normal((1200, 100))
    sw = SlidingWindows(np.linspace(0, 1, 1200)[:, None], y)
    first = sw.overlapping_windowed_cov_estimation(30)
    for _ in range(4):
        assert np.array_equal(sw.overlapping_windowed_cov_estimation(30), first)


@pytest.mark.parametrize("D,S,vector_a", [(1, 1, False), (2, 5, False), (3, 5, False), (4, 3, True), (5, 2, False), (6, 9, False), (7, 4, True)])
def test_loglik_subwarp_kernel(cuda, D, S, vector_a):
    """D <= 7: four problems per warp (8-lane groups, registers + shuffles, likelihood_subwarp.cu); also more time steps
    than one block of groups and a sample count that is not a multiple of anything."""
    _check(N=37, D=D, S=S, seed=300 + D, vector_a=vector_a)


def test_loglik_subwarp_bad_pivot_and_forced_warp_kernel_agree(cuda):
    from fcest_b200 import ops
    N, D, S = 5, 3, 2
    fmean, fvar, y, A, lam, eps = _inputs(N, D, S, 5)
    lam0 = torch.full((D,), -800.0, dtype=torch.float64)
    out = ops.wishart_loglik(ops.as_pitched(torch.zeros_like(fmean).cuda()), ops.as_pitched(fvar.cuda()),
                             torch.zeros_like(eps).cuda(), y.cuda(), A.cuda(), lam0.cuda(), need_grad=False)
    assert (out["info"].cpu().numpy() == 1).all() and torch.isnan(out["ve"]).all()
    code = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r)\n"
        "from fcest_b200 import ops\n"
        "import importlib.util as iu; sp = iu.spec_from_file_location('tl', %r); tl = iu.module_from_spec(sp); sp.loader.exec_module(tl)\n"
        "fm, fv, y, A, lam, eps = tl._inputs(21, 6, 4, 12)\n"
        "o = ops.wishart_loglik(ops.as_pitched(fm.cuda()), ops.as_pitched(fv.cuda()), eps.cuda(), y.cuda(), A.cuda(), lam.cuda())\n"
        "np.savez(sys.argv[1], **{k: v.cpu().numpy() for k, v in o.items()})\n" % (ROOT, os.path.abspath(__file__)))
    outs = []
    for impl in ("subwarp", "warp"):
        path = os.path.join(ROOT, "gpurun_out", f"_lik_small_{impl}.npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=dict(os.environ, FCEST_LIK_IMPL=impl), cwd=ROOT)
        outs.append(np.load(path))
    for k in ("ve", "g_fmean", "g_fvar", "gA", "glam"):
        a, b = outs[0][k], outs[1][k]
        assert np.abs(a - b).max() <= 1e-11 * max(np.abs(b).max(), 1e-300), k
