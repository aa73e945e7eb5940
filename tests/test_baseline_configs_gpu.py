"""Oracle parity AT THE SIZES BASELINE.json NAMES (configs 1-4), through the public classes and the C ABI.

C1  VariationalWishartProcess N=200 D=3 (README dummy data), ELBO + gradient + predict_cov / predict_corr (S*=300)
C2  SparseVariationalWishartProcess N=1200 D=15 M=200 S=3, ELBO + gradient
C3  SlidingWindows N=1200 D=100 window_length=30 and the cross-validated window search, against a fixture the
    REFERENCE ITSELF produced at this size (tests/golden/sliding_c3.npz, make_golden.py sliding_c3)
C4  SparseVariationalWishartProcess N=1200 D=50 S=8 (M=200), ELBO + gradient -- the metric's configuration

Tolerances (north_star): ELBO 1e-8 relative, gradients 1e-7 relative-to-max, predicted covariances / correlations
1e-6, sliding-window estimates 1e-12 relative-to-max.  The oracle evaluations take 0.1 s (C1), 1 s (C2), ~10 s (C4).
"""
import os

import numpy as np
import pytest
import torch

from oracle import wishart_oracle as wo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _assign(m, p):
    m.q_mu.assign(p["q_mu"])
    m.q_sqrt.assign(p["q_sqrt"])
    m.likelihood.A_scale_matrix.assign(p["A"])
    m.likelihood.additive_part.assign(p["lam"])


def _check_grads(m, g_ref, sparse, tol=1e-7, floor=0.0):
    """Each gradient to ``tol`` relative to its largest reference entry (``floor``: lower bound of that scale, for a
    gradient that is analytically zero at the evaluation point)."""
    got = {"q_mu": m.q_mu.grad, "q_sqrt": m.q_sqrt.grad, "A": m.likelihood.A_scale_matrix.grad,
           "lam": m.likelihood.additive_part.grad, "u_var": m.kernel.variance.grad,
           "u_ls": m.kernel.lengthscales.grad}
    if sparse:
        got["Z"] = m.inducing_variable.Z.grad
    worst = {}
    for k, v in got.items():
        ref = g_ref[k].numpy().reshape(-1)
        worst[k] = float(np.abs(v.cpu().numpy().reshape(-1) - ref).max() / max(np.abs(ref).max(), floor, 1e-300))
        assert worst[k] < tol, (k, worst[k])
    return worst


def test_c1_vwp_readme_config(cuda):
    """reference README.md:32-57: N=200, D=3, x = linspace(0, 1, N), y = random((N, D)); S=5 (the default)."""
    from fcest_b200.models import VariationalWishartProcess
    N, D, S = 200, 3, 5
    x, y, eps, p = wo.make_inputs(N, D, S, M=None, seed=1)
    y = torch.tensor(np.random.default_rng(1).random((N, D)))   # README: np.random.random, not centred
    m = VariationalWishartProcess(x.numpy(), y.numpy(), nu=D)
    assert m.likelihood.num_mc_samples == 5
    _assign(m, p)
    elbo_ref, g_ref = wo.elbo_and_grad("vgp", p, x, y, eps)
    elbo = m.elbo_and_grad(None, epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) <= 1e-8 * abs(elbo_ref), (elbo, elbo_ref)
    _check_grads(m, g_ref, sparse=False)
    # predict_cov / predict_corr with the README's default 300 samples at the training locations
    fm_ref, fv_ref = wo.vgp_predict_f(x, x, p)
    e2 = torch.randn(300, N, D, D, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    for name in ("predict_cov", "predict_corr"):
        mean_ref, std_ref = getattr(wo, name)(fm_ref, fv_ref, p["A"], p["lam"], e2, D)
        mean, std = getattr(m, name)(x.numpy(), 300, epsilon=e2.cuda())
        assert mean.shape == (N, D, D) and std.shape == (N, D, D)
        assert (mean.cpu() - mean_ref).abs().max().item() <= 1e-6 * max(1.0, mean_ref.abs().max().item()), name
        assert (std.cpu() - std_ref).abs().max().item() <= 1e-6 * max(1.0, std_ref.abs().max().item()), name


def test_c1_vwp_at_reference_initialisation(cuda):
    """The same config at the point the reference starts from (A = I, q_mu = 0, q_sqrt = 1e-3 I)."""
    from fcest_b200.models import VariationalWishartProcess
    N, D, S = 200, 3, 5
    x, y, eps, p = wo.make_inputs(N, D, S, M=None, seed=2, perturbed=False)
    m = VariationalWishartProcess(x.numpy(), y.numpy(), nu=D)
    elbo_ref, g_ref = wo.elbo_and_grad("vgp", p, x, y, eps)
    elbo = m.elbo_and_grad(None, epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) <= 1e-8 * abs(elbo_ref), (elbo, elbo_ref)
    # with q_mu = 0 and q_sqrt = c I the VGP moments do not depend on the lengthscale: d ELBO / d u_ls is zero
    # analytically (oracle: -6e-17, rounding), so it is held to the scale of the kernel-variance gradient
    _check_grads(m, g_ref, sparse=False, floor=float(g_ref["u_var"].abs().max()))


def test_c2_svwp(cuda):
    from fcest_b200.models import SparseVariationalWishartProcess
    N, D, S, M = 1200, 15, 3, 200
    x, y, eps, p = wo.make_inputs(N, D, S, M=M, seed=0)
    m = SparseVariationalWishartProcess(D, p["Z"].numpy(), num_mc_samples=S, verbose=False)
    _assign(m, p)
    elbo_ref, g_ref = wo.svgp_elbo_and_grad_chunked(p, x, y, eps, chunk=75)
    elbo = m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) <= 1e-8 * abs(elbo_ref), (elbo, elbo_ref)
    _check_grads(m, g_ref, sparse=True)
    assert abs(m.elbo((x.numpy(), y.numpy()), epsilon=eps.cuda()).item() - elbo_ref) <= 1e-8 * abs(elbo_ref)


def test_c4_svwp_metric_config(cuda):
    """The configuration the metric is quoted on: one full oracle evaluation (autograd through the materialised
    GPflow op sequence, 125 latents at a time) against one CUDA evaluation on the same epsilon."""
    from fcest_b200.models import SparseVariationalWishartProcess
    N, D, S, M = 1200, 50, 8, 200
    x, y, eps, p = wo.make_inputs(N, D, S, M=M, seed=0)
    m = SparseVariationalWishartProcess(D, p["Z"].numpy(), num_mc_samples=S, verbose=False)
    _assign(m, p)
    elbo_ref, g_ref = wo.svgp_elbo_and_grad_chunked(p, x, y, eps, chunk=125)
    elbo = m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) <= 1e-8 * abs(elbo_ref), (elbo, elbo_ref)
    _check_grads(m, g_ref, sparse=True)


# ---- C3 against the reference's own outputs at full size -----------------------------------------------------------
def _c3():
    g = np.load(os.path.join(GOLD, "sliding_c3.npz"))
    y = np.random.default_rng(int(g["seed"])).standard_normal((1200, 100))
    return g, y


def test_c3_window_cov_matches_reference(cuda):
    from fcest_b200.models import SlidingWindows
    g, y = _c3()
    sw = SlidingWindows(np.linspace(0, 1, 1200)[:, None], y)
    cov = sw.overlapping_windowed_cov_estimation(30)
    assert cov.shape == (1200, 100, 100)
    ref = g["cov_rows"]
    assert np.abs(cov[g["rows"]] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    sums = np.array([cov.sum(), np.abs(cov).sum(), (cov * cov).sum()])
    assert np.all(np.abs(sums - g["cov_checksum"]) <= 1e-11 * np.abs(g["cov_checksum"]).max())


def test_c3_cross_validated_window_search_matches_reference(cuda):
    """Full 1019 x 80 table: the -inf pattern (window lengths <= 101 are rank deficient at D=100) is exact, finite
    scores agree to 1e-6 relative, and the selected window length is the reference's."""
    from fcest_b200.models import SlidingWindows
    g, y = _c3()
    sw = SlidingWindows(np.linspace(0, 1, 1200)[:, None], y)
    df = sw.find_cross_validated_optimal_window_length()
    assert np.array_equal(df.index.to_numpy(), g["locs"]) and np.array_equal(df.columns.to_numpy(), g["windows"])
    got, ref = df.to_numpy(), g["table"]
    assert got.shape == ref.shape == (1019, 80)
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert fin.any() and np.all(np.abs(got[fin] - ref[fin]) <= 1e-6 * np.abs(ref[fin]))
    assert sw.get_optimal_window_length(df) == int(g["best"])
    assert sw.compute_cross_validated_optimal_window_length() == int(g["best"])
