"""CPU tests (no GPU): the oracle against the reference's golden vectors and against itself."""
import os

import numpy as np
import pytest
import torch

from oracle import filter_oracle as fo
from oracle import sliding_oracle as so
from oracle import summary_oracle as smo
from oracle import wishart_oracle as wo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- sliding windows: oracle vs outputs of the reference itself (tests/golden/make_golden.py) -------
def test_sliding_oracle_matches_reference_fixtures():
    g = np.load(os.path.join(GOLD, "sliding_windows.npz"))
    for tag in "abcd":
        y = g[f"y_{tag}"]
        w, step = (int(v) for v in g[f"meta_{tag}"])
        cov = so.overlapping_windowed_cov(y, w, step)
        corr = so.overlapping_windowed_cov(y, w, step, "correlation")
        assert cov.shape == g[f"cov_{tag}"].shape
        scale = max(1.0, np.abs(g[f"cov_{tag}"]).max())
        assert np.abs(cov - g[f"cov_{tag}"]).max() <= 1e-12 * scale
        assert np.abs(corr - g[f"corr_{tag}"]).max() <= 1e-12
        # rows >= N/step stay zero, result index is the window counter (sliding_windows.py:164-169)
        assert np.all(cov[int(y.shape[0] / step):] == 0)


def test_cv_oracle_matches_reference_fixtures():
    g = np.load(os.path.join(GOLD, "sliding_cv.npz"))
    assert tuple(g["bounds_1200"]) == (21, 181)  # SURVEY.md finding 2
    for tag in "ab":
        y = g[f"y_{tag}"]
        assert so.cv_bounds(y.shape[0]) == tuple(int(v) for v in g[f"bounds_{tag}"])
        windows, locs, table = so.cross_validated_table(y)
        assert np.array_equal(windows, g[f"windows_{tag}"]) and np.array_equal(locs, g[f"locs_{tag}"])
        ref = g[f"table_{tag}"]
        assert np.array_equal(np.isinf(table), np.isinf(ref))
        fin = np.isfinite(ref)
        # w - 2 == D windows are barely full rank (cond ~ 1e10): eigh-driver differences show up at 1e-9
        assert np.all(np.abs(table[fin] - ref[fin]) <= 1e-7 * np.abs(ref[fin]))
        assert so.optimal_window_length(windows, table) == int(g[f"best_{tag}"])
    assert np.isinf(g["table_b"]).any()  # D=25: short windows are rank deficient -> -inf (finding 3)


def test_static_fc_known_answers():
    """The reference's only exact golden vectors on this path
    (tests/fcest/models/test_sliding_windows.py:25-44): np.cov on 3x2 integer data."""
    y = np.array([[1.0, 2.0], [3.0, 5.0], [8.0, 9.0]])
    c = so.np_cov_rows(y)
    assert np.allclose(c, np.cov(y.T), rtol=0, atol=1e-15)
    r = so.correlation_from_covariance(np.eye(3))
    assert np.array_equal(r, np.eye(3))  # unit variance: cov == corr (test_array_operations.py:67-88)


def test_mvn_logpdf_restatement_matches_scipy():
    from scipy.stats import multivariate_normal
    rng = np.random.default_rng(0)
    for d, n in [(3, 10), (6, 5), (4, 30)]:
        rows = rng.standard_normal((n, d))
        cov = so.np_cov_rows(rows)
        x = rng.standard_normal(d)
        a = so.mvn_logpdf_allow_singular(x, cov)
        b = multivariate_normal.logpdf(x, mean=np.zeros(d), cov=cov, allow_singular=True)
        assert (np.isinf(a) and np.isinf(b)) or abs(a - b) <= 1e-10 * abs(b)


# ---- Wishart oracle: regression fixtures, identities, gradient routes -------------------------------
@pytest.mark.parametrize("tag,kind", [("svgp", "svgp"), ("vgp", "vgp")])
def test_wishart_oracle_regression_fixture(tag, kind):
    g = np.load(os.path.join(GOLD, "wishart_oracle.npz"))
    N, D, S, M = (int(v) for v in g[f"{tag}_shape"])
    x, y, eps, p = wo.make_inputs(N, D, S, M=None if M < 0 else M, seed=5)
    elbo, grads = wo.elbo_and_grad(kind, p, x, y, eps)
    assert abs(elbo - float(g[f"{tag}_elbo"])) <= 1e-12 * abs(elbo)
    for k, v in grads.items():
        ref = g[f"{tag}_grad_{k}"]
        assert np.abs(v.numpy() - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-30), k


def test_reference_likelihood_identities():
    """tests/fcest/models/test_likelihoods.py:54-76 (matmul form, no additive noise):
    L L^T = Sigma, 2 sum log diag L = logdet Sigma, |L^-1 y|^2 = y^T Sigma^-1 y."""
    rng = np.random.default_rng(1)
    S, N, D, nu = 2, 6, 2, 3
    F = torch.tensor(rng.random((S, N, D, nu)))
    A = torch.tensor(rng.random((D, D)))
    y = torch.tensor(rng.random((N, D)))
    AF = A @ F
    Sig = AF @ AF.transpose(-1, -2)
    L = torch.linalg.cholesky(Sig)
    assert torch.allclose(L @ L.transpose(-1, -2), Sig, atol=1e-12)
    assert torch.allclose(2 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1), torch.logdet(Sig), atol=1e-10)
    z = torch.linalg.solve_triangular(L, y[None, :, :, None].expand(S, -1, -1, -1), upper=False)
    q1 = (z ** 2).sum((2, 3))
    q2 = (y[None, :, None, :] @ torch.linalg.inv(Sig) @ y[None, :, :, None])[..., 0, 0]
    assert torch.allclose(q1, q2, rtol=1e-8)
    # and the oracle's log_prob with a_matmul=True equals the dense MVN density built from those pieces
    lam = torch.full((D,), -50.0, dtype=torch.float64)
    lp = wo.log_prob(F, y, A, lam, a_matmul=True)
    Sig2 = Sig + torch.diag_embed(torch.nn.functional.softplus(lam).expand(S, N, D))
    ref = (-0.5 * D * np.log(2 * np.pi) - 0.5 * torch.logdet(Sig2)
           - 0.5 * (y[None, :, None, :] @ torch.linalg.inv(Sig2) @ y[None, :, :, None])[..., 0, 0]).mean(0)
    assert torch.allclose(lp, ref, rtol=1e-9)


def test_closed_form_likelihood_gradient_matches_autograd():
    x, y, eps, p = wo.make_inputs(20, 4, 3, M=8, seed=2)
    fm, fv = wo.svgp_predict_f(x, p)
    fm = fm.clone().requires_grad_(); fv = fv.clone().requires_grad_()
    A = p["A"].clone().requires_grad_(); lam = p["lam"].clone().requires_grad_()
    wo.variational_expectations(fm, fv, y, A, lam, eps).sum().backward()
    cf = wo.likelihood_closed_form_grad(fm.detach(), fv.detach(), y, A.detach(), lam.detach(), eps)
    for a, b in zip(cf, (fm.grad, fv.grad, A.grad, lam.grad)):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-12


def test_oracle_gradient_finite_differences():
    x, y, eps, p = wo.make_inputs(12, 3, 2, M=6, seed=4)
    _, g = wo.elbo_and_grad("svgp", p, x, y, eps)
    h = 1e-6
    for name, idx in [("q_mu", (2, 4)), ("q_sqrt", (3, 4, 1)), ("A", (1, 2)), ("lam", (1,)), ("Z", (2, 0)),
                      ("u_ls", (0,)), ("u_var", (0,))]:
        pp = {k: v.clone() for k, v in p.items()}
        pm = {k: v.clone() for k, v in p.items()}
        pp[name][idx] += h
        pm[name][idx] -= h
        fd = (float(wo.svgp_elbo(pp, x, y, eps)) - float(wo.svgp_elbo(pm, x, y, eps))) / (2 * h)
        an = float(g[name][idx])
        assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (name, fd, an)


def test_chunked_oracle_equals_full():
    x, y, eps, p = wo.make_inputs(30, 3, 3, M=10, seed=1)
    e1, g1 = wo.elbo_and_grad("svgp", p, x, y, eps)
    e2, g2 = wo.svgp_elbo_and_grad_chunked(p, x, y, eps, chunk=4)
    assert abs(e1 - e2) <= 1e-12 * abs(e1)
    for k in g1:
        assert float((g1[k] - g2[k]).abs().max()) <= 1e-9 * max(float(g1[k].abs().max()), 1.0)


# ------------------------------------------------------------------------------------------------
# Independent anchors for the GPflow restatement (the oracle is "parity unpinned": TF / GPflow cannot be
# installed here).  Each test derives the same quantity by a different route, from third-party code or
# textbook formulas that do not share the restatement's structure.
# ------------------------------------------------------------------------------------------------
def test_matern52_matches_sklearn():
    """gpflow.kernels.Matern52 restatement == sigma^2 * sklearn Matern(nu=2.5) (independent implementation)."""
    from sklearn.gaussian_process.kernels import Matern
    rng = np.random.default_rng(0)
    X, X2 = rng.random((17, 1)), rng.random((9, 1))
    var, ls = 1.7, 0.3
    ref = var * Matern(length_scale=ls, nu=2.5)(X, X2)
    got = wo.matern52(torch.tensor(X), torch.tensor(X2), torch.tensor(var, dtype=torch.float64),
                      torch.tensor(ls, dtype=torch.float64)).numpy()
    assert np.abs(got - ref).max() < 1e-13


def test_whitened_conditional_matches_unwhitened_svgp_algebra():
    """base_conditional(white=True) == the textbook sparse-GP predictive for q(u) = N(m, S) with
    m = Lm q_mu, S = Lm q_sqrt q_sqrt^T Lm^T:  mean = Kxz Kzz^-1 m,
    var = kxx - diag(Kxz Kzz^-1 Kzx) + diag(Kxz Kzz^-1 S Kzz^-1 Kzx)   (Hensman et al. 2013, eq. for q(f))."""
    rng = np.random.default_rng(1)
    M, N, R = 11, 23, 4
    Z, X = np.sort(rng.random((M, 1)), 0), rng.random((N, 1))
    var, ls = 1.3, 0.4
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
    Kzz = wo.matern52(t(Z), t(Z), t(var), t(ls)).numpy() + 1e-6 * np.eye(M)
    Kzx = wo.matern52(t(Z), t(X), t(var), t(ls)).numpy()
    q_mu = rng.standard_normal((M, R))
    q_sqrt = np.tril(rng.standard_normal((R, M, M))) * 0.3 + np.eye(M)[None]
    fmean, fvar = wo.base_conditional_white(t(Kzx), t(Kzz), t(var * np.ones(N)), t(q_mu), t(q_sqrt))
    Lm = np.linalg.cholesky(Kzz)
    KiK = np.linalg.solve(Kzz, Kzx)                       # Kzz^-1 Kzx
    for r in range(R):
        m = Lm @ q_mu[:, r]
        S = Lm @ q_sqrt[r] @ q_sqrt[r].T @ Lm.T
        mean_ref = KiK.T @ m
        var_ref = var - np.einsum("mn,mn->n", Kzx, KiK) + np.einsum("mn,mk,kn->n", KiK, S, KiK)
        assert np.abs(fmean[:, r].numpy() - mean_ref).max() < 1e-9
        assert np.abs(fvar[:, r].numpy() - var_ref).max() < 1e-9


def test_gauss_kl_white_matches_generic_gaussian_kl():
    """gauss_kl(q_mu, q_sqrt, K=None) == sum_r KL(N(mu_r, L_r L_r^T) || N(0, I)) by the generic formula."""
    rng = np.random.default_rng(2)
    M, R = 7, 5
    q_mu = rng.standard_normal((M, R))
    q_sqrt = np.tril(rng.standard_normal((R, M, M))) * 0.2 + np.eye(M)[None]
    got = float(wo.gauss_kl_white(torch.tensor(q_mu), torch.tensor(q_sqrt)))
    ref = 0.0
    for r in range(R):
        S = q_sqrt[r] @ q_sqrt[r].T
        ref += 0.5 * (np.trace(S) + q_mu[:, r] @ q_mu[:, r] - M - np.linalg.slogdet(S)[1])
    assert abs(got - ref) < 1e-11 * max(1.0, abs(ref))


def test_vgp_moments_match_dense_gaussian_algebra():
    """VGP.elbo moments: f = L v with v ~ N(q_mu, q_sqrt q_sqrt^T)  =>  mean = L q_mu, var = diag(L S L^T)."""
    rng = np.random.default_rng(3)
    N, R = 9, 3
    x = np.sort(rng.random((N, 1)), 0)
    p = {"u_var": torch.tensor(wo.inv_softplus(1.2), dtype=torch.float64),
         "u_ls": torch.tensor(wo.inv_softplus(0.35), dtype=torch.float64),
         "q_mu": torch.tensor(rng.standard_normal((N, R))),
         "q_sqrt": torch.tensor(np.tril(rng.standard_normal((R, N, N))) * 0.3 + np.eye(N)[None])}
    fmean, fvar = wo.vgp_fmean_fvar(torch.tensor(x), p)
    K = wo.matern52(torch.tensor(x), torch.tensor(x), torch.tensor(1.2, dtype=torch.float64),
                    torch.tensor(0.35, dtype=torch.float64)).numpy() + 1e-6 * np.eye(N)
    L = np.linalg.cholesky(K)
    for r in range(R):
        S = p["q_sqrt"][r].numpy() @ p["q_sqrt"][r].numpy().T
        assert np.abs(fmean[:, r].numpy() - L @ p["q_mu"][:, r].numpy()).max() < 1e-12
        assert np.abs(fvar[:, r].numpy() - np.diag(L @ S @ L.T)).max() < 1e-12


def test_keras_adam_equals_torch_adam_at_zero_epsilon():
    """With epsilon = 0 the Keras form (bias correction folded into the step size) and torch.optim.Adam coincide."""
    g = torch.Generator().manual_seed(0)
    theta0 = torch.randn(50, dtype=torch.float64, generator=g)
    w = torch.nn.Parameter(theta0.clone())
    opt = torch.optim.Adam([w], lr=1e-3, betas=(0.9, 0.999), eps=0.0)
    theta, m, v = theta0.clone(), torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64)
    for t in range(1, 6):
        grad = torch.randn(50, dtype=torch.float64, generator=g)
        w.grad = grad.clone()
        opt.step()
        theta, m, v = wo.keras_adam_step(theta, grad, m, v, t, eps=0.0)
        assert (theta - w.detach()).abs().max() < 1e-15


def test_filter_oracle_matches_reference_fixtures_bit_exact():
    """oracle/filter_oracle.py against outputs of the reference's own highpass_filter_data (and its filter design)."""
    g = np.load(os.path.join(GOLD, "filtering.npz"))
    for tag in "abcde":
        w, tr = int(g[f"meta_{tag}"][0]), float(g[f"meta_{tag}"][1])
        b, a = fo.butter_highpass(1 / (w * tr), 1 / tr / 2)
        assert np.array_equal(b, g[f"b_{tag}"]) and np.array_equal(a, g[f"a_{tag}"])
        out = fo.highpass_filter_data(g[f"y_{tag}"], w, tr)
        assert np.array_equal(out, g[f"filtered_{tag}"]), tag
        if f"cov_{tag}" in g:
            cov = so.overlapping_windowed_cov(out, w, 1)
            assert np.abs(cov - g[f"cov_{tag}"]).max() <= 1e-13 * max(1.0, np.abs(cov).max())


def test_host_filter_design_is_bit_identical_to_scipy():
    """The product's host-side design (NumPy only) against scipy.signal.butter / lfilter_zi: no GPU involved."""
    from scipy.signal import butter, lfilter_zi
    from fcest_b200.helpers.filtering import _butter_highpass, _compute_lower_frequency_cutoff, _lfilter_zi
    for w, tr in [(30, 1.0), (15, 2.0), (121, 0.72), (181, 1.0), (7, 1.0), (21, 0.8), (3, 1.0)]:
        cutoff = _compute_lower_frequency_cutoff(w, tr)
        assert cutoff == 1 / (w * tr)
        for order in (5, 2, 3):
            b, a = _butter_highpass(cutoff, nyquist_freq=1 / tr / 2, order=order)
            bs, as_ = butter(order, cutoff / (1 / tr / 2), btype="high", analog=False)
            assert np.array_equal(b, bs) and np.array_equal(a, as_) and a[0] == 1.0
            assert np.array_equal(_lfilter_zi(b, a), lfilter_zi(bs, as_))
    with pytest.raises(ValueError):
        _butter_highpass(1.0, nyquist_freq=0.5)   # Wn = 2: scipy rejects it too


def test_summary_oracle_matches_reference_fixtures_bit_exact():
    """oracle/summary_oracle.py against outputs of the reference's summarize_tvfc_estimates; the time-ordered
    sums the kernel uses reproduce numpy's reduction over the leading axis bit for bit."""
    g = np.load(os.path.join(GOLD, "summary.npz"))
    for tag in "ab":
        c = g[f"c_{tag}"]
        for metric in ("mean", "variance", "std", "rate_of_change"):
            with np.errstate(all="ignore"):
                assert np.array_equal(smo.summarize(c, metric), g[f"{metric}_{tag}"], equal_nan=True), (tag, metric)
        s = np.zeros(c.shape[1:]); cnt = np.zeros(c.shape[1:])
        for t in range(c.shape[0]):
            ok = ~np.isnan(c[t]); s = s + np.where(ok, c[t], 0.0); cnt += ok
        assert np.array_equal(s / cnt, g[f"mean_{tag}"], equal_nan=True)
        q = np.zeros(c.shape[1:])
        for t in range(c.shape[0]):
            d = np.where(np.isnan(c[t]), 0.0, c[t] - s / cnt); q = q + d * d
        assert np.array_equal(q / cnt, g[f"variance_{tag}"], equal_nan=True)


def test_ar1_restatement_is_the_centred_ols_slope():
    rng = np.random.default_rng(2)
    c = rng.standard_normal((60, 3, 3)); c = c + np.swapaxes(c, 1, 2)
    ref = smo.ar1(c)
    y = c[:, 2, 1]; x, z = y[:-1], y[1:]
    slope = ((x - x.mean()) * (z - z.mean())).sum() / ((x - x.mean()) ** 2).sum()
    assert abs(ref[2, 1] - slope) < 1e-12 and ref[1, 2] == ref[2, 1] and ref[0, 0] == 0.0


def test_csv_layout_matches_reference_round_trip():
    """load_tvfc_estimates / to_3d_format on a file the reference's save_tvfc_estimates wrote: same array as the
    reference's own loader returns (including its leading row-index slice); to_2d_format inverts to_3d_format on
    symmetric structures."""
    from fcest_b200.helpers.data import to_2d_format, to_3d_format
    from fcest_b200.models import SlidingWindows
    g = np.load(os.path.join(GOLD, "tvfc_csv.npz"))
    got = SlidingWindows.load_tvfc_estimates(GOLD, "tvfc_reference.csv")
    assert got.shape == g["loaded"].shape == (13, 3, 3)
    assert np.array_equal(np.asarray(got, dtype=np.float64), g["loaded"])
    cov = so.overlapping_windowed_cov(g["y"], 4, 1)
    assert np.abs(np.asarray(got, dtype=np.float64)[1:] - cov).max() <= 1e-15    # what was saved is the estimate
    assert np.array_equal(to_3d_format(to_2d_format(cov)), np.swapaxes(cov, 1, 2))
    assert to_2d_format(cov).shape == (9, 12)


# ---- Wishart path: oracle vs outputs of the reference's OWN source (likelihoods.py / wishart_process.py /
#      array_operations.py executed on the NumPy tensorflow stand-in, tests/golden/make_golden.py) -------------
def _ref_cases(g, key):
    return [c.split(":") for c in g[key].tolist()]


def test_wishart_oracle_matches_reference_source_likelihood():
    g = np.load(os.path.join(GOLD, "wishart_reference.npz"))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
    for tag, option in _ref_cases(g, "lik_cases"):
        A, lam, y = t(g[f"lik_{tag}_A"]), t(g[f"lik_{tag}_lam"]), t(g[f"lik_{tag}_y"])
        ve = wo.variational_expectations(t(g[f"lik_{tag}_f_mean"]), t(g[f"lik_{tag}_f_var"]), y, A, lam,
                                         t(g[f"lik_{tag}_eps"]))
        ref = g[f"lik_{tag}_ve"]
        assert np.abs(ve.numpy() - ref).max() <= 1e-11 * np.abs(ref).max(), tag
        lp = wo.log_prob(t(g[f"lik_{tag}_f_sample"]), y, A, lam)
        ref = g[f"lik_{tag}_log_prob"]
        assert np.abs(lp.numpy() - ref).max() <= 1e-11 * np.abs(ref).max(), tag
        noise = wo.add_diagonal_additive_noise(t(g[f"lik_{tag}_cov_in"]), lam)
        assert np.abs(noise.numpy() - g[f"lik_{tag}_cov_noise"]).max() <= 1e-14
        # constructor state (likelihoods.py:64-86, 208-251) against the oracle's canonical initial point
        D = y.shape[1]
        _, _, _, p = wo.make_inputs(4, D, 1, M=3, perturbed=False, a_option=option)
        assert np.array_equal(p["A"].numpy(), g[f"lik_{tag}_A_init"])
        assert np.abs(p["lam"].numpy() - g[f"lik_{tag}_lam_init"]).max() <= 1e-15
        dims = g[f"lik_{tag}_dims"]
        assert tuple(dims[:5]) == (1, D * D, D, D, D)
        assert bool(dims[5]) == (option in ("train_full_matrix", "train_diagonal")) and bool(dims[6])
    assert str(g["nu_lt_D_message"]) == "Wishart Degrees of Freedom must be >= D."


def test_wishart_oracle_matches_reference_source_prediction():
    g = np.load(os.path.join(GOLD, "wishart_reference.npz"))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
    for tag, option in _ref_cases(g, "pred_cases"):
        A, lam = t(g[f"pred_{tag}_A"]), t(g[f"pred_{tag}_lam"])
        fm, fv, eps = t(g[f"pred_{tag}_f_mean"]), t(g[f"pred_{tag}_f_var"]), t(g[f"pred_{tag}_eps"])
        D = lam.shape[0]
        s = wo.cov_samples(fm, fv, A, lam, eps, D)
        assert np.abs(s.numpy() - g[f"pred_{tag}_samples"]).max() <= 1e-12 * np.abs(g[f"pred_{tag}_samples"]).max()
        c = wo.convert_tensor_to_correlation(s)
        assert np.abs(c.numpy() - g[f"pred_{tag}_corr_samples"]).max() <= 1e-13
        for fn, key in ((wo.predict_cov, "cov"), (wo.predict_corr, "corr")):
            mean, std = fn(fm, fv, A, lam, eps, D)
            scale = np.abs(g[f"pred_{tag}_{key}_mean"]).max()
            assert np.abs(mean.numpy() - g[f"pred_{tag}_{key}_mean"]).max() <= 1e-12 * scale, (tag, key)
            assert np.abs(std.numpy() - g[f"pred_{tag}_{key}_std"]).max() <= 1e-12 * scale, (tag, key)
    truth = [wo.are_all_positive_definite(m) for m in g["pd_inputs"]]
    assert truth == g["pd_truth"].tolist() == [True, False, False]
