"""CUDA path (through the C ABI) against outputs of the REFERENCE'S OWN SOURCE for the FCEst-owned arithmetic of
the Wishart path: ``likelihoods.py`` (constructor, ``variational_expectations``, ``_log_prob``,
``_add_diagonal_additive_noise``), ``wishart_process.py`` (``_get_cov_samples``, ``predict_cov``, ``predict_corr``,
``save_model_params_dict`` / ``load_from_params_dict``) and ``array_operations.py`` (``convert_tensor_to_correlation``,
``are_all_positive_definite``).  The fixtures were produced by executing those files from the read-only reference
checkout on the NumPy tensorflow stand-in (``tests/golden/tf_numpy_shim.py``, ``make_golden.py wishart_reference``)
with the N(0,1) draws injected.  Tolerances: 1e-10 on log densities (north_star: ELBO 1e-8), 1e-10 on predicted
covariances / correlations (north_star: 1e-6)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cases(g, key):
    return [c.split(":") for c in g[key].tolist()]


def _dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def test_likelihood_matches_reference_source(cuda):
    from fcest_b200 import ops
    from fcest_b200.models.likelihoods import WishartProcessLikelihood
    g = np.load(os.path.join(GOLD, "wishart_reference.npz"))
    for tag, option in _cases(g, "lik_cases"):
        y = g[f"lik_{tag}_y"]
        N, D = y.shape
        S = g[f"lik_{tag}_eps"].shape[0]
        lik = WishartProcessLikelihood(D, num_mc_samples=S, A_scale_matrix_option=option, train_additive_noise=True,
                                       verbose=False)
        # constructor state: likelihoods.py:64-86 and :208-251
        assert np.array_equal(lik.A_scale_matrix.numpy(), g[f"lik_{tag}_A_init"]), tag
        assert np.abs(lik.additive_part.numpy() - g[f"lik_{tag}_lam_init"]).max() <= 1e-15
        dims = g[f"lik_{tag}_dims"]
        assert (lik.input_dim, lik.latent_dim, lik.observation_dim, lik.D, lik.nu) == tuple(dims[:5])
        assert (lik.A_scale_matrix.trainable, lik.additive_part.trainable) == (bool(dims[5]), bool(dims[6]))
        lik.A_scale_matrix.assign(g[f"lik_{tag}_A"])
        lik.additive_part.assign(g[f"lik_{tag}_lam"])
        fm = ops.as_pitched(_dev(g[f"lik_{tag}_f_mean"]))
        fv = ops.as_pitched(_dev(g[f"lik_{tag}_f_var"]))
        ve = lik.variational_expectations(None, fm, fv, _dev(y), epsilon=_dev(g[f"lik_{tag}_eps"]))
        ref = g[f"lik_{tag}_ve"]
        assert np.abs(ve.cpu().numpy() - ref).max() <= 1e-10 * np.abs(ref).max(), tag
        lp = lik._log_prob(None, _dev(g[f"lik_{tag}_f_sample"]), y)
        ref = g[f"lik_{tag}_log_prob"]
        assert np.abs(lp.cpu().numpy() - ref).max() <= 1e-10 * np.abs(ref).max(), tag
        noise = lik._add_diagonal_additive_noise(_dev(g[f"lik_{tag}_cov_in"]))
        assert np.abs(noise.cpu().numpy() - g[f"lik_{tag}_cov_noise"]).max() <= 1e-14
    with pytest.raises(Exception, match=str(g["nu_lt_D_message"])):
        WishartProcessLikelihood(4, nu=3, verbose=False)


def test_prediction_matches_reference_source(cuda):
    """predict_cov / predict_corr / _get_cov_samples with GPflow's predict_f replaced by the prescribed (mean,
    variance) pair the reference run used -- exactly how the fixture was produced."""
    from fcest_b200 import ops
    from fcest_b200.models import SparseVariationalWishartProcess, VariationalWishartProcess
    g = np.load(os.path.join(GOLD, "wishart_reference.npz"))
    for tag, option in _cases(g, "pred_cases"):
        eps = g[f"pred_{tag}_eps"]
        S, N, D, _ = eps.shape
        if tag.startswith("svwp"):
            m = SparseVariationalWishartProcess(D, np.linspace(0, 1, 3)[:, None], A_scale_matrix_option=option,
                                                verbose=False)
        else:
            m = VariationalWishartProcess(np.linspace(0, 1, 4)[:, None], np.zeros((4, D)),
                                          A_scale_matrix_option=option)
        m.likelihood.A_scale_matrix.assign(g[f"pred_{tag}_A"])
        m.likelihood.additive_part.assign(g[f"pred_{tag}_lam"])
        fm = ops.as_pitched(_dev(g[f"pred_{tag}_f_mean"]))
        fv = ops.as_pitched(_dev(g[f"pred_{tag}_f_var"]))
        m.predict_f = lambda x_new, fm=fm, fv=fv: (fm, fv)
        x_new = np.linspace(0, 1, N)[:, None]
        smp = m._get_cov_samples(x_new, S, epsilon=_dev(eps))
        ref = g[f"pred_{tag}_samples"]
        assert np.abs(smp.cpu().numpy() - ref).max() <= 1e-11 * np.abs(ref).max(), tag
        cm, cs = m.predict_cov(x_new, S, epsilon=_dev(eps))
        rm, rs = m.predict_corr(x_new, S, epsilon=_dev(eps))
        for got, key in ((cm, "cov_mean"), (cs, "cov_std"), (rm, "corr_mean"), (rs, "corr_std")):
            ref = g[f"pred_{tag}_{key}"]
            assert got.shape == ref.shape
            assert np.abs(got.cpu().numpy() - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (tag, key)


def test_pd_check_matches_reference_source(cuda):
    from fcest_b200.helpers.array_operations import are_all_positive_definite
    g = np.load(os.path.join(GOLD, "wishart_reference.npz"))
    got = [bool(are_all_positive_definite(_dev(m))) for m in g["pd_inputs"]]
    assert got == g["pd_truth"].tolist()


# ---- JSON checkpoints (wishart_process.py:239-308, 526-602) ---------------------------------------------------
def _fresh(kind, D, M, option="train_full_matrix"):
    from fcest_b200.models import SparseVariationalWishartProcess, VariationalWishartProcess
    if kind == "svwp":
        return SparseVariationalWishartProcess(D, np.linspace(0, 1, M)[:, None], A_scale_matrix_option=option,
                                               verbose=False)
    return VariationalWishartProcess(np.linspace(0, 1, M)[:, None], np.zeros((M, D)), A_scale_matrix_option=option)


@pytest.mark.parametrize("key,kind,D,M", [("svwp", "svwp", 2, 3), ("svwp_diag", "svwp", 3, 2), ("vwp", "vwp", 2, 4)])
def test_load_checkpoint_written_by_reference(cuda, tmp_path, key, kind, D, M):
    """A file written by the reference's save_model_params_dict loads into this implementation with the values
    the reference's own load_from_params_dict assigns (incl. 1-D A -> diag(A) for SVWP, :579-581); saving it
    again reproduces the reference's schema, key order and numbers."""
    g = np.load(os.path.join(GOLD, "wishart_reference.npz"))
    fname = f"{key}_reference_params.json"
    m = _fresh(kind, D, M)
    m.load_from_params_dict(GOLD, fname)
    assert np.array_equal(m.likelihood.A_scale_matrix.numpy(), g[f"json_{key}_A"])
    assert np.array_equal(m.likelihood.additive_part.numpy(), g[f"json_{key}_lam"])   # stored unconstrained
    assert abs(float(m.kernel.variance.numpy()) - float(g[f"json_{key}_kvar"])) <= 1e-15
    assert abs(float(m.kernel.lengthscales.numpy()) - float(g[f"json_{key}_kls"])) <= 1e-15
    assert np.array_equal(m.q_mu.numpy(), g[f"json_{key}_q_mu"])
    assert np.array_equal(m.q_sqrt.numpy(), g[f"json_{key}_q_sqrt"])
    if kind == "svwp":
        assert np.array_equal(m.inducing_variable.Z.numpy(), g[f"json_{key}_Z"])
    m.save_model_params_dict(str(tmp_path / "ckpt"), "again.json")   # creates the directory like the reference
    ours = json.load(open(tmp_path / "ckpt" / "again.json"))
    ref = json.load(open(os.path.join(GOLD, fname)))
    assert list(ours.keys()) == list(ref.keys())
    for k in ref:
        a, b = np.asarray(ours[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64)
        if key == "svwp_diag" and k == "A_scale_matrix":
            b = b * np.eye(len(b))   # the loaded model holds the squared form
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 1e-15 * max(1.0, np.abs(b).max()), k
    assert type(ours["D"]) is int and type(ours["kernel_variance"]) is float


@pytest.mark.parametrize("kind", ["svwp", "vwp"])
def test_checkpoint_round_trip_reproduces_elbo_and_prediction(cuda, tmp_path, kind):
    from oracle import wishart_oracle as wo
    N, D, S, M = 30, 4, 3, 9
    x, y, eps, p = wo.make_inputs(N, D, S, M=M if kind == "svwp" else None, seed=3)
    Mq = M if kind == "svwp" else N

    def build():
        from fcest_b200.models import SparseVariationalWishartProcess, VariationalWishartProcess
        if kind == "svwp":
            return SparseVariationalWishartProcess(D, np.linspace(0, 1, Mq)[:, None], num_mc_samples=S, verbose=False)
        return VariationalWishartProcess(x.numpy(), y.numpy(), num_mc_samples=S)

    a = build()
    a.q_mu.assign(p["q_mu"]); a.q_sqrt.assign(p["q_sqrt"])
    a.likelihood.A_scale_matrix.assign(p["A"]); a.likelihood.additive_part.assign(p["lam"])
    a.kernel.variance.assign(1.7); a.kernel.lengthscales.assign(0.21)
    if kind == "svwp":
        a.inducing_variable.Z.assign(np.sort(np.random.default_rng(0).random((Mq, 1)), axis=0))
    a.save_model_params_dict(str(tmp_path), "m.json")
    b = build()
    b.load_from_params_dict(str(tmp_path), "m.json")
    data = (x.numpy(), y.numpy()) if kind == "svwp" else None
    ea = a.elbo(data, epsilon=eps.cuda()).item()
    eb = b.elbo(data, epsilon=eps.cuda()).item()
    # the kernel hyper-parameters pass through softplus / softplus^-1 once (the reference stores constrained values)
    assert abs(ea - eb) <= 1e-12 * abs(ea)
    xn = np.linspace(0, 1, 7)[:, None]
    e2 = torch.randn(11, 7, D, D, dtype=torch.float64, generator=torch.Generator().manual_seed(0)).cuda()
    for u, v in zip(a.predict_cov(xn, 11, epsilon=e2), b.predict_cov(xn, 11, epsilon=e2)):
        assert (u - v).abs().max().item() <= 1e-12 * max(1.0, u.abs().max().item())
    # a second save of the reloaded model is byte-identical to the first once the hyper-parameters have settled
    b.save_model_params_dict(str(tmp_path), "m2.json")
    c = build()
    c.load_from_params_dict(str(tmp_path), "m2.json")
    assert c.elbo(data, epsilon=eps.cuda()).item() == pytest.approx(eb, rel=1e-13)
    ja, jb = json.load(open(tmp_path / "m.json")), json.load(open(tmp_path / "m2.json"))
    for k in ("A_scale_matrix", "additive_noise", "q_mu", "q_sqrt"):
        assert ja[k] == jb[k], k
