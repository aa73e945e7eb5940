"""Generates the committed golden fixtures.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

* sliding_*.npz -- outputs of the REFERENCE's own SlidingWindows class (fcest/models/sliding_windows.py),
  imported from the read-only mount with the unrelated third-party imports stubbed.
* filtering.npz -- outputs of the REFERENCE's own highpass_filter_data (fcest/helpers/filtering.py; numpy + scipy
  only) and of its SlidingWindows.overlapping_windowed_cov_estimation with a repetition time (filter + windows).
* summary.npz -- outputs of the REFERENCE's summarize_tvfc_estimates (fcest/helpers/summary_measures.py) for the
  metrics that need no statsmodels (mean, variance, std, rate_of_change).
* tvfc_reference.csv / tvfc_csv.npz -- a file written by the REFERENCE's SlidingWindows.save_tvfc_estimates and
  what its load_tvfc_estimates reads back from it.
* wishart_reference.npz, svwp_reference_params.json, vwp_reference_params.json -- outputs of the REFERENCE's own
  source text for the FCEst-owned arithmetic of the Wishart path (likelihoods.py, wishart_process.py,
  array_operations.py), executed on the NumPy-backed tensorflow stand-in tests/golden/tf_numpy_shim.py with the
  N(0,1) draws injected.  These pin oracle/wishart_oracle.py (and the CUDA path) to the reference's code for
  everything FCEst itself computes; the GPflow formulas (conditional, KL, Matern52, Adam) stay a restatement.
* wishart_oracle.npz -- outputs of the CPU oracle (oracle/wishart_oracle.py) for the full ELBO + gradient
  (regression fixtures of the restatement, GPflow part included).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference_sliding_windows():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Callable:
        def __init__(self, *a, **k):
            pass

    stub("tensorflow", Tensor=object, Variable=object)
    stub("tensorflow.python"); stub("tensorflow.python.framework")
    stub("tensorflow.python.framework.errors", InvalidArgumentError=Exception)
    stub("gpflow", set_trainable=lambda *a, **k: None)
    stub("gpflow.likelihoods", MonteCarloLikelihood=object)
    stub("gpflow.kernels", Kernel=object, Matern52=_Callable)
    models = stub("gpflow.models")
    models.vgp = stub("gpflow.models.vgp", VGP=object)
    models.svgp = stub("gpflow.models.svgp", SVGP=object)
    stub("gpflow.monitor", ModelToTensorBoard=object, Monitor=object, MonitorTaskGroup=object,
         ScalarToTensorBoard=object)
    sys.modules["gpflow"].models = models
    sys.modules["gpflow"].kernels = sys.modules["gpflow.kernels"]
    stub("rpy2"); stub("rpy2.robjects", pandas2ri=None, r=None)
    stub("rpy2.robjects.packages", importr=lambda *a, **k: None)
    stub("rpy2.robjects.conversion", localconverter=None)
    stub("statsmodels"); stub("statsmodels.tsa"); stub("statsmodels.tsa.ar_model", AutoReg=object)
    sys.path.insert(0, "/root/reference")
    from fcest.models.sliding_windows import SlidingWindows
    return SlidingWindows


def make_sliding():
    SlidingWindows = import_reference_sliding_windows()
    rng = np.random.default_rng(11)
    out = {}
    cases = [("a", 60, 4, 9, 1), ("b", 61, 3, 10, 2), ("c", 48, 7, 15, 3), ("d", 40, 5, 4, 1)]
    for tag, n, d, w, step in cases:
        y = rng.standard_normal((n, d)) + (3.0 if tag == "c" else 0.0)
        x = np.linspace(0, 1, n)[:, None]
        sw = SlidingWindows(x, y)
        out[f"y_{tag}"] = y
        out[f"meta_{tag}"] = np.array([w, step])
        out[f"cov_{tag}"] = sw.overlapping_windowed_cov_estimation(w, step_size=step)
        out[f"corr_{tag}"] = sw.overlapping_windowed_cov_estimation(w, step_size=step,
                                                                     connectivity_metric="correlation")
    np.savez_compressed(os.path.join(HERE, "sliding_windows.npz"), **out)
    # cross-validated window search (reference table as a DataFrame -> arrays)
    cv = {}
    for tag, n, d in [("a", 70, 3), ("b", 64, 25)]:  # b: D=25 makes short windows rank deficient (-inf)
        y = rng.standard_normal((n, d))
        sw = SlidingWindows(np.linspace(0, 1, n)[:, None], y)
        df = sw.find_cross_validated_optimal_window_length()
        cv[f"y_{tag}"] = y
        cv[f"table_{tag}"] = df.to_numpy(dtype=np.float64)
        cv[f"locs_{tag}"] = df.index.to_numpy()
        cv[f"windows_{tag}"] = df.columns.to_numpy()
        cv[f"best_{tag}"] = np.array(sw.get_optimal_window_length(df))
        cv[f"bounds_{tag}"] = np.array([sw.minimum_proposal_window_length, sw.maximum_proposal_window_length])
    sw = SlidingWindows(np.zeros((1200, 1)), np.zeros((1200, 2)))
    cv["bounds_1200"] = np.array([sw.minimum_proposal_window_length, sw.maximum_proposal_window_length])
    np.savez_compressed(os.path.join(HERE, "sliding_cv.npz"), **cv)
    print("sliding fixtures written")


def make_sliding_c3():
    """BASELINE.json config 3 (N=1200, D=100, window_length=30 + CV window search) run through the REFERENCE
    itself (about three minutes: 81 520 eigendecompositions and a DataFrame that grows cell by cell)."""
    SlidingWindows = import_reference_sliding_windows()
    y = np.random.default_rng(0).standard_normal((1200, 100))
    sw = SlidingWindows(np.linspace(0, 1, 1200)[:, None], y)
    cov = sw.overlapping_windowed_cov_estimation(30)
    rows = np.array([0, 15, 600, 1199])
    df = sw.find_cross_validated_optimal_window_length()
    np.savez_compressed(os.path.join(HERE, "sliding_c3.npz"), seed=np.array(0), rows=rows, cov_rows=cov[rows],
                        cov_checksum=np.array([cov.sum(), np.abs(cov).sum(), (cov * cov).sum()]),
                        table=df.to_numpy(dtype=np.float64), locs=df.index.to_numpy(), windows=df.columns.to_numpy(),
                        best=np.array(sw.get_optimal_window_length(df)))
    print("C3 sliding fixtures written; best window", sw.get_optimal_window_length(df))


def make_filtering():
    SlidingWindows = import_reference_sliding_windows()
    from fcest.helpers.filtering import _butter_highpass, highpass_filter_data
    rng = np.random.default_rng(23)
    out = {}
    # (tag, N, D, window length in TRs, repetition time in seconds)
    cases = [("a", 80, 5, 9, 1.0), ("b", 150, 4, 30, 0.72), ("c", 64, 3, 15, 2.0), ("d", 19, 2, 7, 1.0),
             ("e", 400, 6, 121, 1.0)]
    for tag, n, d, w, tr in cases:
        y = rng.standard_normal((n, d)) + (5.0 if tag == "b" else 0.0)
        y += np.sin(np.linspace(0, 3, n))[:, None]  # slow drift for the high-pass to remove
        out[f"y_{tag}"] = y
        out[f"meta_{tag}"] = np.array([w, tr])
        out[f"filtered_{tag}"] = highpass_filter_data(y, window_length=w, repetition_time=tr)
        b, a = _butter_highpass(1 / (w * tr), nyquist_freq=1 / tr / 2)
        out[f"b_{tag}"], out[f"a_{tag}"] = b, a
        if n > 200:
            continue  # keep the fixture small: the long case pins the filter only
        sw = SlidingWindows(np.linspace(0, 1, n)[:, None], y, repetition_time=tr)
        out[f"cov_{tag}"] = sw.overlapping_windowed_cov_estimation(w, repetition_time=tr)
        out[f"corr_{tag}"] = sw.overlapping_windowed_cov_estimation(w, repetition_time=tr,
                                                                     connectivity_metric="correlation")
    np.savez_compressed(os.path.join(HERE, "filtering.npz"), **out)
    print("filtering fixtures written")


def make_summary():
    import_reference_sliding_windows()   # installs the stubs (statsmodels among them) and the import path
    from fcest.helpers.summary_measures import summarize_tvfc_estimates
    rng = np.random.default_rng(31)
    out = {}
    for tag, n, d in [("a", 40, 4), ("b", 25, 7)]:
        a = rng.standard_normal((n, d, d + 2))
        c = a @ np.swapaxes(a, 1, 2) / (d + 2)
        if tag == "b":
            c[3, 1, 2] = c[3, 2, 1] = np.nan      # nan-aware metrics
            c[7, 0, 4] = c[7, 4, 0] = 0.0         # division by zero in the rate of change -> inf -> 1
            c[9:12, 5, 5] = 1e-9                  # step > 10 -> 1
        out[f"c_{tag}"] = c
        with np.errstate(divide="ignore", invalid="ignore"):
            for metric in ("mean", "variance", "std", "rate_of_change"):
                out[f"{metric}_{tag}"] = summarize_tvfc_estimates(c, metric)
    np.savez_compressed(os.path.join(HERE, "summary.npz"), **out)
    print("summary fixtures written")


def make_csv():
    """The reference's CSV round trip: save_tvfc_estimates -> file text, load_tvfc_estimates -> array."""
    import tempfile
    SlidingWindows = import_reference_sliding_windows()
    rng = np.random.default_rng(41)
    y = rng.standard_normal((12, 3))
    sw = SlidingWindows(np.linspace(0, 1, 12)[:, None], y)
    with tempfile.TemporaryDirectory() as d:
        sw.save_tvfc_estimates(4, d, "tvfc.csv", connectivity_metric="covariance")
        text = open(os.path.join(d, "tvfc.csv")).read()
        loaded = SlidingWindows.load_tvfc_estimates(d, "tvfc.csv")
    open(os.path.join(HERE, "tvfc_reference.csv"), "w").write(text)
    np.savez_compressed(os.path.join(HERE, "tvfc_csv.npz"), y=y, loaded=np.asarray(loaded, dtype=np.float64))
    print("csv fixtures written", np.asarray(loaded).shape)


def make_wishart():
    import torch
    from oracle import wishart_oracle as wo
    out = {}
    for tag, kind, N, D, S, M in [("svgp", "svgp", 36, 4, 3, 11), ("vgp", "vgp", 25, 3, 4, None)]:
        x, y, eps, p = wo.make_inputs(N, D, S, M=M, seed=5)
        elbo, g = wo.elbo_and_grad(kind, p, x, y, eps)
        out[f"{tag}_elbo"] = np.array(elbo)
        for k, v in g.items():
            out[f"{tag}_grad_{k}"] = v.numpy()
        xn = torch.linspace(-0.05, 1.05, 9, dtype=torch.float64)[:, None]
        fm, fv = wo.svgp_predict_f(xn, p) if kind == "svgp" else wo.vgp_predict_f(xn, x, p)
        gen = torch.Generator().manual_seed(3)
        e2 = torch.randn(16, 9, D, D, dtype=torch.float64, generator=gen)
        cm, cs = wo.predict_cov(fm, fv, p["A"], p["lam"], e2, D)
        rm, rs = wo.predict_corr(fm, fv, p["A"], p["lam"], e2, D)
        out[f"{tag}_fmean"] = fm.numpy(); out[f"{tag}_fvar"] = fv.numpy()
        out[f"{tag}_cov_mean"] = cm.numpy(); out[f"{tag}_cov_std"] = cs.numpy()
        out[f"{tag}_corr_mean"] = rm.numpy(); out[f"{tag}_corr_std"] = rs.numpy()
        out[f"{tag}_shape"] = np.array([N, D, S, -1 if M is None else M])
    np.savez_compressed(os.path.join(HERE, "wishart_oracle.npz"), **out)
    print("wishart fixtures written")


def make_wishart_reference():
    """Run the reference's own likelihood / prediction / checkpoint source on the NumPy tensorflow stand-in."""
    import json
    import tempfile
    sys.path.insert(0, HERE)
    import tf_numpy_shim as shim
    tf = shim.install()
    for name in [k for k in sys.modules if k == "fcest" or k.startswith("fcest.")]:
        del sys.modules[name]   # a stub-imported copy (sliding fixtures) must not be reused
    from fcest.models.likelihoods import WishartProcessLikelihood
    from fcest.models.wishart_process import SparseVariationalWishartProcess, VariationalWishartProcess
    from fcest.helpers.array_operations import are_all_positive_definite, convert_tensor_to_correlation
    import io
    import contextlib

    rng = np.random.default_rng(77)
    out = {}
    # (tag, N, D, S, A option, perturb the parameters away from their initial values)
    lik_cases = [("full", 12, 5, 3, "train_full_matrix", True), ("ident", 10, 4, 2, "identity", True),
                 ("scaled", 9, 6, 4, "scaled", True), ("diag", 11, 3, 5, "train_diagonal", True),
                 ("init", 8, 4, 3, "train_full_matrix", False), ("d1", 7, 1, 2, "train_full_matrix", True),
                 ("d17", 6, 17, 2, "train_full_matrix", True), ("c1", 200, 3, 5, "train_full_matrix", True),
                 ("d50", 3, 50, 8, "train_full_matrix", True), ("d60", 2, 60, 2, "train_diagonal", True)]
    tags = []
    for tag, N, D, S, option, perturb in lik_cases:
        with contextlib.redirect_stdout(io.StringIO()):
            lik = WishartProcessLikelihood(D=D, num_mc_samples=S, A_scale_matrix_option=option,
                                           train_additive_noise=True, verbose=True)
        out[f"lik_{tag}_A_init"] = lik.A_scale_matrix.numpy().copy()
        out[f"lik_{tag}_lam_init"] = lik.additive_part.numpy().copy()
        out[f"lik_{tag}_dims"] = np.array([lik.input_dim, lik.latent_dim, lik.observation_dim, lik.D, lik.nu,
                                           int(lik.A_scale_matrix.trainable), int(lik.additive_part.trainable)])
        if perturb:
            if option in ("train_full_matrix", "train_diagonal"):
                lik.A_scale_matrix.assign(lik.A_scale_matrix.numpy() + 0.1 * rng.standard_normal(lik.A_scale_matrix.shape))
            lik.additive_part.assign(lik.additive_part.numpy() + 0.3 * rng.standard_normal(D))
        R = D * D
        x = np.linspace(0, 1, N)[:, None]
        y = rng.standard_normal((N, D))
        f_mean = 0.5 * rng.standard_normal((N, R))
        f_var = 0.05 + rng.random((N, R))
        eps = rng.standard_normal((S, N, R))
        shim.push_normal(eps)
        ve = lik.variational_expectations(x, tf.constant(f_mean), tf.constant(f_var), y)
        f_sample = rng.standard_normal((S, N, D, D))
        lp = lik._log_prob(x, tf.constant(f_sample), y)
        cov = rng.standard_normal((2, D, D))
        out[f"lik_{tag}_A"] = lik.A_scale_matrix.numpy().copy()
        out[f"lik_{tag}_lam"] = lik.additive_part.numpy().copy()
        out[f"lik_{tag}_y"] = y
        out[f"lik_{tag}_f_mean"] = f_mean
        out[f"lik_{tag}_f_var"] = f_var
        out[f"lik_{tag}_eps"] = eps
        out[f"lik_{tag}_ve"] = ve.numpy()
        out[f"lik_{tag}_f_sample"] = f_sample
        out[f"lik_{tag}_log_prob"] = lp.numpy()
        out[f"lik_{tag}_cov_in"] = cov
        out[f"lik_{tag}_cov_noise"] = lik._add_diagonal_additive_noise(tf.constant(cov)).numpy()
        tags.append(f"{tag}:{option}")
    out["lik_cases"] = np.array(tags)
    try:
        WishartProcessLikelihood(D=4, nu=3)
        raise AssertionError("nu < D must raise")
    except Exception as e:
        out["nu_lt_D_message"] = np.array(str(e))

    # prediction: the reference's predict_cov / predict_corr / _get_cov_samples run on a model object whose
    # predict_f (GPflow's) returns a prescribed (mean, variance) pair
    ptags = []
    for tag, cls, N, D, S, option in [("vwp", VariationalWishartProcess, 6, 3, 7, "train_full_matrix"),
                                      ("svwp", SparseVariationalWishartProcess, 5, 4, 9, "train_full_matrix"),
                                      ("svwp_diag", SparseVariationalWishartProcess, 4, 9, 6, "train_diagonal"),
                                      ("vwp_scaled", VariationalWishartProcess, 3, 18, 5, "scaled")]:
        model = object.__new__(cls)
        model.D, model.nu = D, D
        with contextlib.redirect_stdout(io.StringIO()):
            model.likelihood = WishartProcessLikelihood(D=D, num_mc_samples=2, A_scale_matrix_option=option,
                                                        train_additive_noise=True, verbose=False)
        lik = model.likelihood
        if option in ("train_full_matrix", "train_diagonal"):
            lik.A_scale_matrix.assign(lik.A_scale_matrix.numpy() + 0.1 * rng.standard_normal(lik.A_scale_matrix.shape))
        lik.additive_part.assign(lik.additive_part.numpy() + 0.3 * rng.standard_normal(D))
        f_mean = 0.7 * rng.standard_normal((N, D * D))
        f_var = 0.05 + rng.random((N, D * D))
        model.predict_f = lambda x_new, fm=f_mean, fv=f_var: (tf.constant(fm), tf.constant(fv))
        x_new = np.linspace(0, 1, N)[:, None]
        eps = rng.standard_normal((S, N, D, D))
        shim.push_normal(eps)
        samples = model._get_cov_samples(x_new, S)
        shim.push_normal(eps)
        cov_mean, cov_std = model.predict_cov(x_new, S)
        shim.push_normal(eps)
        corr_mean, corr_std = model.predict_corr(x_new, S)
        if cls is SparseVariationalWishartProcess:
            shim.push_normal(eps)
            assert np.array_equal(model.predict_cov_samples(x_new, S).numpy(), samples.numpy())
        out[f"pred_{tag}_A"] = lik.A_scale_matrix.numpy().copy()
        out[f"pred_{tag}_lam"] = lik.additive_part.numpy().copy()
        out[f"pred_{tag}_f_mean"] = f_mean
        out[f"pred_{tag}_f_var"] = f_var
        out[f"pred_{tag}_eps"] = eps
        out[f"pred_{tag}_samples"] = samples.numpy()
        out[f"pred_{tag}_cov_mean"] = cov_mean.numpy()
        out[f"pred_{tag}_cov_std"] = cov_std.numpy()
        out[f"pred_{tag}_corr_mean"] = corr_mean.numpy()
        out[f"pred_{tag}_corr_std"] = corr_std.numpy()
        out[f"pred_{tag}_corr_samples"] = convert_tensor_to_correlation(samples).numpy()
        ptags.append(f"{tag}:{option}")
    out["pred_cases"] = np.array(ptags)

    # the PD check on tensors (array_operations.py:75-103)
    a = rng.standard_normal((4, 5, 7))
    pd = a @ np.swapaxes(a, 1, 2) + 0.1 * np.eye(5)
    not_pd = pd.copy(); not_pd[2] -= 3.0 * np.linalg.eigvalsh(pd[2]).max() * np.eye(5)
    asym = pd.copy(); asym[1, 0, 3] += 1e-3
    with contextlib.redirect_stdout(io.StringIO()):
        out["pd_inputs"] = np.stack([pd, not_pd, asym])
        out["pd_truth"] = np.array([are_all_positive_definite(tf.constant(m)) for m in (pd, not_pd, asym)])

    # JSON checkpoints written by the reference's save_model_params_dict, and what its load_from_params_dict
    # assigns from them (wishart_process.py:239-308, 526-602)
    class _Obj:
        pass

    def fake(cls, D, M, option, sparse):
        m = object.__new__(cls)
        m.D, m.nu = D, D
        with contextlib.redirect_stdout(io.StringIO()):
            m.likelihood = WishartProcessLikelihood(D=D, A_scale_matrix_option=option, train_additive_noise=True,
                                                    verbose=False)
        m.kernel = _Obj()
        m.kernel.variance = tf.Variable(1.0)
        m.kernel.lengthscales = tf.Variable(0.3)
        if sparse:
            m.inducing_variable = _Obj()
            m.inducing_variable.Z = tf.Variable(np.linspace(0, 1, M)[:, None])
        m.q_mu = tf.Variable(np.zeros((M, D * D)))
        m.q_sqrt = tf.Variable(np.tile(1e-3 * np.eye(M)[None], (D * D, 1, 1)))
        return m

    def randomise(m, sparse):
        lik = m.likelihood
        lik.A_scale_matrix.assign(lik.A_scale_matrix.numpy() + 0.1 * rng.standard_normal(lik.A_scale_matrix.shape))
        lik.additive_part.assign(lik.additive_part.numpy() + 0.3 * rng.standard_normal(lik.additive_part.shape))
        m.kernel.variance.assign(1.0 + rng.random())
        m.kernel.lengthscales.assign(0.2 + rng.random())
        if sparse:
            m.inducing_variable.Z.assign(np.sort(rng.random(m.inducing_variable.Z.shape), axis=0))
        m.q_mu.assign(rng.standard_normal(m.q_mu.shape))
        m.q_sqrt.assign(np.tril(rng.standard_normal(m.q_sqrt.shape)))

    logging_off = __import__("logging").disable
    logging_off(50)
    for fname, cls, D, M, option, sparse in [("svwp_reference_params.json", SparseVariationalWishartProcess, 2, 3, "train_full_matrix", True),
                                             ("svwp_diag_reference_params.json", SparseVariationalWishartProcess, 3, 2, "train_diagonal", True),
                                             ("vwp_reference_params.json", VariationalWishartProcess, 2, 4, "train_full_matrix", False)]:
        m = fake(cls, D, M, option, sparse)
        randomise(m, sparse)
        with tempfile.TemporaryDirectory() as d:
            m.save_model_params_dict(d, fname)
            text = open(os.path.join(d, fname)).read()
            # a fresh model of the class the reference can load this file into (a 1-D A becomes diag(A) for SVWP)
            m2 = fake(cls, D, M, "train_full_matrix" if sparse else option, sparse)
            with contextlib.redirect_stdout(io.StringIO()):
                m2.load_from_params_dict(d, fname)
        open(os.path.join(HERE, fname), "w").write(text)
        key = fname.replace("_reference_params.json", "")
        out[f"json_{key}_A"] = m2.likelihood.A_scale_matrix.numpy().copy()
        out[f"json_{key}_lam"] = m2.likelihood.additive_part.numpy().copy()
        out[f"json_{key}_kvar"] = np.array(float(m2.kernel.variance.numpy()))
        out[f"json_{key}_kls"] = np.array(float(m2.kernel.lengthscales.numpy()))
        out[f"json_{key}_q_mu"] = m2.q_mu.numpy().copy()
        out[f"json_{key}_q_sqrt"] = m2.q_sqrt.numpy().copy()
        if sparse:
            out[f"json_{key}_Z"] = m2.inducing_variable.Z.numpy().copy()
        assert list(json.loads(text).keys()) == (["D", "nu", "A_scale_matrix", "additive_noise", "kernel_variance",
                                                  "kernel_lengthscales"] + (["Z"] if sparse else []) + ["q_mu", "q_sqrt"])
    logging_off(0)
    np.savez_compressed(os.path.join(HERE, "wishart_reference.npz"), **out)
    print("reference-source wishart fixtures written:", os.path.getsize(os.path.join(HERE, "wishart_reference.npz")), "bytes")


if __name__ == "__main__":
    which = sys.argv[1:] or ["wishart", "sliding", "filtering", "summary", "csv", "wishart_reference"]
    if "wishart_reference" in which:
        import subprocess
        if len(which) > 1:   # own process: the tensorflow stand-in must not leak into the other generators
            subprocess.run([sys.executable, os.path.abspath(__file__), "wishart_reference"], check=True)
            which = [w for w in which if w != "wishart_reference"]
        else:
            make_wishart_reference()
            sys.exit(0)
    if "summary" in which:
        make_summary()
    if "csv" in which:
        make_csv()
    if "wishart" in which:
        make_wishart()
    if "sliding" in which:
        make_sliding()
    if "sliding_c3" in which:   # not in the default list: takes minutes
        make_sliding_c3()
    if "filtering" in which:
        make_filtering()
