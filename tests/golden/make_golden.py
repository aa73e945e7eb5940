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
* wishart_*.npz -- outputs of the CPU oracle (oracle/wishart_oracle.py).  TensorFlow / GPflow cannot be
  installed here, so these pin the *restatement* (regression fixtures), not the reference: "parity
  unpinned" for the Wishart path, as stated in the oracle header and DESIGN.md.
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


if __name__ == "__main__":
    which = sys.argv[1:] or ["wishart", "sliding", "filtering", "summary", "csv"]
    if "summary" in which:
        make_summary()
    if "csv" in which:
        make_csv()
    if "wishart" in which:
        make_wishart()
    if "sliding" in which:
        make_sliding()
    if "filtering" in which:
        make_filtering()
