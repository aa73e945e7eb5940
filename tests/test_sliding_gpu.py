"""Sliding-window kernels against the reference's own outputs (golden fixtures) and the CPU oracle.
Tolerance from BASELINE.json's north_star: 1e-12 (relative to the largest entry) for covariance /
correlation estimates.  The CV table is held to 1e-6 relative (Cholesky here vs eigh in SciPy; windows with
w - 2 close to D are ill conditioned) with an exact match of the -inf pattern."""
import os

import numpy as np
import pytest

from oracle import sliding_oracle as so

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_window_cov_matches_reference_fixtures(cuda):
    from fcest_b200.models import SlidingWindows
    g = np.load(os.path.join(GOLD, "sliding_windows.npz"))
    for tag in "abcd":
        y = g[f"y_{tag}"]
        w, step = (int(v) for v in g[f"meta_{tag}"])
        sw = SlidingWindows(np.linspace(0, 1, len(y))[:, None], y)
        cov = sw.overlapping_windowed_cov_estimation(w, step_size=step)
        corr = sw.overlapping_windowed_cov_estimation(w, step_size=step, connectivity_metric="correlation")
        scale = max(1.0, np.abs(g[f"cov_{tag}"]).max())
        assert np.abs(cov - g[f"cov_{tag}"]).max() <= 1e-12 * scale
        assert np.abs(corr - g[f"corr_{tag}"]).max() <= 1e-12


# bulk-store kernel: (1200,100,30) two staging buffers, 8-9 windows per CTA; (128,64,31) one window per CTA;
# (2500,100,30) 16 windows per CTA over two waves; (700,112,25) widest D with one buffer; (1200,64,121) long
# windows; (60,4,9) tiny tiles.  Register-store kernel: step 2, odd D, (1200,100,181) rows exceed shared memory
# next to a staging buffer, (400,130,40) D > 112.
@pytest.mark.parametrize("N,D,w,step", [(1200, 100, 30, 1), (300, 17, 45, 2), (128, 64, 31, 1), (90, 3, 2, 1),
                                        (2500, 100, 30, 1), (700, 112, 25, 1), (1200, 64, 121, 1), (60, 4, 9, 1),
                                        (1200, 100, 181, 1), (400, 130, 40, 1), (1200, 100, 30, 3)])
def test_window_cov_matches_oracle(cuda, N, D, w, step):
    from fcest_b200.models import SlidingWindows
    rng = np.random.default_rng(N + D)
    y = rng.standard_normal((N, D)) * 2.0 + rng.standard_normal(D) * 5.0  # non-centred on purpose
    sw = SlidingWindows(np.linspace(0, 1, N)[:, None], y)
    for metric in ("covariance", "correlation"):
        got = sw.overlapping_windowed_cov_estimation(w, step_size=step, connectivity_metric=metric)
        ref = so.overlapping_windowed_cov(y, w, step, metric)
        assert got.shape == (N, D, D)
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), metric
        assert np.array_equal(got, np.swapaxes(got, 1, 2)) or np.abs(got - np.swapaxes(got, 1, 2)).max() < 1e-13


def test_window_cov_badly_scaled_data(cuda):
    """Scale guard (direct accumulation for blocks whose windows differ in scale) in the bulk-store kernel."""
    from fcest_b200.models import SlidingWindows
    rng = np.random.default_rng(5)
    y = rng.standard_normal((600, 20)) * np.exp(np.linspace(0, 18, 600))[:, None] + 100.0
    sw = SlidingWindows(np.zeros((600, 1)), y)
    got = sw.overlapping_windowed_cov_estimation(12)
    ref = so.overlapping_windowed_cov(y, 12, 1, "covariance")
    scale = np.abs(ref).reshape(600, -1).max(axis=1)[:, None, None]
    assert (np.abs(got - ref) <= 1e-12 * np.maximum(scale, 1e-300)).all()


def test_window_cov_zero_columns(cuda):
    """corr[cov == 0] = 0 (array_operations.py:52): a constant column gives zeros, not NaN."""
    from fcest_b200.models import SlidingWindows
    rng = np.random.default_rng(0)
    y = rng.standard_normal((40, 4))
    y[:, 2] = 0.0
    sw = SlidingWindows(np.zeros((40, 1)), y)
    got = sw.overlapping_windowed_cov_estimation(8, connectivity_metric="correlation")
    ref = so.overlapping_windowed_cov(y, 8, 1, "correlation")
    assert np.all(np.isfinite(got)) and np.abs(got - ref).max() <= 1e-12


def test_cv_matches_reference_fixtures(cuda):
    from fcest_b200.models import SlidingWindows
    g = np.load(os.path.join(GOLD, "sliding_cv.npz"))
    for tag in "ab":
        y = g[f"y_{tag}"]
        sw = SlidingWindows(np.linspace(0, 1, len(y))[:, None], y)
        assert (sw.minimum_proposal_window_length, sw.maximum_proposal_window_length) == tuple(g[f"bounds_{tag}"])
        df = sw.find_cross_validated_optimal_window_length()
        assert np.array_equal(df.index.to_numpy(), g[f"locs_{tag}"])
        assert np.array_equal(df.columns.to_numpy(), g[f"windows_{tag}"])
        got, ref = df.to_numpy(), g[f"table_{tag}"]
        assert np.array_equal(np.isinf(got), np.isinf(ref))
        fin = np.isfinite(ref)
        assert np.all(np.abs(got[fin] - ref[fin]) <= 1e-6 * np.abs(ref[fin]))
        assert sw.compute_cross_validated_optimal_window_length() == int(g[f"best_{tag}"])


def test_cv_default_bounds_n1200(cuda):
    from fcest_b200.models import SlidingWindows
    sw = SlidingWindows(np.zeros((1200, 1)), np.zeros((1200, 2)))
    assert (sw.minimum_proposal_window_length, sw.maximum_proposal_window_length) == (21, 181)


def test_cv_larger_problem_matches_oracle(cuda):
    from fcest_b200.models import SlidingWindows
    rng = np.random.default_rng(5)
    y = rng.standard_normal((260, 12))
    sw = SlidingWindows(np.linspace(0, 1, 260)[:, None], y)
    df = sw.find_cross_validated_optimal_window_length(eval_location_step_size=7)
    windows, locs = so.cv_grid(260, 2, 7)
    _, _, ref = so.cross_validated_table(y, windows=windows, locs=locs)
    got = df.to_numpy()
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.all(np.abs(got[fin] - ref[fin]) <= 1e-8 * np.abs(ref[fin]))


def test_save_and_load_tvfc_estimates_match_reference_file(cuda, tmp_path):
    """save_tvfc_estimates writes the reference's CSV layout: parsed, it equals the file the reference wrote for
    the same data (golden fixture) to 1e-12, and load_tvfc_estimates reads our file back."""
    import pandas as pd
    from fcest_b200.models import SlidingWindows
    g = np.load(os.path.join(GOLD, "tvfc_csv.npz"))
    sw = SlidingWindows(np.linspace(0, 1, 12)[:, None], g["y"])
    sw.save_tvfc_estimates(4, str(tmp_path / "out"), "tvfc.csv", connectivity_metric="covariance")
    ours = pd.read_csv(tmp_path / "out" / "tvfc.csv")
    ref = pd.read_csv(os.path.join(GOLD, "tvfc_reference.csv"))
    assert list(ours.columns) == list(ref.columns) and ours.shape == ref.shape
    assert np.abs(ours.to_numpy() - ref.to_numpy()).max() <= 1e-12
    back = SlidingWindows.load_tvfc_estimates(str(tmp_path / "out"), "tvfc.csv")
    assert np.abs(np.asarray(back, dtype=np.float64) - g["loaded"]).max() <= 1e-12
