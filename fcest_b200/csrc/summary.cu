// Temporal summaries of a TVFC estimate: (N, D, D) -> (D, D).
//
// Reference: fcest/helpers/summary_measures.py:16-126 (summarize_tvfc_estimates: np.nanmean / nanvar / nanstd over
// the time axis, compute_rate_of_change = mean_t |c_{t+1} / c_t - 1| with inf and > 10 mapped to 1, and the
// AR(1) coefficient of every edge series, statsmodels AutoReg(lags=1) = least squares of c_{t+1} on [1, c_t]).
// Consumers of the sliding-window / Wishart-process estimates.
//
// One thread per matrix entry walks the time axis; consecutive threads read consecutive entries of a time
// step (coalesced), SM_UNROLL steps are in flight per thread.  Sums run in time order, which is numpy's order
// for a reduction over the leading axis, so mean / variance / std / rate of change are bit-identical to numpy.
#include <math.h>

#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

constexpr int SM_UNROLL = 32;  // loads in flight per thread: D^2 threads alone (10^4 at D = 100) do not cover the HBM latency
constexpr int SM_BLOCK = 64;   // small CTAs so that D^2 / 64 of them reach every SM

enum { SUM_MEAN = 0, SUM_VAR = 1, SUM_STD = 2, SUM_ROC = 3, SUM_AR1 = 4 };

template <int METRIC>
__global__ void __launch_bounds__(SM_BLOCK) tvfc_summary_kernel(const double* __restrict__ c, int N, int D,
                                                           double* __restrict__ out) {
  const long long E = (long long)D * D;
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long eo = e;
  if (METRIC == SUM_AR1) {   // lower-triangle series for both (i, j) and (j, i); diagonal stays 0
    const int i = (int)(e / D), j = (int)(e - (long long)i * D);
    if (i == j) { out[eo] = 0.0; return; }
    if (j > i) e = (long long)j * D + i;
  }
  const double* p = c + e;
  double v[SM_UNROLL];
  if (METRIC == SUM_MEAN || METRIC == SUM_VAR || METRIC == SUM_STD) {
    double s = 0.0;
    int cnt = 0;
    for (int t0 = 0; t0 < N; t0 += SM_UNROLL) {
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u) v[u] = (t0 + u < N) ? p[(long long)(t0 + u) * E] : nan("");
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u)
        if (t0 + u < N) { const bool ok = !isnan(v[u]); s = __dadd_rn(s, ok ? v[u] : 0.0); cnt += ok; }
    }
    const double mean = s / (double)cnt;   // 0 / 0 = NaN for an all-NaN entry, like numpy
    if (METRIC == SUM_MEAN) { out[eo] = mean; return; }
    double q = 0.0;
    for (int t0 = 0; t0 < N; t0 += SM_UNROLL) {
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u) v[u] = (t0 + u < N) ? p[(long long)(t0 + u) * E] : nan("");
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u)
        if (t0 + u < N) {
          const double d = isnan(v[u]) ? 0.0 : __dsub_rn(v[u], mean);
          q = __dadd_rn(q, __dmul_rn(d, d));
        }
    }
    const double var = q / (double)cnt;
    out[eo] = (METRIC == SUM_VAR) ? var : sqrt(var);
  } else if (METRIC == SUM_ROC) {
    double s = 0.0, prev = p[0];
    for (int t0 = 1; t0 < N; t0 += SM_UNROLL) {
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u) v[u] = (t0 + u < N) ? p[(long long)(t0 + u) * E] : 0.0;
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u)
        if (t0 + u < N) {
          double r = fabs(__dsub_rn(v[u] / prev, 1.0));
          if (isinf(r)) r = 1.0;
          if (r > 10.0) r = 1.0;
          s = __dadd_rn(s, r);
          prev = v[u];
        }
    }
    out[eo] = s / (double)(N - 1);
  } else {  // AR(1): slope of the least-squares line of c_{t+1} on c_t (with intercept), centred two-pass
    double sx = 0.0;
    for (int t = 0; t < N; ++t) sx += p[(long long)t * E];
    const double first = p[0], last = p[(long long)(N - 1) * E];
    const double mx = (sx - last) / (double)(N - 1), mz = (sx - first) / (double)(N - 1);
    double sxz = 0.0, sxx = 0.0, prev = first;
    for (int t0 = 1; t0 < N; t0 += SM_UNROLL) {
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u) v[u] = (t0 + u < N) ? p[(long long)(t0 + u) * E] : 0.0;
#pragma unroll
      for (int u = 0; u < SM_UNROLL; ++u)
        if (t0 + u < N) {
          const double dx = prev - mx;
          sxz += dx * (v[u] - mz);
          sxx += dx * dx;
          prev = v[u];
        }
    }
    out[eo] = sxz / sxx;
  }
}

}  // namespace fcest

extern "C" int fcest_tvfc_summary(const double* cov, int N, int D, int metric, double* out, cudaStream_t stream) {
  using namespace fcest;
  if (D <= 0) return 0;
  if (N < 1 || !cov || !out) return FCEST_ERR_ARGUMENT;
  if ((metric == SUM_ROC || metric == SUM_AR1) && N < 2) return FCEST_ERR_ARGUMENT;
  const long long E = (long long)D * D;
  const unsigned grid = (unsigned)((E + SM_BLOCK - 1) / SM_BLOCK);
  switch (metric) {
    case SUM_MEAN: tvfc_summary_kernel<SUM_MEAN><<<grid, SM_BLOCK, 0, stream>>>(cov, N, D, out); break;
    case SUM_VAR: tvfc_summary_kernel<SUM_VAR><<<grid, SM_BLOCK, 0, stream>>>(cov, N, D, out); break;
    case SUM_STD: tvfc_summary_kernel<SUM_STD><<<grid, SM_BLOCK, 0, stream>>>(cov, N, D, out); break;
    case SUM_ROC: tvfc_summary_kernel<SUM_ROC><<<grid, SM_BLOCK, 0, stream>>>(cov, N, D, out); break;
    case SUM_AR1: tvfc_summary_kernel<SUM_AR1><<<grid, SM_BLOCK, 0, stream>>>(cov, N, D, out); break;
    default: return FCEST_ERR_ARGUMENT;
  }
  FCEST_CHECK_LAUNCH();
  return 0;
}
