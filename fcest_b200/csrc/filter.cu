// Zero-phase IIR filtering of every column of a (N, C) array: the high-pass step the reference applies in
// front of the sliding-window estimator when a repetition time is given.
//
// Reference: fcest/helpers/filtering.py:12-67 (highpass_filter_data -> scipy.signal.filtfilt(b, a, column) with
// scipy's defaults: padtype 'odd', padlen = 3 * max(len(a), len(b)), method 'pad'), called from
// fcest/models/sliding_windows.py:147-154.  scipy's algorithm, restated:
//   ext      = [2 x_0 - x_edge .. 2 x_0 - x_1,  x_0 .. x_{N-1},  2 x_{N-1} - x_{N-2} .. 2 x_{N-1} - x_{N-1-edge}]
//   forward  = lfilter(b, a, ext,           zi * ext[0])        (direct form II transposed)
//   backward = lfilter(b, a, forward[::-1], zi * forward[-1])[::-1],  result = backward[edge : -edge]
//   lfilter step:  y = z_0 + b_0 x;   z_i = (z_{i+1} + x b_{i+1}) - y a_{i+1};   z_last = x b_n - y a_n
//
// The recurrence is sequential in time and independent per column, so one thread owns one column (consecutive
// threads = consecutive columns: every load / store of a time step is one coalesced row segment) and the
// samples of the next FF_AHEAD steps are fetched in one batch ahead of the dependent chain; a software-pipelined
// variant (next batch in flight during the current one) gained 10 % on one subject and lost 24 % on a
// 1000-subject batch, which runs at 4.9 TB/s as is (3 fp64 operations per step).
// Every product and sum is rounded separately (__dmul_rn / __dadd_rn, never contracted into FMA) in scipy's
// order of operations, which makes the result bit-identical to scipy's C loop.
#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

constexpr int FF_MAXORD = 8;
constexpr int FF_AHEAD = 8;

struct FiltParams {
  const double* x;   // (N, C)
  double* tmp;       // (N + 2 edge, C): forward pass
  double* out;       // (N, C)
  int N, edge, order;
  long long C;
  double b[FF_MAXORD + 1], a[FF_MAXORD + 1], zi[FF_MAXORD];
};

template <int ORD>
__device__ __forceinline__ double ff_step(const FiltParams& p, double (&z)[FF_MAXORD], double xn) {
  const double yn = __dadd_rn(z[0], __dmul_rn(p.b[0], xn));
#pragma unroll
  for (int i = 0; i < ORD - 1; ++i)
    z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(xn, p.b[i + 1])), __dmul_rn(yn, p.a[i + 1]));
  z[ORD - 1] = __dsub_rn(__dmul_rn(xn, p.b[ORD]), __dmul_rn(yn, p.a[ORD]));
  return yn;
}

// element k of the odd extension of column c (k in [0, N + 2 edge))
__device__ __forceinline__ double ff_ext(const FiltParams& p, long long c, int k, double x0, double xl) {
  const int j = k - p.edge;
  if (j < 0) return __dsub_rn(__dmul_rn(2.0, x0), p.x[(long long)(-j) * p.C + c]);
  if (j >= p.N) return __dsub_rn(__dmul_rn(2.0, xl), p.x[(long long)(2 * (p.N - 1) - j) * p.C + c]);
  return p.x[(long long)j * p.C + c];
}

template <int ORD>
__global__ void __launch_bounds__(32) filtfilt_kernel(const __grid_constant__ FiltParams p) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const int L = p.N + 2 * p.edge;
  const double x0 = p.x[c], xl = p.x[(long long)(p.N - 1) * p.C + c];
  double z[FF_MAXORD], buf[FF_AHEAD];
  // forward pass over the extended signal
  const double e0 = ff_ext(p, c, 0, x0, xl);
#pragma unroll
  for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(p.zi[i], e0);
  double ylast = 0.0;
  for (int k0 = 0; k0 < L; k0 += FF_AHEAD) {
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u) buf[u] = (k0 + u < L) ? ff_ext(p, c, k0 + u, x0, xl) : 0.0;
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u)
      if (k0 + u < L) {
        ylast = ff_step<ORD>(p, z, buf[u]);
        p.tmp[(long long)(k0 + u) * p.C + c] = ylast;
      }
  }
  // backward pass over the reversed forward output (same thread wrote every element it reads)
#pragma unroll
  for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(p.zi[i], ylast);
  for (int k0 = L - 1; k0 >= 0; k0 -= FF_AHEAD) {
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u) buf[u] = (k0 - u >= 0) ? p.tmp[(long long)(k0 - u) * p.C + c] : 0.0;
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u)
      if (k0 - u >= 0) {
        const double yn = ff_step<ORD>(p, z, buf[u]);
        const int j = k0 - u - p.edge;
        if (j >= 0 && j < p.N) p.out[(long long)j * p.C + c] = yn;
      }
  }
}

template <int ORD>
static int launch_filtfilt(const FiltParams& p, cudaStream_t stream) {
  const long long blocks = (p.C + 31) / 32;   // one warp per CTA: spreads the columns over as many SMs as possible
  filtfilt_kernel<ORD><<<(unsigned)blocks, 32, 0, stream>>>(p);
  FCEST_CHECK_LAUNCH();
  return 0;
}

}  // namespace fcest

extern "C" long long fcest_filtfilt_scratch_bytes(int N, long long C, int order) {
  if (N < 1 || C < 1 || order < 1) return 0;
  return (long long)(N + 6 * (order + 1)) * C * (long long)sizeof(double);
}

extern "C" int fcest_filtfilt(const double* x, int N, long long C, const double* b, const double* a,
                              const double* zi, int order, double* scratch, double* out, cudaStream_t stream) {
  using namespace fcest;
  if (order < 1 || order > FF_MAXORD || C < 0 || !b || !a || !zi) return -3;
  const int edge = 3 * (order + 1);
  if (N <= edge) return -2;   // scipy: "The length of the input vector x must be greater than padlen"
  if (C == 0) return 0;
  if (!x || !scratch || !out) return -3;
  if (a[0] != 1.0) return -3;  // coefficients must be normalised (scipy's butter always returns a[0] = 1)
  FiltParams p;
  p.x = x; p.tmp = scratch; p.out = out; p.N = N; p.edge = edge; p.order = order; p.C = C;
  for (int i = 0; i <= FF_MAXORD; ++i) { p.b[i] = i <= order ? b[i] : 0.0; p.a[i] = i <= order ? a[i] : 0.0; }
  for (int i = 0; i < FF_MAXORD; ++i) p.zi[i] = i < order ? zi[i] : 0.0;
  switch (order) {
    case 1: return launch_filtfilt<1>(p, stream);
    case 2: return launch_filtfilt<2>(p, stream);
    case 3: return launch_filtfilt<3>(p, stream);
    case 4: return launch_filtfilt<4>(p, stream);
    case 5: return launch_filtfilt<5>(p, stream);
    case 6: return launch_filtfilt<6>(p, stream);
    case 7: return launch_filtfilt<7>(p, stream);
    default: return launch_filtfilt<8>(p, stream);
  }
}
