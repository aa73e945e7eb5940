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
// 1000-subject batch, which runs at HBM speed as is (3 fp64 operations per step).
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

// The extended signal has three segments -- edge reflected samples, the N samples themselves, edge reflected samples --
// and each pass walks them in separate loops, so the long middle loop has no branches: batches of FF_AHEAD
// loads, then FF_AHEAD recurrence steps (and stores).
template <int ORD>
__global__ void __launch_bounds__(32) filtfilt_kernel(const __grid_constant__ FiltParams p) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const int N = p.N, edge = p.edge;
  const long long C = p.C;
  const double* x = p.x + c;
  double* tmp = p.tmp + c;
  double* out = p.out + c;
  const double x0 = x[0], xl = x[(long long)(N - 1) * C];
  const double two_x0 = __dmul_rn(2.0, x0), two_xl = __dmul_rn(2.0, xl);
  double z[FF_MAXORD], buf[FF_AHEAD];
  // ---- forward pass: tmp[k] = lfilter(ext)[k], zi scaled by ext[0] = 2 x_0 - x_edge
  {
    const double e0 = __dsub_rn(two_x0, x[(long long)edge * C]);
#pragma unroll
    for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(p.zi[i], e0);
  }
  for (int k = 0; k < edge; ++k)
    tmp[(long long)k * C] = ff_step<ORD>(p, z, __dsub_rn(two_x0, x[(long long)(edge - k) * C]));
  int j = 0;
  for (; j + FF_AHEAD <= N; j += FF_AHEAD) {
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u) buf[u] = x[(long long)(j + u) * C];
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u) tmp[(long long)(edge + j + u) * C] = ff_step<ORD>(p, z, buf[u]);
  }
  for (; j < N; ++j) tmp[(long long)(edge + j) * C] = ff_step<ORD>(p, z, x[(long long)j * C]);
  double ylast = 0.0;
  for (int t = 0; t < edge; ++t) {   // ext[N + edge + t] = 2 x_{N-1} - x_{N-2-t}
    ylast = ff_step<ORD>(p, z, __dsub_rn(two_xl, x[(long long)(N - 2 - t) * C]));
    tmp[(long long)(edge + N + t) * C] = ylast;
  }
  // ---- backward pass over the reversed forward output (this thread wrote every element it reads), zi scaled by
  //      its last sample; only the middle segment is kept
#pragma unroll
  for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(p.zi[i], ylast);
  for (int t = edge - 1; t >= 0; --t) (void)ff_step<ORD>(p, z, tmp[(long long)(edge + N + t) * C]);
  j = N;
  for (; j - FF_AHEAD >= 0; j -= FF_AHEAD) {
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u) buf[u] = tmp[(long long)(edge + j - 1 - u) * C];
#pragma unroll
    for (int u = 0; u < FF_AHEAD; ++u) out[(long long)(j - 1 - u) * C] = ff_step<ORD>(p, z, buf[u]);
  }
  for (; j > 0; --j) out[(long long)(j - 1) * C] = ff_step<ORD>(p, z, tmp[(long long)(edge + j - 1) * C]);
  // the leading edge of the reversed pass only feeds samples that are cut off: nothing left to do
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
