// Sliding-window covariance estimator and its cross-validated window-length search.
//
// Reference: fcest/models/sliding_windows.py:143-174 (overlapping_windowed_cov_estimation: zero padding
// of floor(w/2) rows at both ends, np.cov per window = two-pass centred X^T X * 1/(w-1), optional
// cov -> corr with corr[cov == 0] = 0, fcest/helpers/array_operations.py:37-53) and :219-272
// (leave-centre-out window covariance + zero-mean MVN log density of the held-out row).
//
// window_cov_kernel: one CTA per output window; the window is staged (centred) in shared memory and
// the D x D Gram matrix is produced on the fp64 tensor pipe, each CTA writing its 8 D^2 bytes in
// 64-byte row segments.  Output traffic (8 N D^2 bytes, 96 MB at N=1200, D=100) bounds the kernel.
// Each window is computed directly (two-pass), NOT by differencing prefix sums: the reference's
// 1e-12 tolerance does not survive a global prefix-sum on non-centred data.
#include <math.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

__global__ void __launch_bounds__(256)
window_cov_kernel(const double* __restrict__ y, int N, int D, int w, int step, int nwin, int corr,
                  double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const int Dp = (D + 7) / 8 * 8, Pn = pitch4(Dp), wp = (w + 3) / 4 * 4;
  double* Xs = sm;              // wp x Pn, centred window (zero padded)
  double* mean = Xs + wp * Pn;  // Dp
  double* vd = mean + Dp;       // Dp: sqrt(diag(cov))
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int pad = w / 2;
  const double fact = 1.0 / (double)(w - 1);
  const int nb = Dp / 8, ntiles = nb * (nb + 1) / 2;
  const bool vec = ((D & 1) == 0);
  // persistent over output windows: the grid is sized to one resident wave (see fcest_window_cov)
  for (int i = blockIdx.x; i < N; i += gridDim.x) {
    double* o = out + (long long)i * D * D;
    if (i >= nwin) {  // rows >= int(N / step) stay zero (sliding_windows.py:164-165)
      for (int e = tid; e < D * D; e += blockDim.x) o[e] = 0.0;
      continue;
    }
    const int start = i * step;
    __syncthreads();
    for (int e = tid; e < wp * Pn; e += blockDim.x) {
      const int t = e / Pn, d = e - t * Pn;
      double v = 0.0;
      if (t < w && d < D) {
        const int src = start + t - pad;
        if (src >= 0 && src < N) v = y[(long long)src * D + d];
      }
      Xs[e] = v;
    }
    __syncthreads();
    for (int d = tid; d < Dp; d += blockDim.x) {
      double s = 0.0;
      for (int t = 0; t < w; ++t) s += Xs[t * Pn + d];
      mean[d] = s / w;
    }
    __syncthreads();
    for (int e = tid; e < w * Pn; e += blockDim.x) {
      const int d = e % Pn;
      if (d < D) Xs[e] -= mean[d];
    }
    __syncthreads();
    if (corr) {
      for (int d = tid; d < Dp; d += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < w; ++t) s += Xs[t * Pn + d] * Xs[t * Pn + d];
        vd[d] = sqrt(s * fact);
      }
      __syncthreads();
    }
    // lower tiles only (the Gram matrix is symmetric); each tile is also stored transposed
    for (int t = warp; t < ntiles; t += nwarps) {
      int ib = 0, rem = t;
      while (rem > ib) { rem -= ib + 1; ++ib; }
      const int jb = rem;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // two accumulator pairs: even / odd k4-steps
      const double* xa = Xs + tig * Pn + 8 * ib + gid;
      const double* xb = Xs + tig * Pn + 8 * jb + gid;
      int k0 = 0;
      for (; k0 + 4 < wp; k0 += 8) {
        dmma(c0, c1, xa[k0 * Pn], xb[k0 * Pn]);
        dmma(e0, e1, xa[(k0 + 4) * Pn], xb[(k0 + 4) * Pn]);
      }
      if (k0 < wp) dmma(c0, c1, xa[k0 * Pn], xb[k0 * Pn]);
      c0 += e0;
      c1 += e1;
      const int row = 8 * ib + gid, col = 8 * jb + 2 * tig;
      if (row >= D || col >= D) continue;
      c0 *= fact;
      c1 *= fact;
      const bool has1 = col + 1 < D;
      if (corr) {
        c0 = (c0 == 0.0) ? 0.0 : c0 / (vd[row] * vd[col]);
        if (has1) c1 = (c1 == 0.0) ? 0.0 : c1 / (vd[row] * vd[col + 1]);
      }
      double* dst = o + (long long)row * D + col;
      if (vec && has1) {
        *reinterpret_cast<double2*>(dst) = make_double2(c0, c1);
      } else {
        dst[0] = c0;
        if (has1) dst[1] = c1;
      }
      if (ib != jb) {  // mirror: 8 consecutive rows -> 64-byte segments
        o[(long long)col * D + row] = c0;
        if (has1) o[(long long)(col + 1) * D + row] = c1;
      }
    }
  }
}

// One CTA per (eval location t, window length w).  Rank-deficient windows (w - 2 < D) score -inf
// (scipy: x outside the support of a singular covariance); otherwise the covariance of the w-1 rows
// around t is accumulated with DMMA straight from global memory (y is L2 resident), factorised with
// the blocked tile Cholesky and used for -1/2 (D log 2pi + log det + |L^-1 y_t|^2).
__global__ void __launch_bounds__(256)
window_cv_kernel(const double* __restrict__ y, int N, int D, int w_min, int w_step, int t_min,
                 int t_step, int num_t, double rel_tol, double* __restrict__ table, long long ldt) {
  extern __shared__ __align__(16) double sm[];
  const int Dp = (D + 7) / 8 * 8, Pd = pitch4(Dp), nbD = Dp / 8;
  double* T = sm;                 // Dp x Pd
  double* Yb = T + Dp * Pd;       // Dp x 12
  double* scr = Yb + Dp * 12;     // 8 x 96
  double* dinv = scr + 8 * 96;    // Dp
  double* mean = dinv + Dp;       // Dp
  double* scal = mean + Dp;       // 2
  const int ti = blockIdx.x, wi = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = 8;
  const int gid = lane >> 2, tig = lane & 3;
  const int w = w_min + wi * w_step, t = t_min + ti * t_step;
  double* dst = table + (long long)ti * ldt + wi;
  const int rows = w - 1;
  if (rows - 1 < D) {  // rank <= w - 2 < D
    if (tid == 0) *dst = -INFINITY;
    return;
  }
  const int start = t - w / 2;            // int(t - floor(w/2))
  const int removed = start + (w - 1) / 2;  // np.delete(window, int((w-1)/2))
  // row index map: k-th kept row -> start + k + (k >= removed - start)
  for (int d = tid; d < Dp; d += blockDim.x) {
    double s = 0.0;
    if (d < D)
      for (int k = 0; k < rows; ++k) {
        const int r = start + k + (start + k >= removed ? 1 : 0);
        s += y[(long long)r * D + d];
      }
    mean[d] = s / rows;
    Yb[d * 12] = d < D ? y[(long long)t * D + d] : 0.0;
#pragma unroll
    for (int c = 1; c < 12; ++c) Yb[d * 12 + c] = 0.0;
  }
  __syncthreads();
  const double fact = 1.0 / (double)(rows - 1);
  const int kp = (rows + 3) / 4 * 4;
  for (int tl = warp; tl < nbD * (nbD + 1) / 2; tl += nwarps) {
    int ib = 0, rem = tl;
    while (rem > ib) { rem -= ib + 1; ++ib; }
    const int jb = rem;
    const int ca = 8 * ib + gid, cb = 8 * jb + gid;
    const double ma = mean[ca], mb = mean[cb];
    const bool oka = ca < D, okb = cb < D;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < kp; k0 += 4) {
      const int k = k0 + tig;
      double a = 0.0, b = 0.0;
      if (k < rows) {
        const int r = start + k + (start + k >= removed ? 1 : 0);
        const double* yr = y + (long long)r * D;
        if (oka) a = yr[ca] - ma;
        if (okb) b = yr[cb] - mb;
      }
      dmma(c0, c1, a, b);
    }
    const int row = 8 * ib + gid, col = 8 * jb + 2 * tig;
    c0 *= fact;
    c1 *= fact;
    if (row >= D) { c0 = (row == col) ? 1.0 : 0.0; c1 = (row == col + 1) ? 1.0 : 0.0; }  // identity padding
    T[row * Pd + col] = c0;
    T[row * Pd + col + 1] = c1;
  }
  __syncthreads();
  // largest diagonal entry (scale for the singularity threshold)
  if (warp == 0) {
    double mx = 0.0;
    for (int d = lane; d < D; d += 32) mx = fmax(mx, T[d * Pd + d]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) scal[0] = mx;
  }
  __syncthreads();
  int bad = blocked_cholesky(T, dinv, Pd, nbD, warp, lane, nwarps, 0);
  if (warp == 0) {
    solve_yblock(T, dinv, Pd, nbD, Yb, scr, lane);
    double q = 0.0, ld = 0.0, mn = INFINITY;
    for (int d = lane; d < D; d += 32) {
      q += Yb[d * 12] * Yb[d * 12];
      const double l = 1.0 / dinv[d];
      ld += log(l);
      mn = fmin(mn, l * l);
    }
    q = warp_sum(q);
    ld = warp_sum(ld);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) {
      double v = -0.5 * (D * 1.8378770664093453 + 2.0 * ld + q);
      if (bad || !(mn > rel_tol * scal[0])) v = -INFINITY;  // numerically rank deficient
      *dst = v;
    }
  }
}

}  // namespace fcest

using namespace fcest;

extern "C" int fcest_window_cov(const double* y, int N, int D, int window_length, int step_size, int corr,
                                double* out, cudaStream_t stream) {
  if (N <= 0 || D <= 0) return 0;
  if (window_length < 2 || step_size < 1) return FCEST_ERR_ARGUMENT;
  const int Dp = (D + 7) / 8 * 8, Pn = pitch4(Dp), wp = (window_length + 3) / 4 * 4;
  const size_t smem = sizeof(double) * ((size_t)wp * Pn + 2 * Dp);
  if (smem > 227 * 1024) return FCEST_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(window_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int nwin = N / step_size;
  // one resident wave: CTAs per SM limited by shared memory (and 8 x 256 threads), windows dealt round-robin
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int slots = 148 * per_sm;
  const int per_cta = (N + slots - 1) / slots;
  const int grid = (N + per_cta - 1) / per_cta;
  window_cov_kernel<<<grid, 256, smem, stream>>>(y, N, D, window_length, step_size, nwin, corr, out);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_window_cv(const double* y, int N, int D, int w_min, int w_step, int num_w, int t_min,
                               int t_step, int num_t, double rel_tol, double* table, long long ldt,
                               cudaStream_t stream) {
  if (num_w <= 0 || num_t <= 0) return 0;
  const int w_max = w_min + (num_w - 1) * w_step;
  if (w_min < 3 || t_min - w_max / 2 < 0 || t_min + (num_t - 1) * t_step + (w_max + 1) / 2 > N)
    return FCEST_ERR_ARGUMENT;
  const int Dp = (D + 7) / 8 * 8, Pd = pitch4(Dp);
  const size_t smem = sizeof(double) * ((size_t)Dp * Pd + 12 * (size_t)Dp + 8 * 96 + 2 * Dp + 8);
  if (smem > 227 * 1024) return FCEST_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(window_cv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(num_t, num_w);
  window_cv_kernel<<<grid, 256, smem, stream>>>(y, N, D, w_min, w_step, t_min, t_step, num_t, rel_tol, table, ldt);
  FCEST_CHECK_LAUNCH();
  return 0;
}
