// Cross-validated window-length search, tile kernel (D <= 103): SlidingWindows.find_cross_validated_optimal_window_length
// (fcest/models/sliding_windows.py:199-274): for every (window length w, location t) the covariance of the w - 1
// rows around t (centre row deleted, np.cov) and the zero-mean MVN log density of the held-out row.
//
// A CTA owns one window length and a block of G consecutive locations.  The centred Gram matrix
// M = sum_rows u u^T (u = y - c, c = column means of the block's first window) stays in registers across the
// block: accumulated directly for the first location, then moved to the next location by ONE DMMA k-step per
// tile, the rank-4 update  + u[entering] - u[leaving] + u[old centre] - u[new centre]   (k = 4 exactly).
// cov = (M - n d d^T) / (n - 1) with d = mean of u goes to swizzled 8 x 8 shared-memory tiles with the held-out
// row as bordered row D (unit pivot): the left-looking blocked Cholesky (lik_tiles.cuh) then yields
// z = L^-1 y_t as row D of L.  A scale guard (max_d M_dd within kCvGuard across the block) falls back to the
// direct accumulation for every location.  Rank-deficient windows (w - 2 < D) score -inf without work, as do
// numerically singular ones (smallest pivot^2 below rel_tol * largest variance; scipy: eigenvalue cut-off).
#include <math.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"
#include "lik_tiles.cuh"

namespace fcest {

constexpr int CV_WARPS = 8;
constexpr int CV_MAXT = 12;      // tiles per warp: NB <= 13 (D <= 103)
constexpr double kCvGuard = 32.0;

__global__ void __launch_bounds__(CV_WARPS * 32, 2)
window_cv_tiles_kernel(const double* __restrict__ y, int N, int D, int w_min, int w_step, int wi0, int t_min,
                       int t_step, int num_t, double rel_tol, double* __restrict__ table, long long ldt, int G) {
  extern __shared__ __align__(16) double sm[];
  const int NB = D / 8 + 1, Dp = 8 * NB, NT = NB * (NB + 1) / 2;
  double* tiles = sm;              // NT x 64
  double* ldiag = tiles + NT * 64; // 64: factor block of the last diagonal tile (holds the tail of z)
  double* dinv = ldiag + 64;       // Dp
  double* cvec = dinv + Dp;        // Dp: block centre
  double* svec = cvec + Dp;        // Dp: column sums of u over the current window
  double* red = svec + Dp;         // 2 x 16: max_d M_dd per location (guard), max variance (threshold)
  __shared__ int direct_all;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int oA0 = swz(gid, tig), oA1 = swz(gid, 4 + tig), oC = swz(gid, 2 * tig);
  const int wi = wi0 + blockIdx.y, w = w_min + wi * w_step;
  const int ti0 = blockIdx.x * G, Gc = min(G, num_t - ti0);
  const int n = w - 1;                       // rows per window (centre deleted)
  const double fact = 1.0 / (double)(n - 1), dn = (double)n;
  const int yb = D >> 3, yr = D & 7;
  const int t0 = t_min + ti0 * t_step;
  // kept row k of the window around t: start + k + (k >= rem - start), start = t - w/2, rem = start + (w-1)/2
  const int hw = w / 2, rm = (w - 1) / 2;

  // block centre c = column means of the first window; svec = 0 by construction
  for (int d = tid; d < Dp; d += blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    if (d < D) {
      const int start = t0 - hw;
      int k = 0;
      for (; k + 1 < n; k += 2) {
        const int r0 = start + k + (k >= rm ? 1 : 0), r1 = start + k + 1 + (k + 1 >= rm ? 1 : 0);
        s0 += y[(long long)r0 * D + d];
        s1 += y[(long long)r1 * D + d];
      }
      if (k < n) s0 += y[(long long)(start + k + (k >= rm ? 1 : 0)) * D + d];
    }
    cvec[d] = (s0 + s1) / dn;
    svec[d] = 0.0;
  }
  if (tid < 32) reinterpret_cast<unsigned long long*>(red)[tid] = 0ull;
  __syncthreads();
  // scale guard: max_d sum_rows u^2 per location of the block (direct; n adds per column and location are cheap
  // next to the factorisation) -> largest / smallest ratio decides incremental vs direct accumulation
  for (int e = tid; e < Gc * D; e += blockDim.x) {
    const int g = e / D, d = e - g * D;
    const int start = t0 + g * t_step - hw;
    const double c = cvec[d];
    double q0 = 0.0, q1 = 0.0;
    int k = 0;
    for (; k + 1 < n; k += 2) {
      const double a = y[(long long)(start + k + (k >= rm ? 1 : 0)) * D + d] - c;
      const double b = y[(long long)(start + k + 1 + (k + 1 >= rm ? 1 : 0)) * D + d] - c;
      q0 += a * a; q1 += b * b;
    }
    if (k < n) { const double a = y[(long long)(start + k + (k >= rm ? 1 : 0)) * D + d] - c; q0 += a * a; }
    atomicMax(reinterpret_cast<unsigned long long*>(red) + g, (unsigned long long)__double_as_longlong(q0 + q1));
  }
  __syncthreads();
  if (tid == 0) {
    double lo = INFINITY, hi = 0.0;
    for (int g = 0; g < Gc; ++g) { const double v = red[g]; lo = fmin(lo, v); hi = fmax(hi, v); }
    direct_all = (t_step != 1) || !(hi <= kCvGuard * lo);
  }
  __syncthreads();
  const bool all_direct = direct_all != 0;

  // owned tiles: t = warp + 8 q
  int tib[CV_MAXT], tjb[CV_MAXT];
  double acc[CV_MAXT][2];
#pragma unroll
  for (int q = 0; q < CV_MAXT; ++q) {
    const int t = warp + q * CV_WARPS;
    const int tt = t < NT ? t : 0;
    int ib = (int)((sqrtf(8.0f * (float)tt + 1.0f) - 1.0f) * 0.5f);
    if ((ib + 1) * (ib + 2) / 2 <= tt) ++ib;
    if (ib * (ib + 1) / 2 > tt) --ib;
    tib[q] = t < NT ? ib : -1;
    tjb[q] = tt - ib * (ib + 1) / 2;
    acc[q][0] = acc[q][1] = 0.0;
  }

  for (int g = 0; g < Gc; ++g) {
    const int t = t0 + g * t_step;
    const int start = t - hw;
    if (g == 0 || all_direct) {
      // direct accumulation over the kept rows of window(t)
#pragma unroll
      for (int q = 0; q < CV_MAXT; ++q) acc[q][0] = acc[q][1] = 0.0;
      for (int k0 = 0; k0 < n; k0 += 4) {
        const int k = k0 + tig;
        const bool live = k < n;
        const double* yr_ = y + (long long)(start + k + (k >= rm ? 1 : 0)) * D;
#pragma unroll
        for (int q = 0; q < CV_MAXT; ++q) {
          if (tib[q] < 0) continue;
          const int ca = 8 * tib[q] + gid, cb = 8 * tjb[q] + gid;
          const double a = (live && ca < D) ? yr_[ca] - cvec[ca] : 0.0;
          const double b = (live && cb < D) ? yr_[cb] - cvec[cb] : 0.0;
          dmma(acc[q][0], acc[q][1], a, b);
        }
      }
      __syncthreads();
      for (int d = tid; d < Dp; d += blockDim.x) {
        double s0 = 0.0;
        if (d < D) {
          const double c = cvec[d];
          for (int k = 0; k < n; ++k) s0 += y[(long long)(start + k + (k >= rm ? 1 : 0)) * D + d] - c;
        }
        svec[d] = s0;
      }
    } else {
      // window(t) from window(t - 1): + row (start + w - 1), - row (start - 1), + old centre (rem - 1), - new centre (rem)
      const int rem = start + rm;
      const int rrow = (tig == 0) ? start + w - 1 : (tig == 1) ? start - 1 : (tig == 2) ? rem - 1 : rem;
      const double sgn = (tig & 1) ? -1.0 : 1.0;
      const double* yr_ = y + (long long)rrow * D;
#pragma unroll
      for (int q = 0; q < CV_MAXT; ++q) {
        if (tib[q] < 0) continue;
        const int ca = 8 * tib[q] + gid, cb = 8 * tjb[q] + gid;
        const double a = (ca < D) ? sgn * (yr_[ca] - cvec[ca]) : 0.0;
        const double b = (cb < D) ? yr_[cb] - cvec[cb] : 0.0;
        dmma(acc[q][0], acc[q][1], a, b);
      }
      __syncthreads();
      for (int d = tid; d < D; d += blockDim.x) {
        svec[d] += (y[(long long)(start + w - 1) * D + d] - y[(long long)(start - 1) * D + d]) +
                   (y[(long long)(rem - 1) * D + d] - y[(long long)rem * D + d]);
      }
    }
    if (tid == 0) red[16] = 0.0;
    __syncthreads();
    // covariance tiles + bordered held-out row -> shared memory; largest variance
    double mx = 0.0;
#pragma unroll
    for (int q = 0; q < CV_MAXT; ++q) {
      if (tib[q] < 0) continue;
      const int i = 8 * tib[q] + gid;
      double v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * tjb[q] + 2 * tig + e;
        double x;
        if (i < D) {
          x = (j < D) ? (acc[q][e] - svec[i] * svec[j] / dn) * fact : 0.0;
          if (i == j) mx = fmax(mx, x);
        } else if (i == D) {
          x = (j < D) ? y[(long long)t * D + j] : (j == D ? 1.0 : 0.0);
        } else {
          x = (j == i) ? 1.0 : 0.0;
        }
        v[e] = x;
      }
      *reinterpret_cast<double2*>(tiles + TI(tib[q], tjb[q]) * 64 + oC) = make_double2(v[0], v[1]);
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(red) + 16, (unsigned long long)__double_as_longlong(mx));
    __syncthreads();
    // left-looking blocked Cholesky (NB <= 13: at most two tiles of a block column per warp)
    int bad = 0;
    for (int jb = 0; jb < NB; ++jb) {
      {
        double c[2][2];
        int ibs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          ibs[u] = jb + warp + u * CV_WARPS;
          c[u][0] = c[u][1] = 0.0;
          if (ibs[u] < NB) {
            const double2 vv = *reinterpret_cast<const double2*>(tiles + TI(ibs[u], jb) * 64 + oC);
            c[u][0] = vv.x; c[u][1] = vv.y;
          }
        }
        if (ibs[0] < NB) {
          for (int kb = 0; kb < jb; ++kb) {
            const double* Tj = tiles + TI(jb, kb) * 64;
            const double nb0 = -Tj[oA0], nb1 = -Tj[oA1];
            dmma(c[0][0], c[0][1], tiles[TI(ibs[0], kb) * 64 + oA0], nb0);
            if (ibs[1] < NB) dmma(c[1][0], c[1][1], tiles[TI(ibs[1], kb) * 64 + oA0], nb0);
            dmma(c[0][0], c[0][1], tiles[TI(ibs[0], kb) * 64 + oA1], nb1);
            if (ibs[1] < NB) dmma(c[1][0], c[1][1], tiles[TI(ibs[1], kb) * 64 + oA1], nb1);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (ibs[u] < NB) *reinterpret_cast<double2*>(tiles + TI(ibs[u], jb) * 64 + oC) = make_double2(c[u][0], c[u][1]);
        }
      }
      __syncthreads();
      if (warp == 0)
        bad = factor_diag_swz(tiles + TI(jb, jb) * 64, dinv + 8 * jb, lane, (jb == yb) ? yr : -1, 8 * jb, D, bad,
                              (jb == yb) ? ldiag : nullptr);
      __syncthreads();
      {
        const double* Ij = tiles + TI(jb, jb) * 64;
        const double b0 = Ij[oA0], b1 = Ij[oA1];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int ib = jb + 1 + warp + u * CV_WARPS;
          if (ib < NB) {
            const double* X = tiles + TI(ib, jb) * 64;
            double r0 = 0.0, r1 = 0.0;
            const double x0 = X[oA0], x1 = X[oA1];
            dmma(r0, r1, x0, b0);
            dmma(r0, r1, x1, b1);
            __syncwarp();
            *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(r0, r1);
          }
        }
      }
      __syncthreads();
    }
    if (warp == 0) {
      // z = row D of L: off-diagonal tiles (yb, jb < yb) + the factor block of the last diagonal tile
      double qd = 0.0, ld = 0.0, mn = INFINITY;
      for (int j = lane; j < D; j += 32) {
        const int jb = j >> 3;
        const double z = (jb < yb) ? tiles[TI(yb, jb) * 64 + swz(yr, j & 7)] : ldiag[yr * 8 + (j & 7)];
        qd += z * z;
        const double l = 1.0 / dinv[j];
        ld += log(l);
        mn = fmin(mn, l * l);
      }
      qd = warp_sum(qd);
      ld = warp_sum(ld);
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o2));
      if (lane == 0) {
        double v = -0.5 * (D * 1.8378770664093453 + 2.0 * ld + qd);
        if (bad || !(mn > rel_tol * red[16])) v = -INFINITY;  // numerically rank deficient
        table[(long long)(ti0 + g) * ldt + wi] = v;
      }
    }
    __syncthreads();
  }
}

bool window_cv_tiles_supported(int D) { return D >= 1 && D / 8 + 1 <= 13; }

// launches the tile kernel for the window lengths wi >= wi0 (the non-singular ones); returns 0 / cudaError_t
int launch_window_cv_tiles(const double* y, int N, int D, int w_min, int w_step, int wi0, int num_w, int t_min,
                           int t_step, int num_t, double rel_tol, double* table, long long ldt, cudaStream_t stream) {
  const int NB = D / 8 + 1, NT = NB * (NB + 1) / 2, Dp = 8 * NB;
  const size_t smem = sizeof(double) * ((size_t)NT * 64 + 64 + 3 * (size_t)Dp + 32);
  cudaError_t e = cudaFuncSetAttribute(window_cv_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int G = (t_step == 1) ? 16 : 1;
  dim3 grid((num_t + G - 1) / G, num_w - wi0);
  window_cv_tiles_kernel<<<grid, CV_WARPS * 32, smem, stream>>>(y, N, D, w_min, w_step, wi0, t_min, t_step, num_t,
                                                                rel_tol, table, ldt, G);
  return 0;
}

}  // namespace fcest
