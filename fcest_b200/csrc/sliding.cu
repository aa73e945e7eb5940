// Sliding-window covariance estimator and its cross-validated window-length search.
//
// Reference: fcest/models/sliding_windows.py:143-174 (overlapping_windowed_cov_estimation: zero padding
// of floor(w/2) rows at both ends, np.cov per window = two-pass centred X^T X * 1/(w-1), optional
// cov -> corr with corr[cov == 0] = 0, fcest/helpers/array_operations.py:37-53) and :219-272
// (leave-centre-out window covariance + zero-mean MVN log density of the held-out row).
//
// window_cov: output traffic (8 N D^2 bytes, 96 MB at N=1200, D=100) bounds the estimator.  Two kernels share the
// anchored prefix-sum formulation described below (a block of consecutive windows per CTA, re-anchored by a direct
// two-pass accumulation at every block start -- a GLOBAL prefix sum would not survive the reference's 1e-12
// tolerance on non-centred data):
//   window_cov_bulk_kernel  step 1, even D <= 112: windows assembled in shared memory and written with cp.async.bulk
//   window_cov_kernel       everything else: 64-byte row segments straight from the accumulator registers
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

// Anchored prefix-sum formulation.  A CTA owns a block of G consecutive windows (step 1) and keeps the Gram
// matrix M_g = sum_{t in window g} u_t u_t^T of the block-centred rows u_t = x_t - c (c = column means of the
// block's first window) in registers (16 x 16 macro tiles of 2 x 2 DMMA tiles, lower macro tiles only):
//   window 0 of the block : M_0 accumulated directly over its w rows (two-pass, as np.cov does);
//   window g > 0          : M_g = M_{g-1} + u_new u_new^T - u_old u_old^T   -- ONE DMMA k-step per tile;
//   cov_g = (M_g - w d_g d_g^T) / (w - 1),  d_g = window mean of u (exactly the centred second moment).
// Re-anchoring every G windows bounds the accumulated rounding at ~2 G eps |M|; a guard falls back to the
// direct accumulation for every window of a block whose windows differ in scale (max_d M_dd) by more than
// kScaleGuard, which keeps the 1e-12 relative-to-max contract on badly scaled data too.
constexpr int WC_MAXM = 4;               // macro tiles held per warp
constexpr double kScaleGuard = 32.0;

__global__ void __launch_bounds__(256, 2)
window_cov_kernel(const double* __restrict__ y, int N, int D, int w, int step, int nwin, int corr, int G,
                  double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const int Dm = (D + 15) / 16 * 16, Pn = pitch4(Dm);
  const int RB = w + (G - 1) * step, RBp = (RB + 3) / 4 * 4;
  double* Xs = sm;                  // RBp x Pn: block rows, centred in place (zero padded rows / columns)
  double* delta = Xs + RBp * Pn;    // G x Dm: window mean of the centred rows
  double* isd = delta + G * Dm;     // G x Dm: 1 / sqrt(var) (0 for a constant column)
  double* red = isd + G * Dm;       // 16: max_d sum_t u^2 per window (scale guard)
  __shared__ int direct_all;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int pad = w / 2;
  const double fact = 1.0 / (double)(w - 1);
  const bool vec = ((D & 1) == 0);
  // rows >= int(N / step) stay zero (sliding_windows.py:164-165)
  for (long long e = (long long)nwin * D * D + (long long)blockIdx.x * blockDim.x + tid; e < (long long)N * D * D;
       e += (long long)gridDim.x * blockDim.x) out[e] = 0.0;
  const int g0 = blockIdx.x * G;
  if (g0 >= nwin) return;
  const int Gc = min(G, nwin - g0);
  const int start = g0 * step - pad;
  // stage the block's rows: 16-byte cp.async copies (zero-filled outside the data / past column D) when
  // rows are 16-byte aligned (D even), all in flight at once
  if (vec) {
    const int h = Pn / 2;
    for (int e = tid; e < RBp * h; e += blockDim.x) {
      const int t = e / h, d = 2 * (e - t * h);
      const int src = start + t;
      const bool live = t < RB && src >= 0 && src < N && d < D;
      cp_async16(Xs + t * Pn + d, live ? y + (long long)src * D + d : y, live ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else {
    for (int t = warp; t < RBp; t += nwarps) {
      const int src = start + t;
      const bool live = t < RB && src >= 0 && src < N;
      const double* yr = y + (long long)(live ? src : 0) * D;
      for (int d = lane; d < Pn; d += 32) Xs[t * Pn + d] = (live && d < D) ? yr[d] : 0.0;
    }
  }
  if (tid < 16) reinterpret_cast<unsigned long long*>(red)[tid] = 0ull;
  __syncthreads();
  // per column: block centre c (mean of the block's first window), centre every row in place
  for (int d = tid; d < Dm; d += blockDim.x) {
    double* xc = Xs + d;
    double s0 = 0.0, s1 = 0.0;
    int t = 0;
    for (; t + 1 < w; t += 2) { s0 += xc[t * Pn]; s1 += xc[(t + 1) * Pn]; }
    if (t < w) s0 += xc[t * Pn];
    const double c = (d < D) ? (s0 + s1) / w : 0.0;
    for (t = 0; t < RB; ++t) xc[t * Pn] -= c;
  }
  __syncthreads();
  // per (window, column): two-pass mean offset d_g, variance, scale sum u^2 (-> max over columns per window)
  for (int e = tid; e < Gc * Dm; e += blockDim.x) {
    const int g = e / Dm, d = e - g * Dm;
    const double* xw = Xs + (long long)g * step * Pn + d;
    double a0 = 0.0, a1 = 0.0;
    int t = 0;
    for (; t + 1 < w; t += 2) { a0 += xw[t * Pn]; a1 += xw[(t + 1) * Pn]; }
    if (t < w) a0 += xw[t * Pn];
    const double dl = (a0 + a1) / w;
    double q0 = 0.0, q1 = 0.0, r0 = 0.0, r1 = 0.0;
    for (t = 0; t + 1 < w; t += 2) {
      const double ua = xw[t * Pn], ub = xw[(t + 1) * Pn];
      const double va = ua - dl, vb = ub - dl;
      q0 += va * va; q1 += vb * vb; r0 += ua * ua; r1 += ub * ub;
    }
    if (t < w) { const double ua = xw[t * Pn], va = ua - dl; q0 += va * va; r0 += ua * ua; }
    const double var = (q0 + q1) * fact;
    delta[e] = dl;
    isd[e] = var > 0.0 ? 1.0 / sqrt(var) : 0.0;
    // non-negative doubles order like their bit patterns: deterministic max without a reduction tree
    atomicMax(reinterpret_cast<unsigned long long*>(red) + g, (unsigned long long)__double_as_longlong(r0 + r1));
  }
  __syncthreads();
  if (tid == 0) {
    double lo = INFINITY, hi = 0.0;
    for (int g = 0; g < Gc; ++g) { const double v = red[g]; lo = fmin(lo, v); hi = fmax(hi, v); }
    direct_all = (step != 1) || !(hi <= kScaleGuard * lo);
  }
  __syncthreads();
  const bool all_direct = direct_all != 0;
  const int nbm = Dm / 16, nmac = nbm * (nbm + 1) / 2;
  const int w4 = w & ~3;
  // macro tiles in chunks of (warps x WC_MAXM); D <= 112 needs one chunk
  for (int mbase = 0; mbase < nmac; mbase += nwarps * WC_MAXM) {
    int Is[WC_MAXM], Js[WC_MAXM];
    bool on[WC_MAXM];
#pragma unroll
    for (int q = 0; q < WC_MAXM; ++q) {
      const int mt = mbase + warp + q * nwarps;
      on[q] = mt < nmac;
      int I = 0, rem = on[q] ? mt : 0;
      while (rem > I) { rem -= I + 1; ++I; }
      Is[q] = I; Js[q] = rem;
    }
    double acc[WC_MAXM][2][2][2];
    for (int g = 0; g < Gc; ++g) {
      const double* Xw = Xs + (long long)g * step * Pn;
      if (g == 0 || all_direct) {
#pragma unroll
        for (int q = 0; q < WC_MAXM; ++q) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int v = 0; v < 2; ++v) acc[q][u][v][0] = acc[q][u][v][1] = 0.0;
          if (!on[q]) continue;
          const double* pa = Xw + tig * Pn + 16 * Is[q] + gid;
          const double* pb = Xw + tig * Pn + 16 * Js[q] + gid;
          for (int k0 = 0; k0 < w4; k0 += 4) {
            const double a0 = pa[0], a1 = pa[8], b0 = pb[0], b1 = pb[8];
            dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
            dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
            dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
            dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
            pa += 4 * Pn;
            pb += 4 * Pn;
          }
          if (w4 < w) {  // ragged last k-step: rows >= w belong to later windows
            const bool live = w4 + tig < w;
            const double a0 = live ? pa[0] : 0.0, a1 = live ? pa[8] : 0.0, b0 = live ? pb[0] : 0.0, b1 = live ? pb[8] : 0.0;
            dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
            dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
            dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
            dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
          }
        }
      } else {
        // M_g = M_{g-1} + u_new u_new^T - u_old u_old^T: k-slot 0 = entering row, 1 = leaving row, 2, 3 = 0
        const double* rnew = Xw + (w - 1) * Pn;   // last row of window g
        const double* rold = Xw - Pn;             // first row of window g - 1
        const double* rsel = (tig == 0) ? rnew : rold;
        const double sgn = (tig == 0) ? 1.0 : -1.0;
#pragma unroll
        for (int q = 0; q < WC_MAXM; ++q) {
          if (!on[q]) continue;
          double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
          if (tig < 2) {
            a0 = rsel[16 * Is[q] + gid]; a1 = rsel[16 * Is[q] + 8 + gid];
            b0 = rsel[16 * Js[q] + gid]; b1 = rsel[16 * Js[q] + 8 + gid];
            a0 *= sgn; a1 *= sgn;
          }
          dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
          dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
          dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
          dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
        }
      }
      // epilogue of window g (accumulators stay live for the next window)
      double* o = out + (long long)(g0 + g) * D * D;
      const double* dg = delta + g * Dm;
      const double* ig = isd + g * Dm;
      const double wd = (double)w;
#pragma unroll
      for (int q = 0; q < WC_MAXM; ++q) {
        if (!on[q]) continue;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int ib = 2 * Is[q] + u, jb = 2 * Js[q] + v;
            if (jb > ib) continue;  // upper tile of a diagonal macro tile
            const int row = 8 * ib + gid, col = 8 * jb + 2 * tig;
            if (row >= D || col >= D) continue;
            const bool has1 = col + 1 < D;
            const double dr = dg[row];
            double c0 = (acc[q][u][v][0] - wd * dr * dg[col]) * fact;
            double c1 = (acc[q][u][v][1] - wd * dr * dg[col + 1]) * fact;
            if (corr) {
              c0 = (c0 == 0.0) ? 0.0 : c0 * (ig[row] * ig[col]);
              c1 = (c1 == 0.0) ? 0.0 : c1 * (ig[row] * ig[col + 1]);
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
    }
  }
}

// ---- bulk-store variant (step 1, even D <= 112): the default at the reference's sizes --------------------------
// The register-store kernel above spends ~8 000 warp instructions per window, most of them address arithmetic and
// predicates of the per-tile global stores, at two warps per scheduler (ncu: issue slots 23 % busy, 74 k cycles
// for ~8 windows per SM).  Here a CTA of 16 warps per SM owns an even contiguous share of the windows; a finished
// window is assembled (both triangles) in one of `nbuf` D x D shared-memory buffers with all offsets hoisted out of
// the window loop, and leaves as ONE bulk-copy group (cp.async.bulk shared -> global, 16 KB pieces) issued by a
// single thread: the rank-2 update and assembly of window g + 1 run underneath the store of window g.
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int WB_THREADS = 512;
constexpr int WB_LANES = 8;  // threads that issue the bulk copies of a window
constexpr int WB_MAXM = 2;   // macro tiles per warp: 16 warps x 2 >= 28 macro tiles of D <= 112

template <bool CORR>
__global__ void __launch_bounds__(WB_THREADS, 1)
window_cov_bulk_kernel(const double* __restrict__ y, int N, int D, int w, int G, int nbuf,
                       double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const int Dm = (D + 15) / 16 * 16, Pn = pitch4(Dm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = WB_THREADS / 32;
  const int gid = lane >> 2, tig = lane & 3;
  const int g0 = (int)((long long)blockIdx.x * N / gridDim.x);
  const int Gc = (int)((long long)(blockIdx.x + 1) * N / gridDim.x) - g0;
  if (Gc <= 0) return;
  const int RB = w + Gc - 1, RBp = (w + G - 1 + 3) / 4 * 4;
  double* Obuf = sm;                               // nbuf x (D x D)
  double* Xs = sm + (size_t)nbuf * D * D;          // RBp x Pn block rows, centred in place
  double* delta = Xs + (size_t)RBp * Pn;           // G x Dm window mean of the centred rows
  double* isd = delta + (size_t)G * Dm;            // G x Dm 1 / sqrt(var) (CORR)
  double* red = isd + (size_t)G * Dm;              // 16: scale guard
  double* cen = red + 16;                          // 4 x Dm partial column sums
  __shared__ int direct_all;
  const int pad = w / 2, start = g0 - pad;
  const double fact = 1.0 / (double)(w - 1), wd = (double)w;
  // stage the block's rows, one warp per row: 16-byte cp.async copies (zero-filled outside the data / past D)
  for (int t = warp; t < RBp; t += nwarps) {
    const int src = start + t;
    const bool rowlive = t < RB && src >= 0 && src < N;
    const double* yr = y + (long long)(rowlive ? src : 0) * D;
    for (int d = 2 * lane; d < Pn; d += 64) {
      const bool live = rowlive && d < D;
      cp_async16(Xs + t * Pn + d, live ? yr + d : y, live ? 16 : 0);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  if (tid < 16) reinterpret_cast<unsigned long long*>(red)[tid] = 0ull;
  __syncthreads();
  // block centre c[d] ~ mean of the block's first window: only an anchor (the result does not depend on it),
  // so four partial sums per column in any order will do
  if (tid < 4 * 128) {
    const int part = tid >> 7, d = tid & 127;
    if (d < Dm) {
      double s = 0.0;
      for (int t = part; t < w; t += 4) s += Xs[t * Pn + d];
      cen[part * Dm + d] = s;
    }
  }
  __syncthreads();
  {
    double c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = lane + 32 * i;
      c[i] = (d < D) ? (cen[d] + cen[Dm + d] + cen[2 * Dm + d] + cen[3 * Dm + d]) / wd : 0.0;
    }
    for (int t = warp; t < RB; t += nwarps) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = lane + 32 * i;
        if (d < Dm) Xs[t * Pn + d] -= c[i];
      }
    }
  }
  __syncthreads();
  // per (window, column): mean offset d_g, two-pass variance (CORR), scale sum u^2 (-> max over columns)
  const unsigned inv_nbm = (65536u + (unsigned)(Dm >> 4) - 1u) / (unsigned)(Dm >> 4);  // exact e / Dm for e < 2048
  for (int e = tid; e < Gc * Dm; e += WB_THREADS) {
    const int g = (int)((((unsigned)e >> 4) * inv_nbm) >> 16), d = e - g * Dm;
    const double* xw = Xs + (size_t)g * Pn + d;
    double a0 = 0.0, a1 = 0.0, r0 = 0.0, r1 = 0.0;
    int t = 0;
    for (; t + 1 < w; t += 2) {
      const double ua = xw[t * Pn], ub = xw[(t + 1) * Pn];
      a0 += ua; a1 += ub; r0 += ua * ua; r1 += ub * ub;
    }
    if (t < w) { const double ua = xw[t * Pn]; a0 += ua; r0 += ua * ua; }
    const double dl = (a0 + a1) / wd;
    delta[e] = dl;
    if (CORR) {
      double q0 = 0.0, q1 = 0.0;
      for (t = 0; t + 1 < w; t += 2) {
        const double va = xw[t * Pn] - dl, vb = xw[(t + 1) * Pn] - dl;
        q0 += va * va; q1 += vb * vb;
      }
      if (t < w) { const double va = xw[t * Pn] - dl; q0 += va * va; }
      const double var = (q0 + q1) * fact;
      isd[e] = var > 0.0 ? 1.0 / sqrt(var) : 0.0;
    }
    atomicMax(reinterpret_cast<unsigned long long*>(red) + g, (unsigned long long)__double_as_longlong(r0 + r1));
  }
  __syncthreads();
  if (tid == 0) {
    double lo = INFINITY, hi = 0.0;
    for (int g = 0; g < Gc; ++g) { const double v = red[g]; lo = fmin(lo, v); hi = fmax(hi, v); }
    direct_all = !(hi <= kScaleGuard * lo);
  }
  __syncthreads();
  const bool all_direct = direct_all != 0;
  const int nbm = Dm / 16, nmac = nbm * (nbm + 1) / 2;
  const int w4 = w & ~3;
  // per-lane constants of this warp's macro tiles (hoisted out of the window loop)
  int xa[WB_MAXM], xb[WB_MAXM];        // column offsets of the A / B fragments in a row of Xs
  int od[WB_MAXM], om[WB_MAXM];        // element offsets in a D x D buffer: direct (row, col), mirror (col, row)
  bool on[WB_MAXM], offd[WB_MAXM], rok[WB_MAXM][2], cok[WB_MAXM][2];
#pragma unroll
  for (int q = 0; q < WB_MAXM; ++q) {
    const int mt = warp + q * nwarps;
    on[q] = mt < nmac;
    int I = 0, rem = on[q] ? mt : 0;
    while (rem > I) { rem -= I + 1; ++I; }
    const int J = rem;
    xa[q] = 16 * I + gid; xb[q] = 16 * J + gid;
    const int row = 16 * I + gid, col = 16 * J + 2 * tig;
    od[q] = row * D + col; om[q] = col * D + row;
    offd[q] = I != J;
    rok[q][0] = row < D; rok[q][1] = row + 8 < D;
    cok[q][0] = col < D; cok[q][1] = col + 8 < D;
  }
  double acc[WB_MAXM][2][2][2];
  for (int g = 0; g < Gc; ++g) {
    const double* Xw = Xs + (size_t)g * Pn;
    if (g == 0 || all_direct) {
#pragma unroll
      for (int q = 0; q < WB_MAXM; ++q) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int v = 0; v < 2; ++v) acc[q][u][v][0] = acc[q][u][v][1] = 0.0;
        if (!on[q]) continue;
        const double* pa = Xw + tig * Pn + xa[q];
        const double* pb = Xw + tig * Pn + xb[q];
        for (int k0 = 0; k0 < w4; k0 += 4) {
          const double a0 = pa[0], a1 = pa[8], b0 = pb[0], b1 = pb[8];
          dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
          dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
          dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
          dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
          pa += 4 * Pn;
          pb += 4 * Pn;
        }
        if (w4 < w) {  // ragged last k-step: rows >= w belong to later windows
          const bool live = w4 + tig < w;
          const double a0 = live ? pa[0] : 0.0, a1 = live ? pa[8] : 0.0, b0 = live ? pb[0] : 0.0, b1 = live ? pb[8] : 0.0;
          dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
          dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
          dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
          dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
        }
      }
    } else {
      // M_g = M_{g-1} + u_new u_new^T - u_old u_old^T: k-slot 0 = entering row, 1 = leaving row, 2, 3 = 0
      const double* rsel = (tig == 0) ? Xw + (w - 1) * Pn : Xw - Pn;
      const double sgn = (tig == 0) ? 1.0 : -1.0;
#pragma unroll
      for (int q = 0; q < WB_MAXM; ++q) {
        if (!on[q]) continue;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        if (tig < 2) {
          a0 = rsel[xa[q]] * sgn; a1 = rsel[xa[q] + 8] * sgn;
          b0 = rsel[xb[q]]; b1 = rsel[xb[q] + 8];
        }
        dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
        dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
        dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
        dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
      }
    }
    // assemble window g in shared memory (accumulators stay live for the next window).  Buffer g & 1 is free:
    // the store of window g - 2 was waited for before the barrier of window g - 1.
    double* o = Obuf + (size_t)(nbuf == 2 ? (g & 1) : 0) * D * D;
    const double* dg = delta + g * Dm;
    const double* ig = isd + g * Dm;
#pragma unroll
    for (int q = 0; q < WB_MAXM; ++q) {
      if (!on[q]) continue;
      const int ca = xb[q] - gid + 2 * tig;   // first of this lane's columns
      const double2 dc0 = *reinterpret_cast<const double2*>(dg + ca);
      const double2 dc1 = *reinterpret_cast<const double2*>(dg + ca + 8);
      double2 ic0, ic1;
      if (CORR) {
        ic0 = *reinterpret_cast<const double2*>(ig + ca);
        ic1 = *reinterpret_cast<const double2*>(ig + ca + 8);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!rok[q][u]) continue;
        const double tr = wd * dg[xa[q] + 8 * u];
        const double ir = CORR ? ig[xa[q] + 8 * u] : 0.0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (!cok[q][v]) continue;
          const double2 dc = v ? dc1 : dc0;
          double c0 = (acc[q][u][v][0] - tr * dc.x) * fact;
          double c1 = (acc[q][u][v][1] - tr * dc.y) * fact;
          if (CORR) {
            const double2 ic = v ? ic1 : ic0;
            c0 = (c0 == 0.0) ? 0.0 : c0 * (ir * ic.x);
            c1 = (c1 == 0.0) ? 0.0 : c1 * (ir * ic.y);
          }
          *reinterpret_cast<double2*>(o + od[q] + 8 * u * D + 8 * v) = make_double2(c0, c1);
          if (offd[q]) {
            double* m = o + om[q] + 8 * v * D + 8 * u;
            m[0] = c0;
            m[D] = c1;
          }
        }
      }
    }
    fence_async_smem();   // generic-proxy writes of the window -> visible to the bulk-copy engine
    // the store of window g - 1 has had the whole assembly of window g to drain; once it has left shared
    // memory its buffer is free for window g + 1 (with one buffer: wait here, after the store, instead)
    if (nbuf == 2 && tid < WB_LANES) bulk_wait_read<0>();
    __syncthreads();
    if (tid < WB_LANES) {   // lanes 0..7 of warp 0: one piece of the window each
      const uint32_t total = (uint32_t)D * D * 8u;
      const uint32_t piece = ((total / WB_LANES) + 15u) & ~15u;
      const uint32_t off = (uint32_t)tid * piece;
      if (off < total) {
        bulk_store(out + (long long)(g0 + g) * D * D + off / 8, o + off / 8, min(piece, total - off));
        bulk_commit();
      }
      if (nbuf == 1) bulk_wait_read<0>();
    }
    if (nbuf == 1) __syncthreads();
  }
  if (tid < WB_LANES) bulk_wait_read<0>();  // shared memory must outlive the last copies
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

// table[:, wi] = -inf for the rank-deficient window lengths wi < wi0 (w - 2 < D)
__global__ void window_cv_fill_singular_kernel(double* __restrict__ table, long long ldt, int num_t, int wi0) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)num_t * wi0;
       e += (long long)gridDim.x * blockDim.x)
    table[(e / wi0) * ldt + (e % wi0)] = -INFINITY;
}

bool window_cv_tiles_supported(int D);
int launch_window_cv_tiles(const double* y, int N, int D, int w_min, int w_step, int wi0, int num_w, int t_min,
                           int t_step, int num_t, double rel_tol, double* table, long long ldt, cudaStream_t stream);

}  // namespace fcest

using namespace fcest;

extern "C" int fcest_window_cov(const double* y, int N, int D, int window_length, int step_size, int corr,
                                double* out, cudaStream_t stream) {
  if (N <= 0 || D <= 0) return 0;
  if (window_length < 2 || step_size < 1) return FCEST_ERR_ARGUMENT;
  const int nwin = N / step_size;
  const int Dm = (D + 15) / 16 * 16, Pn = pitch4(Dm);
  // Bulk-store kernel: one CTA per SM, an even contiguous share of the windows each (at most 16 per CTA: the
  // anchored update chain restarts at every block), one or two D x D staging buffers.
  static const int use_bulk = getenv("FCEST_WCOV_IMPL") ? strcmp(getenv("FCEST_WCOV_IMPL"), "direct") != 0 : 1;
  if (use_bulk && step_size == 1 && (D & 1) == 0 && Dm <= 112) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = N < sms ? N : sms;
    if ((N + grid - 1) / grid > 16) grid = (N + 15) / 16;
    const int Gt = (N + grid - 1) / grid;
    const int RBp = (window_length + Gt - 1 + 3) / 4 * 4;
    const size_t rest = sizeof(double) * ((size_t)RBp * Pn + 2 * (size_t)Gt * Dm + 16 + 4 * (size_t)Dm);
    const size_t obytes = sizeof(double) * (size_t)D * D;
    const int nbuf = (2 * obytes + rest <= 227 * 1024) ? 2 : ((obytes + rest <= 227 * 1024) ? 1 : 0);
    if (nbuf > 0) {
      const size_t smem_b = nbuf * obytes + rest;
      auto kern = corr ? window_cov_bulk_kernel<true> : window_cov_bulk_kernel<false>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
      if (e != cudaSuccess) return (int)e;
      kern<<<grid, WB_THREADS, smem_b, stream>>>(y, N, D, window_length, Gt, nbuf, out);
      FCEST_CHECK_LAUNCH();
      return 0;
    }
  }
  // windows per CTA: blocks sized so that two CTAs per SM cover all windows in one resident wave
  // (step 1 only; larger steps share too few rows for the incremental update to pay)
  int G = 1;
  if (step_size == 1) {
    G = (nwin + 2 * 148 - 1) / (2 * 148);
    G = G < 2 ? 2 : (G > 16 ? 16 : G);
  }
  size_t smem = 0;
  for (;; --G) {
    const int RB = window_length + (G - 1) * step_size, RBp = (RB + 3) / 4 * 4;
    smem = sizeof(double) * ((size_t)RBp * Pn + 2 * (size_t)G * Dm + 16);
    if (smem <= 110 * 1024 || G == 1) break;
  }
  if (smem > 227 * 1024) return FCEST_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(window_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int grid = (nwin + G - 1) / G;
  if (grid < 1) grid = 1;
  window_cov_kernel<<<grid, 256, smem, stream>>>(y, N, D, window_length, step_size, nwin, corr, G, out);
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
  static const bool force_cta = [] { const char* v = getenv("FCEST_CV_IMPL"); return v && !strcmp(v, "cta"); }();
  if (!force_cta && window_cv_tiles_supported(D) && w_step >= 1) {
    // tile kernel (sliding_cv_tiles.cu): window lengths with w - 2 < D are rank deficient -> -inf, no work
    int wi0 = 0;
    while (wi0 < num_w && (w_min + wi0 * w_step) - 2 < D) ++wi0;
    if (wi0 > 0) {
      window_cv_fill_singular_kernel<<<148, 256, 0, stream>>>(table, ldt, num_t, wi0);
      FCEST_CHECK_LAUNCH();
    }
    if (wi0 < num_w) {
      const int rc = launch_window_cv_tiles(y, N, D, w_min, w_step, wi0, num_w, t_min, t_step, num_t, rel_tol, table, ldt,
                                            stream);
      if (rc != 0) return rc;
      FCEST_CHECK_LAUNCH();
    }
    return 0;
  }
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
