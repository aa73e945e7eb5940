// Tile-level (8x8, DMMA) building blocks for small dense SPD problems held in shared memory:
// blocked right-looking Cholesky with register-resident diagonal-tile factorisation + inversion, and the
// block forward substitution z = L^-1 y.  Shared by the Wishart likelihood kernel (likelihood.cu) and
// the cross-validated sliding-window kernel (sliding.cu).
//
// Storage convention for the Dp x Dp matrix T (row pitch Pd = 4 mod 8 doubles, Dp = 8 nbD):
//   lower triangle (incl. diagonal): Sigma on entry, L = chol(Sigma) on exit;
//   strict upper part of DIAGONAL tile (b,b): (L_bb^-1)^T;  dinv[d] = 1 / L[d][d].
#pragma once
#include "common.cuh"

namespace fcest {

__host__ __device__ inline int pitch4(int x) {  // smallest p >= x with p % 8 == 4
  int p = (x + 7) / 8 * 8 + 4;
  return (p - 8 >= x) ? p - 8 : p;
}

// I_bb[r][k]: inverse of the b-th 8x8 diagonal Cholesky block.  Its strict lower part is stored
// transposed in the strict upper part of diagonal tile (b,b) of T, its diagonal in dinv.
__device__ __forceinline__ double invdiag(const double* T, const double* dinv, int Pd, int b, int r, int k) {
  return k < r ? T[(8 * b + k) * Pd + 8 * b + r] : (k == r ? dinv[8 * b + r] : 0.0);
}

// Cholesky + inverse of one 8x8 diagonal tile, rows held in lanes 0..7 of the calling warp
// (registers + shuffles; no shared-memory round trips inside the 8-step dependency chain).
__device__ __forceinline__ int factor_diag_tile(double* T, double* dinv, int Pd, int b, int lane, int bad) {
  double row[8], x[8], my_rinv = 0.0;
  const int i = lane & 7;
  double* base = T + (8 * b + i) * Pd + 8 * b;
#pragma unroll
  for (int c = 0; c < 8; ++c) { row[c] = (c <= i) ? base[c] : 0.0; x[c] = (c == i) ? 1.0 : 0.0; }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double piv = __shfl_sync(0xffffffffu, row[j], j);
    if (!(piv > 0.0) && bad == 0) bad = 8 * b + j + 1;
    const double rinv = rsqrt(piv);
    if (i == j) my_rinv = rinv;
    row[j] = (i == j) ? piv * rinv : row[j] * rinv;
#pragma unroll
    for (int k = j + 1; k < 8; ++k) {
      const double lkj = __shfl_sync(0xffffffffu, row[j], k);
      if (i >= k) row[k] -= row[j] * lkj;
    }
  }
  // X = L^-1 by forward substitution, row i in lane i: finalise row k, broadcast it, eliminate below
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (i == k) {
#pragma unroll
      for (int c = 0; c <= k; ++c) x[c] *= my_rinv;
    }
#pragma unroll
    for (int c = 0; c <= k; ++c) {
      const double xk = __shfl_sync(0xffffffffu, x[c], k);
      if (i > k) x[c] -= row[k] * xk;
    }
  }
  if (lane < 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c <= i) base[c] = row[c];                                   // L (lower, incl. diagonal)
      if (c < i) T[(8 * b + c) * Pd + 8 * b + i] = x[c];              // I^T (strict upper)
    }
    dinv[8 * b + i] = my_rinv;
  }
  return bad;
}

// Blocked right-looking Cholesky of T (all threads of the CTA must call; contains block barriers).
// Returns the updated `bad` flag (meaningful in warp 0): 1-based index of the first non-positive pivot.
__device__ __forceinline__ int blocked_cholesky(double* T, double* dinv, int Pd, int nbD, int warp, int lane,
                                                int nwarps, int bad) {
  const int gid = lane >> 2, tig = lane & 3;
  if (warp == 0) bad = factor_diag_tile(T, dinv, Pd, 0, lane, bad);
  __syncthreads();
  for (int jb = 0; jb < nbD; ++jb) {
    // panel: L[ib,jb] = S[ib,jb] I_jj^T
    for (int ib = jb + 1 + warp; ib < nbD; ib += nwarps) {
      double* tile = T + (8 * ib + gid) * Pd + 8 * jb;
      double c0 = 0.0, c1 = 0.0;
      const double a0 = tile[tig], a1 = tile[4 + tig];
      dmma(c0, c1, a0, invdiag(T, dinv, Pd, jb, gid, tig));
      dmma(c0, c1, a1, invdiag(T, dinv, Pd, jb, gid, 4 + tig));
      __syncwarp();
      tile[2 * tig] = c0;
      tile[2 * tig + 1] = c1;
    }
    __syncthreads();
    // trailing update S[ib,kb] -= L[ib,jb] L[kb,jb]^T (jb < kb <= ib); warp 0 owns the next
    // diagonal tile and factors it straight away (critical path), warps 1..7 share the rest
    const int m = nbD - 1 - jb;
    if (m > 0) {
    if (warp == 0) {
        const int ib = jb + 1;
        double* ct = T + (8 * ib + gid) * Pd + 8 * ib + 2 * tig;
        double c0 = ct[0], c1 = ct[1];
        const double* la = T + (8 * ib + gid) * Pd + 8 * jb;
        dmma(c0, c1, -la[tig], la[tig]);
        dmma(c0, c1, -la[4 + tig], la[4 + tig]);
        ct[0] = c0;
        ct[1] = c1;
        __syncwarp();
        bad = factor_diag_tile(T, dinv, Pd, ib, lane, bad);
      } else {
        for (int t = warp; t < m * (m + 1) / 2; t += nwarps - 1) {  // t = 0 is warp 0's tile
          int ii = 0, rem = t;
          while (rem > ii) { rem -= ii + 1; ++ii; }
          const int ib = jb + 1 + ii, kb = jb + 1 + rem;
          double* ct = T + (8 * ib + gid) * Pd + 8 * kb + 2 * tig;
          double c0 = ct[0], c1 = ct[1];
          const double* la = T + (8 * ib + gid) * Pd + 8 * jb;
          const double* lb = T + (8 * kb + gid) * Pd + 8 * jb;
          dmma(c0, c1, -la[tig], lb[tig]);
          dmma(c0, c1, -la[4 + tig], lb[4 + tig]);
          ct[0] = c0;
          ct[1] = c1;
        }
      }
    }
    __syncthreads();
  }
  return bad;
}

// z = L^-1 y for the 8-wide column block Yb (Dp x 12, y in column 0), in place; one warp, warp-level
// synchronisation only.  `myscr` is a 96-double scratch tile private to the calling warp.
__device__ __forceinline__ void solve_yblock(const double* T, const double* dinv, int Pd, int nbD, double* Yb,
                                             double* myscr, int lane) {
  const int gid = lane >> 2, tig = lane & 3;
  for (int ib = 0; ib < nbD; ++ib) {
    double* yt = Yb + (8 * ib + gid) * 12 + 2 * tig;
    double c0 = yt[0], c1 = yt[1];
    for (int kb = 0; kb < ib; ++kb) {
      const double* la = T + (8 * ib + gid) * Pd + 8 * kb;
      dmma(c0, c1, -la[tig], Yb[(8 * kb + tig) * 12 + gid]);
      dmma(c0, c1, -la[4 + tig], Yb[(8 * kb + 4 + tig) * 12 + gid]);
    }
    myscr[gid * 12 + 2 * tig] = c0;
    myscr[gid * 12 + 2 * tig + 1] = c1;
    __syncwarp();
    double x0 = 0.0, x1 = 0.0;
    dmma(x0, x1, invdiag(T, dinv, Pd, ib, gid, tig), myscr[tig * 12 + gid]);
    dmma(x0, x1, invdiag(T, dinv, Pd, ib, gid, 4 + tig), myscr[(4 + tig) * 12 + gid]);
    yt[0] = x0;
    yt[1] = x1;
    __syncwarp();
  }
}

}  // namespace fcest
