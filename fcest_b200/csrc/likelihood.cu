// Fused Wishart-process likelihood: forward + analytic backward in one kernel.
//
// Reference semantics (fcest/models/likelihoods.py:126-206, :253-264), per time step n and MC sample s:
//   F = mu_n + sqrt(v_n) * eps_sn  (D x nu),  G = A o F (Hadamard),  Sigma = G G^T + diag(softplus(lam))
//   L = chol(Sigma),  z = L^-1 y_n,  l_sn = -D/2 log(2 pi) - sum log L_dd - 1/2 |z|^2,  ve[n] = mean_s l_sn
// Backward (SURVEY.md App. A "likelihood grad"), w = 1/S, Li = L^-1, alpha = Li^T z:
//   Gbar = w (-Sigma^-1 G + alpha (alpha^T G)),  Fbar = A o Gbar,  Abar = sum F o Gbar,
//   lam_bar_d = sigmoid(lam_d) sum w (-1/2 (Sigma^-1)_dd + 1/2 alpha_d^2),
//   mu_bar = sum_s Fbar,  v_bar = sum_s Fbar o eps / (2 sqrt v).
//
// Mapping: one CTA per time step n (all S samples, so the sums over s stay on chip), everything for
// one D x D problem lives in shared memory (~68 KB at D = 50 -> 3 CTAs per SM).  All O(D^3) work runs
// on the fp64 tensor pipe (DMMA.8x8x4) from conflict-free padded tiles (pitch = 4 mod 8 doubles):
// G G^T, a blocked (8x8) right-looking Cholesky whose diagonal tiles are factored + inverted in
// registers by one warp (shuffles), the block-column recurrences for L^-1 and L^-1 y (one warp per
// column block, warp-level sync only), and Sigma^-1 G = Li^T (Li G) done in place per column block.
// Per-CTA partial sums for Abar / lam_bar go to a workspace and are reduced in fixed order
// (deterministic) by colsum_kernel.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"
#include "lik_params.cuh"

namespace fcest {



// One CTA (8 warps) per time step; per MC sample a blocked (8x8 tiles) Cholesky / inverse on the fp64
// tensor pipe.  Shared memory: G (D x nu, later H = Li G, then Q = Sigma^-1 G, in place per column
// block), T (lower: Sigma -> L; strict upper: Li^T), a y column block and per-warp scratch tiles.
// BIG = false: the whole problem lives in shared memory (D <= ~80), one CTA per time step.
// BIG = true : G, T and the y block live in a per-CTA slot of an L2-resident global workspace (same
//              tile algorithm, block / warp barriers order the global accesses inside the CTA); a
//              persistent grid walks the time steps.  Functional path for D up to 400; a cluster /
//              TMA-staged variant is the planned replacement.
template <bool BIG>
__global__ void __launch_bounds__(256, 3)
wishart_loglik_kernel(const LikParams p) {
  extern __shared__ __align__(16) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu;
  const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8;
  const int Pd = pitch4(Dp), Pg = pitch4(Np);
  const int nbD = Dp / 8, nbN = Np / 8;
  double* big = BIG ? p.ws + (long long)blockIdx.x * p.ws_stride : sm;
  double* G = big;                // Dp x Pg
  double* T = G + Dp * Pg;        // Dp x Pd
  double* Yb = T + Dp * Pd;       // Dp x 12 : column 0 = y -> z -> alpha
  double* scr = BIG ? sm : Yb + Dp * 12;  // 8 warps x 96
  double* dinv = scr + 8 * 96;    // Dp  1 / L_dd
  double* yv = dinv + Dp;         // Dp
  double* spl = yv + Dp;          // Dp  softplus(lam)
  double* ds = spl + Dp;          // Dp  diag(Sigma^-1)
  double* cv = ds + Dp;           // Np  c = alpha^T G = y^T Q
  double* scal = cv + Np;         // [0] = |z|^2
  __shared__ int bad_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = 8;
  const int gid = lane >> 2, tig = lane & 3;
  const double w = 1.0 / p.S;
  double* myscr = scr + warp * 96;

  for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
  const double* fm = p.fmean + (long long)n * p.ldf;
  const double* fv = p.fvar + (long long)n * p.ldf;
  __syncthreads();
  for (int d = tid; d < Dp; d += blockDim.x) {
    yv[d] = d < D ? p.y[(long long)n * D + d] : 0.0;
    spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  }
  if (tid == 0) bad_s = 0;
  double ve_acc = 0.0;
  int bad = 0;

  for (int s = 0; s < p.S; ++s) {
    const double* ep = p.eps + ((long long)s * p.N + n) * R;
    __syncthreads();
    // P1. G = A o (mu + sqrt(v) eps) with zero padding; y block
    for (int e = tid; e < Dp * Pg; e += blockDim.x) {
      const int d = e / Pg, k = e - d * Pg;
      double v = 0.0;
      if (d < D && k < nu) {
        const int r = d * nu + k;
        v = (p.a_mode == 0 ? p.A[r] : p.A[k]) * (fm[r] + sqrt(fv[r]) * ep[r]);
      }
      G[e] = v;
    }
    for (int e = tid; e < Dp * 12; e += blockDim.x) Yb[e] = (e % 12 == 0) ? yv[e / 12] : 0.0;
    __syncthreads();
    // P2. Sigma (lower tiles) = G G^T + diag(softplus(lam)); padded diagonal = 1
    for (int t = warp; t < nbD * (nbD + 1) / 2; t += nwarps) {
      int ib = 0, rem = t;
      while (rem > ib) { rem -= ib + 1; ++ib; }
      const int jb = rem;
      double c0 = 0.0, c1 = 0.0;
      const double* ga = G + (8 * ib + gid) * Pg + tig;
      const double* gb = G + (8 * jb + gid) * Pg + tig;
      for (int k0 = 0; k0 < Np; k0 += 4) dmma(c0, c1, ga[k0], gb[k0]);
      const int row = 8 * ib + gid, col = 8 * jb + 2 * tig;
      if (row == col) c0 += spl[row];
      if (row == col + 1) c1 += spl[row];
      T[row * Pd + col] = c0;
      T[row * Pd + col + 1] = c1;
    }
    __syncthreads();
    // P3. blocked right-looking Cholesky (chol_tiles.cuh); diagonal tiles factored + inverted by warp 0
    bad = blocked_cholesky(T, dinv, Pd, nbD, warp, lane, nwarps, bad);
    // P4. columns of Li = L^-1 (one warp per column block, no block barriers) and z = L^-1 y
    for (int task = warp; task <= nbD; task += nwarps) {
      if (task < nbD) {
        if (!p.need_grad) continue;
        const int jb = task;
        for (int ib = jb + 1; ib < nbD; ++ib) {
          double c0 = 0.0, c1 = 0.0;
          for (int kb = jb; kb < ib; ++kb) {
            const double* la = T + (8 * ib + gid) * Pd + 8 * kb;
            double b0, b1;
            if (kb == jb) {
              b0 = invdiag(T, dinv, Pd, jb, tig, gid);
              b1 = invdiag(T, dinv, Pd, jb, 4 + tig, gid);
            } else {
              const double* xb = T + (8 * jb + gid) * Pd + 8 * kb;
              b0 = xb[tig];
              b1 = xb[4 + tig];
            }
            dmma(c0, c1, la[tig], b0);
            dmma(c0, c1, la[4 + tig], b1);
          }
          myscr[gid * 12 + 2 * tig] = c0;
          myscr[gid * 12 + 2 * tig + 1] = c1;
          __syncwarp();
          double x0 = 0.0, x1 = 0.0;
          dmma(x0, x1, invdiag(T, dinv, Pd, ib, gid, tig), myscr[tig * 12 + gid]);
          dmma(x0, x1, invdiag(T, dinv, Pd, ib, gid, 4 + tig), myscr[(4 + tig) * 12 + gid]);
          T[(8 * jb + 2 * tig) * Pd + 8 * ib + gid] = -x0;      // Li[ib,jb]^T into the upper tile (jb,ib)
          T[(8 * jb + 2 * tig + 1) * Pd + 8 * ib + gid] = -x1;
          __syncwarp();
        }
      } else {
        solve_yblock(T, dinv, Pd, nbD, Yb, myscr, lane);
        double q = 0.0;
        for (int d = lane; d < D; d += 32) q += Yb[d * 12] * Yb[d * 12];
        q = warp_sum(q);
        if (lane == 0) scal[0] = q;
      }
    }
    __syncthreads();
    if (warp == 0) {
      double ld = 0.0;
      for (int d = lane; d < D; d += 32) ld -= log(dinv[d]);
      ld = warp_sum(ld);
      ve_acc += w * (-0.5 * D * 1.8378770664093453 - ld - 0.5 * scal[0]);
    }
    if (!p.need_grad) continue;
    // P5. per column block (one warp each, in place): H = Li G, then Q = Li^T H = Sigma^-1 G;
    //     the y block gives alpha = Li^T z
    for (int task = warp; task <= nbN; task += nwarps) {
      const bool ycol = (task == nbN);
      double* B = ycol ? Yb : G + 8 * task;
      const int Pb = ycol ? 12 : Pg;
      if (!ycol) {
        // H = Li B, bottom-up: tile ib only needs input tiles kb <= ib, so it can be overwritten at once
        for (int ib = nbD - 1; ib >= 0; --ib) {
          double c0 = 0.0, c1 = 0.0;
          for (int kb = 0; kb <= ib; ++kb) {
            double a0, a1;
            if (kb == ib) {
              a0 = invdiag(T, dinv, Pd, ib, gid, tig);
              a1 = invdiag(T, dinv, Pd, ib, gid, 4 + tig);
            } else {
              a0 = T[(8 * kb + tig) * Pd + 8 * ib + gid];
              a1 = T[(8 * kb + 4 + tig) * Pd + 8 * ib + gid];
            }
            dmma(c0, c1, a0, B[(8 * kb + tig) * Pb + gid]);
            dmma(c0, c1, a1, B[(8 * kb + 4 + tig) * Pb + gid]);
          }
          __syncwarp();
          B[(8 * ib + gid) * Pb + 2 * tig] = c0;
          B[(8 * ib + gid) * Pb + 2 * tig + 1] = c1;
        }
        __syncwarp();
      }
      // Q = Li^T H, top-down: tile mb only needs input tiles kb >= mb
      for (int mb = 0; mb < nbD; ++mb) {
        double c0 = 0.0, c1 = 0.0;
        for (int kb = mb; kb < nbD; ++kb) {
          double a0, a1;
          if (kb == mb) {
            a0 = invdiag(T, dinv, Pd, mb, tig, gid);
            a1 = invdiag(T, dinv, Pd, mb, 4 + tig, gid);
          } else {
            a0 = T[(8 * mb + gid) * Pd + 8 * kb + tig];
            a1 = T[(8 * mb + gid) * Pd + 8 * kb + 4 + tig];
          }
          dmma(c0, c1, a0, B[(8 * kb + tig) * Pb + gid]);
          dmma(c0, c1, a1, B[(8 * kb + 4 + tig) * Pb + gid]);
        }
        __syncwarp();
        B[(8 * mb + gid) * Pb + 2 * tig] = c0;
        B[(8 * mb + gid) * Pb + 2 * tig + 1] = c1;
      }
    }
    __syncthreads();
    // P6. c = y^T Q (= alpha^T G), diag(Sigma^-1)_d = sum_k Li[k,d]^2
    for (int k = tid; k < Np; k += blockDim.x) {
      double c = 0.0;
      if (k < nu)
        for (int d = 0; d < D; ++d) c += yv[d] * G[d * Pg + k];
      cv[k] = c;
    }
    for (int d = tid; d < Dp; d += blockDim.x) {
      double sq = dinv[d] * dinv[d];
      const int b = d >> 3;
      for (int k = d + 1; k < 8 * (b + 1); ++k) { const double l = T[d * Pd + k]; sq += l * l; }
      for (int k = 8 * (b + 1); k < D; ++k) { const double l = T[d * Pd + k]; sq += l * l; }
      ds[d] = sq;
    }
    __syncthreads();
    // P7. gradients
    for (int r = tid; r < R; r += blockDim.x) {
      const int d = r / nu, k = r - d * nu;
      const double gbar = w * (-G[d * Pg + k] + Yb[d * 12] * cv[k]);
      const double a = (p.a_mode == 0 ? p.A[r] : p.A[k]);
      const double sd = sqrt(fv[r]), e = ep[r];
      const double f = fm[r] + sd * e;
      const double fbar = a * gbar;
      const long long o = (long long)n * p.ldg + r;
      const long long oa = (long long)n * R + r;
      if (s == 0) {
        p.g_fmean[o] = fbar;
        p.g_fvar[o] = fbar * e / (2.0 * sd);
        p.gA_part[oa] = f * gbar;
      } else {
        p.g_fmean[o] += fbar;
        p.g_fvar[o] += fbar * e / (2.0 * sd);
        p.gA_part[oa] += f * gbar;
      }
    }
    for (int d = tid; d < D; d += blockDim.x) {
      const double al = Yb[d * 12];
      const double v = sigmoid_d(p.lam[d]) * w * (-0.5 * ds[d] + 0.5 * al * al);
      const long long o = (long long)n * D + d;
      if (s == 0) p.glam_part[o] = v; else p.glam_part[o] += v;
    }
  }
  if (warp == 0 && lane == 0 && bad) atomicMax(&bad_s, bad);
  __syncthreads();
  if (tid == 0) {
    p.ve[n] = bad_s ? nan("") : ve_acc;
    p.info[n] = bad_s;
  }
  }  // time-step loop
}

// out[c] = sum_{row} X[row, c] in a fixed (deterministic) order; block = (32 columns) x (32 row lanes).
// fold_len > 0 (vector-A mode) folds the cols/fold_len column groups onto fold_len outputs.
__global__ void colsum_kernel(const double* __restrict__ X, long long ld, int rows, int cols,
                              double* __restrict__ out, int fold_len) {
  __shared__ double part[32][33];
  const int nout = fold_len ? fold_len : cols;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int groups = fold_len ? cols / fold_len : 1;
  double s = 0.0;
  if (c < nout)
    for (int g = 0; g < groups; ++g)
      for (int r = threadIdx.y; r < rows; r += 32) s += X[(long long)r * ld + g * nout + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < nout) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += part[k][threadIdx.x];
    out[c] = t;
  }
}

}  // namespace fcest

using namespace fcest;

extern "C" long long fcest_wishart_loglik_smem_bytes(int D, int nu) {
  const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8;
  const int Pd = pitch4(Dp), Pg = pitch4(Np);
  return (long long)sizeof(double) * ((long long)Dp * Pg + (long long)Dp * Pd + 12LL * Dp + 8 * 96 + 4 * Dp + Np + 8);
}

// Large-D variant: number of persistent CTAs and bytes of global scratch the caller must provide
// (0 when the problem fits shared memory and no scratch is needed).
static const int kBigCtas = 148 * 3;
extern "C" long long fcest_wishart_loglik_scratch_bytes(int D, int nu) {
  long long need = 0;
  if (!loglik_warp_supported(D, nu) && loglik_tiles_supported(D, nu)) need = loglik_tiles_scratch_bytes(D, nu);
  // the cluster-tiled kernel serves the sizes neither of the above takes (and any size under FCEST_LIK_IMPL=cluster)
  static const bool force_cluster_scr = [] { const char* v = getenv("FCEST_LIK_IMPL"); return v && !strcmp(v, "cluster"); }();
  if ((force_cluster_scr || (!loglik_warp_supported(D, nu) && !loglik_tiles_supported(D, nu))) && loglik_cluster_supported(D, nu) &&
      loglik_cluster_scratch_bytes(D, nu) > need)
    need = loglik_cluster_scratch_bytes(D, nu);
  if (fcest_wishart_loglik_smem_bytes(D, nu) > 227 * 1024) {
    const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8;
    const long long stride = (long long)Dp * pitch4(Np) + (long long)Dp * pitch4(Dp) + 12LL * Dp;
    const long long big = (long long)sizeof(double) * stride * kBigCtas;
    if (big > need) need = big;
  }
  return need;
}

extern "C" int fcest_wishart_loglik(const double* fmean, const double* fvar, long long ldf,
                                    const double* eps, const double* y, const double* A, int a_mode,
                                    const double* lam, int S, int N, int D, int nu, double* ve,
                                    double* g_fmean, double* g_fvar, long long ldg, double* gA,
                                    double* glam, double* ws_gA_part, double* ws_glam_part,
                                    int* info, double* scratch, cudaStream_t stream) {
  if (N <= 0) return 0;
  long long smem = fcest_wishart_loglik_smem_bytes(D, nu);
  const bool bigD = smem > 227 * 1024;
  if (D > 400 || nu > 400) return FCEST_ERR_UNSUPPORTED;
  if (bigD && scratch == nullptr) return FCEST_ERR_ARGUMENT;
  LikParams p;
  p.fmean = fmean; p.fvar = fvar; p.ldf = ldf; p.eps = eps; p.y = y; p.A = A; p.a_mode = a_mode;
  p.lam = lam; p.S = S; p.N = N; p.D = D; p.nu = nu; p.ve = ve; p.g_fmean = g_fmean;
  p.g_fvar = g_fvar; p.ldg = ldg; p.gA_part = ws_gA_part; p.glam_part = ws_glam_part; p.info = info;
  p.need_grad = (g_fmean != nullptr);
  if (p.need_grad && (!g_fvar || !ws_gA_part || !ws_glam_part || !gA || !glam)) return FCEST_ERR_ARGUMENT;
  p.ws = scratch;
  p.ws_stride = 0;
  // D <= 55: warp-per-matrix kernel (likelihood_warp.cu); FCEST_LIK_IMPL=cta forces the older CTA-per-time-step
  // kernels of this file (shared memory for D <= ~80, L2 scratch above)
  static const bool force_cta = [] { const char* v = getenv("FCEST_LIK_IMPL"); return v && !strcmp(v, "cta"); }();
  static const bool force_warp = [] { const char* v = getenv("FCEST_LIK_IMPL"); return v && !strcmp(v, "warp"); }();
  static const bool force_cluster = [] { const char* v = getenv("FCEST_LIK_IMPL"); return v && !strcmp(v, "cluster"); }();
  bool launched = false;
  int part_rows = N;  // rows of the A_bar / lam_bar partial workspaces that the kernel filled
  // D <= 7: sub-warp kernel (likelihood_subwarp.cu): four problems per warp, registers and shuffles only
  // FCEST_LIK_IMPL=cluster: the cluster-tiled kernel at any size it supports (A/B runs against the tiles kernel)
  if (force_cluster && loglik_cluster_supported(D, nu) && scratch != nullptr) {
    const int rc = launch_loglik_cluster(p, stream, &part_rows);
    if (rc > 0) return rc;
    launched = (rc == 0);
  }
  if (!launched && !force_cta && !force_warp && loglik_subwarp_supported(D, nu)) {
    const int rc = launch_loglik_subwarp(p, stream, &part_rows);
    if (rc > 0) return rc;
    launched = (rc == 0);
  }
  if (!launched && !force_cta && loglik_warp_supported(D, nu)) {
    const int rc = launch_loglik_warp(p, stream, &part_rows);
    if (rc > 0) return rc;
    launched = (rc == 0);
  }
  // 56 <= D <= 207: shared-memory-tiled CTA-per-matrix kernel (likelihood_tiles.cu)
  if (!launched && !force_cta && loglik_tiles_supported(D, nu) && scratch != nullptr) {
    const int rc = launch_loglik_tiles(p, stream, &part_rows);
    if (rc > 0) return rc;
    launched = (rc == 0);
  }
  // 208 <= D <= 407: cluster-tiled kernel (likelihood_cluster.cu): the factor's tiles in the distributed shared memory
  // of a thread-block cluster
  if (!launched && !force_cta && loglik_cluster_supported(D, nu) && scratch != nullptr) {
    const int rc = launch_loglik_cluster(p, stream, &part_rows);
    if (rc > 0) return rc;
    launched = (rc == 0);
  }
  if (launched) {
  } else if (!bigD) {
    cudaError_t e = cudaFuncSetAttribute(wishart_loglik_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    wishart_loglik_kernel<false><<<N, 256, smem, stream>>>(p);
  } else {
    const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8;
    p.ws_stride = (long long)Dp * pitch4(Np) + (long long)Dp * pitch4(Dp) + 12LL * Dp;
    smem = sizeof(double) * (8 * 96 + 4 * Dp + Np + 8);
    wishart_loglik_kernel<true><<<N < kBigCtas ? N : kBigCtas, 256, smem, stream>>>(p);
  }
  FCEST_CHECK_LAUNCH();
  if (p.need_grad) {
    const int R = D * nu;
    const dim3 cb(32, 32);
    if (a_mode == 0) {
      colsum_kernel<<<(R + 31) / 32, cb, 0, stream>>>(ws_gA_part, R, part_rows, R, gA, 0);
    } else {
      colsum_kernel<<<(nu + 31) / 32, cb, 0, stream>>>(ws_gA_part, R, part_rows, R, gA, nu);
    }
    FCEST_CHECK_LAUNCH();
    colsum_kernel<<<(D + 31) / 32, cb, 0, stream>>>(ws_glam_part, D, part_rows, D, glam, 0);
    FCEST_CHECK_LAUNCH();
  }
  return 0;
}
