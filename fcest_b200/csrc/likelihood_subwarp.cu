// Sub-warp Wishart likelihood (forward + analytic backward) for D <= 7: one 8-lane group per time step, four time
// steps per warp, everything in registers, shuffle-based factorisation (north_star: "one warp or sub-warp per matrix,
// held in registers with shuffle-based factorisation").
//
// Reference semantics: fcest/models/likelihoods.py:126-206, :253-264 (same math as likelihood_warp.cu).
//
// A D x D problem with D <= 7 and its bordered row (row D = y^T, so that z = L^-1 y falls out of the factorisation)
// is at most 8 x 8: lane i of the group owns row i of G (nu <= 8 values), of Sigma / L and of every right-hand side.
// The warp-per-matrix kernel pads such a problem to one 8 x 8 DMMA tile on 32 lanes; here four problems share a warp
// and no tensor-core instruction is involved.  A CTA holds 16 groups: Gs = min(S, 16) groups share a time step (group gs
// takes the samples s = gs, gs + Gs, ...), 16 / Gs time steps per CTA; the per-group sums over its samples stay in
// registers and are combined across the groups of a time step through shared memory in group order: every output
// has one owner and a fixed summation order (deterministic, no atomics).
#include <math.h>

#include "common.cuh"
#include "fcest_b200.h"
#include "lik_params.cuh"

namespace fcest {

__device__ __forceinline__ double gshfl(unsigned mask, double v, int src) { return __shfl_sync(mask, v, src, 8); }
__device__ __forceinline__ double gsum(unsigned mask, double v) {
  v += __shfl_xor_sync(mask, v, 1, 8);
  v += __shfl_xor_sync(mask, v, 2, 8);
  v += __shfl_xor_sync(mask, v, 4, 8);
  return v;
}

__global__ void __launch_bounds__(128) wishart_loglik_subwarp_kernel(const LikParams p) {
  __shared__ double part[16][8][27];        // per group and lane: mu_bar, v_bar, A_bar (8 each), lam_bar, ve, bad
  const int D = p.D, nu = p.nu, R = D * nu, N = p.N;
  const int Gs = p.S < 16 ? p.S : 16;       // groups per time step
  const int T = 16 / Gs;                    // time steps per CTA
  const int grp = threadIdx.x >> 3;
  const int tl = grp / Gs, gs = grp - tl * Gs;
  const int n = blockIdx.x * T + tl;
  const bool active = tl < T && n < N;      // whole groups are active or not: the shuffle masks name one group only
  const int i = threadIdx.x & 7;            // row owned by this lane
  const unsigned mask = 0xFFu << (8 * ((threadIdx.x & 31) >> 3));
  const double wS = 1.0 / p.S;
  const bool row = i < D;
  if (active) {

  // sample-independent pieces of row i: A, fmean, sqrt(fvar), y, softplus(lam)
  double Ar[8], fm[8], sd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool ok = row && k < nu;
    const int r = ok ? i * nu + k : 0;
    Ar[k] = ok ? (p.a_mode == 0 ? p.A[r] : p.A[k]) : 0.0;
    fm[k] = ok ? p.fmean[(long long)n * p.ldf + r] : 0.0;
    sd[k] = ok ? sqrt(p.fvar[(long long)n * p.ldf + r]) : 1.0;
  }
  const double yi = row ? p.y[(long long)n * D + i] : 0.0;
  const double lam_i = row ? p.lam[i] : 0.0;
  const double spl = row ? softplus_d(lam_i) : 1.0;
  double yrow[8];                           // row D of the bordered matrix: y^T, then the unit pivot
#pragma unroll
  for (int j = 0; j < 8; ++j) yrow[j] = gshfl(mask, yi, j);

  double mu_bar[8], v_bar[8], A_bar[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) mu_bar[k] = v_bar[k] = A_bar[k] = 0.0;
  double lam_bar = 0.0, ve_acc = 0.0;
  int bad = 0;

  for (int s = gs; s < p.S; s += Gs) {
    const double* ep = p.eps + ((long long)s * N + n) * R;
    double e[8], gk[8], a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      e[k] = (row && k < nu) ? __ldg(ep + i * nu + k) : 0.0;
      gk[k] = Ar[k] * (fm[k] + sd[k] * e[k]);                     // G = A o F  (likelihoods.py:177); zero off the matrix
    }
    // Sigma row i = G_i . G_j + softplus(lam) on the diagonal (:179-180); rows >= D: bordered y row / identity
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += gk[k] * gshfl(mask, gk[k], j);
      if (i < D) a[j] = acc + (j == i ? spl : 0.0);
      else if (i == D) a[j] = j < D ? yrow[j] : (j == D ? 1.0 : 0.0);
      else a[j] = j == i ? 1.0 : 0.0;
    }
    // right-looking Cholesky over the D real columns; the bordered row rides along and ends up holding z = L^-1 y
    double diag = 1.0;                      // l_ii of this lane's row
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j < D) {
        const double piv = gshfl(mask, a[j], j);
        if (!(piv > 0.0) && !bad) bad = j + 1;
        const double ljj = sqrt(piv), inv = 1.0 / ljj;
        if (i == j) { a[j] = ljj; diag = ljj; }
        else if (i > j) a[j] *= inv;
#pragma unroll
        for (int k = j + 1; k < 8; ++k) {
          const double lkj = gshfl(mask, a[j], k);
          if (i >= k) a[k] -= a[j] * lkj;
        }
      }
    }
    // log density of this sample (:172, 190-204): -D/2 log 2 pi - sum log l_ii - |z|^2 / 2
    const double ld = gsum(mask, row ? log(diag) : 0.0);
    double zz = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) zz += (i == D && j < D) ? a[j] * a[j] : 0.0;
    zz = gsum(mask, zz);
    ve_acc += wS * (-0.5 * D * 1.8378770664093453 - ld - 0.5 * zz);

    if (p.need_grad) {
      // column i of L (and z_i in slot D): ct[k] = L[k][i]
      double ct[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double mine = 0.0;
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
          const double v = gshfl(mask, a[idx], k);
          if (idx == i) mine = v;
        }
        ct[k] = mine;
      }
      const double inv_d = 1.0 / diag;
      // alpha = L^-T z (back substitution, lane k finalises alpha_k)
      double zi = 0.0, alpha = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k == D) zi = ct[k];
#pragma unroll
      for (int k = 6; k >= 0; --k) {
        if (k < D) {
          const double ak = gshfl(mask, zi * inv_d, k);
          if (i == k) alpha = ak;
          else if (i < k) zi -= ct[k] * ak;
        }
      }
      // H = L^-1 G (forward substitution on the nu columns), then Q = L^-T H (back substitution)
      double h[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) h[c] = gk[c];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        if (k < D) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double hk = gshfl(mask, h[c] * inv_d, k);
            if (i == k) h[c] = hk;
            else if (i > k && row) h[c] -= a[k] * hk;
          }
        }
      }
#pragma unroll
      for (int k = 6; k >= 0; --k) {
        if (k < D) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double qk = gshfl(mask, h[c] * inv_d, k);
            if (i == k) h[c] = qk;
            else if (i < k) h[c] -= ct[k] * qk;
          }
        }
      }
      // Gbar = w (-Sigma^-1 G + alpha (alpha^T G));  cotangents of fmean, fvar, A; lam_bar through
      // M_dd Lam_d = -1/2 (1 - alpha_d y_d - sum_c Q''_dc G_dc),  Q'' = Sigma^-1 G - alpha (alpha^T G)
      double rs = 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double t = gsum(mask, row ? alpha * gk[c] : 0.0);
        const double qpp = h[c] - alpha * t;
        const double gbar = -wS * qpp;
        if (row && c < nu) {
          rs += qpp * gk[c];
          mu_bar[c] += Ar[c] * gbar;
          v_bar[c] += Ar[c] * gbar * e[c] / (2.0 * sd[c]);
          A_bar[c] += (fm[c] + sd[c] * e[c]) * gbar;
        }
      }
      if (row) lam_bar += sigmoid_d(lam_i) * wS * (-0.5) * (1.0 - alpha * yi - rs) / spl;
    }
  }

  // this group's partial sums -> shared memory
#pragma unroll
  for (int c = 0; c < 8; ++c) { part[grp][i][c] = mu_bar[c]; part[grp][i][8 + c] = v_bar[c]; part[grp][i][16 + c] = A_bar[c]; }
  part[grp][i][24] = lam_bar;
  part[grp][i][25] = ve_acc;
  part[grp][i][26] = (double)bad;
  }  // active
  __syncthreads();
  if (!active || gs != 0) return;
  // group 0 of the time step: sum over the groups in order, write the outputs
  double mu_bar[8], v_bar[8], A_bar[8], lam_bar = 0.0, ve_acc = 0.0;
  int bad = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) mu_bar[c] = v_bar[c] = A_bar[c] = 0.0;
  for (int g2 = 0; g2 < Gs; ++g2) {
    const double* q = part[grp + g2][i];
#pragma unroll
    for (int c = 0; c < 8; ++c) { mu_bar[c] += q[c]; v_bar[c] += q[8 + c]; A_bar[c] += q[16 + c]; }
    lam_bar += q[24];
    ve_acc += q[25];
    const int b = (int)q[26];
    if (b && !bad) bad = b;
  }
  if (i == 0) {   // every lane of a group saw the same pivots
    p.ve[n] = bad ? nan("") : ve_acc;
    p.info[n] = bad;
  }
  if (p.need_grad && row) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < nu) {
        const int r = i * nu + c;
        p.g_fmean[(long long)n * p.ldg + r] = mu_bar[c];
        p.g_fvar[(long long)n * p.ldg + r] = v_bar[c];
        p.gA_part[(long long)n * R + r] = A_bar[c];
      }
    }
    p.glam_part[(long long)n * D + i] = lam_bar;
  }
}

bool loglik_subwarp_supported(int D, int nu) { return D >= 1 && D <= 7 && nu >= 1 && nu <= 8; }

int launch_loglik_subwarp(const LikParams& p, cudaStream_t stream, int* part_rows) {
  if (!loglik_subwarp_supported(p.D, p.nu)) return -1;
  const int Gs = p.S < 16 ? p.S : 16, T = 16 / Gs;
  wishart_loglik_subwarp_kernel<<<(p.N + T - 1) / T, 128, 0, stream>>>(p);
  *part_rows = p.N;
  return 0;
}

}  // namespace fcest
