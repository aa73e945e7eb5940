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
// one D x D problem lives in shared memory, the three matmul-shaped pieces (G G^T, Li G, Li^T H) run
// on the fp64 tensor pipe (DMMA.8x8x4) from conflict-free padded tiles (pitch = 4 mod 8 doubles).
// Per-CTA partial sums for Abar / lam_bar go to a workspace and are reduced in fixed order
// (deterministic) by colsum_kernel.
#include <math.h>

#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

__host__ __device__ inline int pitch4(int x) {  // smallest p >= x with p % 8 == 4
  int p = (x + 7) / 8 * 8 + 4;
  return (p - 8 >= x) ? p - 8 : p;
}

struct LikParams {
  const double* fmean; const double* fvar; long long ldf;
  const double* eps;  // (S, N, R) contiguous
  const double* y;    // (N, D) contiguous
  const double* A; int a_mode;  // 0: (D, nu) matrix, 1: (nu,) vector broadcast over rows
  const double* lam;  // (D,) unconstrained
  int S, N, D, nu;
  double* ve;  // (N,)
  double* g_fmean; double* g_fvar; long long ldg;  // (N, R) or null
  double* gA_part;    // (N, D*nu) per-time-step partials
  double* glam_part;  // (N, D)
  int* info;          // (N,) 0 = ok, else 1-based index of the first non-positive pivot (any sample)
  int need_grad;
};

__global__ void __launch_bounds__(256)
wishart_loglik_kernel(const LikParams p) {
  extern __shared__ __align__(16) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu;
  const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8;
  const int Pd = pitch4(Dp), Pg = pitch4(Np);
  double* G = sm;                 // Dp x Pg   (later holds Q = Sigma^-1 G)
  double* H = G + Dp * Pg;        // Dp x Pg
  double* Sg = H + Dp * Pg;       // Dp x Pd   Sigma -> L (lower)
  double* Li = Sg + Dp * Pd;      // Dp x Pd   L^-1 (lower, upper zero)
  double* yv = Li + Dp * Pd;      // Dp
  double* zv = yv + Dp;           // Dp
  double* al = zv + Dp;           // Dp  alpha
  double* ds = al + Dp;           // Dp  diag(Sigma^-1)
  double* spl = ds + Dp;          // Dp  softplus(lam)
  double* cv = spl + Dp;          // Np  c = alpha^T G
  double* red = cv + Np;          // 32

  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nbD = Dp / 8, nbN = Np / 8;
  const double w = 1.0 / p.S;
  const double* fm = p.fmean + (long long)n * p.ldf;
  const double* fv = p.fvar + (long long)n * p.ldf;

  for (int d = tid; d < Dp; d += blockDim.x) {
    yv[d] = d < D ? p.y[(long long)n * D + d] : 0.0;
    spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  }
  double ve_acc = 0.0;
  int bad = 0;

  for (int s = 0; s < p.S; ++s) {
    const double* ep = p.eps + ((long long)s * p.N + n) * R;
    __syncthreads();
    // 1. G = A o (mu + sqrt(v) eps), zero padding; Li = 0
    for (int e = tid; e < Dp * Pg; e += blockDim.x) {
      const int d = e / Pg, k = e - d * Pg;
      double v = 0.0;
      if (d < D && k < nu) {
        const int r = d * nu + k;
        const double f = fm[r] + sqrt(fv[r]) * ep[r];
        v = (p.a_mode == 0 ? p.A[r] : p.A[k]) * f;
      }
      G[e] = v;
    }
    for (int e = tid; e < Dp * Pd; e += blockDim.x) Li[e] = 0.0;
    __syncthreads();
    // 2. Sigma (lower tiles) = G G^T + diag(softplus(lam)); padded diagonal = 1
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
      Sg[row * Pd + col] = c0;
      Sg[row * Pd + col + 1] = c1;
    }
    __syncthreads();
    // 3. in-place right-looking Cholesky of the leading D x D block (lower)
    for (int j = 0; j < D; ++j) {
      const double piv = Sg[j * Pd + j];
      if (!(piv > 0.0)) { if (!bad) bad = j + 1; break; }  // uniform: every thread reads the same piv
      const double dj = sqrt(piv), inv = 1.0 / dj;
      for (int i = j + 1 + tid; i < D; i += blockDim.x) Sg[i * Pd + j] *= inv;
      __syncthreads();
      if (tid == 0) Sg[j * Pd + j] = dj;
      for (int i = j + 1 + warp; i < D; i += nwarps) {
        const double lij = Sg[i * Pd + j];
        for (int k = j + 1 + lane; k <= i; k += 32) Sg[i * Pd + k] -= lij * Sg[k * Pd + j];
      }
      __syncthreads();
    }
    if (bad) break;
    // 4. forward substitution with RHS [I | y]: column c < D -> Li[:, c], column D -> z
    if (tid <= D) {
      const int c = tid;
      if (c < D) {
        for (int i = c; i < D; ++i) {
          double sacc = (i == c) ? 1.0 : 0.0;
          const double* Lrow = Sg + i * Pd;
          for (int k = c; k < i; ++k) sacc -= Lrow[k] * Li[k * Pd + c];
          Li[i * Pd + c] = sacc / Lrow[i];
        }
      } else {
        for (int i = 0; i < D; ++i) {
          double sacc = yv[i];
          const double* Lrow = Sg + i * Pd;
          for (int k = 0; k < i; ++k) sacc -= Lrow[k] * zv[k];
          zv[i] = sacc / Lrow[i];
        }
      }
    }
    __syncthreads();
    // 5. log-density; alpha = Li^T z, diag(Sigma^-1)
    if (warp == 0) {
      double ld = 0.0, q = 0.0;
      for (int d = lane; d < D; d += 32) { ld += log(Sg[d * Pd + d]); q += zv[d] * zv[d]; }
      ld = warp_sum(ld);
      q = warp_sum(q);
      ve_acc += w * (-0.5 * D * 1.8378770664093453 - ld - 0.5 * q);
    }
    if (!p.need_grad) continue;
    for (int d = tid; d < Dp; d += blockDim.x) {
      double a = 0.0, sq = 0.0;
      if (d < D)
        for (int k = d; k < D; ++k) { const double l = Li[k * Pd + d]; a += l * zv[k]; sq += l * l; }
      al[d] = a;
      ds[d] = sq;
    }
    __syncthreads();
    for (int k = tid; k < Np; k += blockDim.x) {
      double c = 0.0;
      if (k < nu)
        for (int d = 0; d < D; ++d) c += al[d] * G[d * Pg + k];
      cv[k] = c;
    }
    // 6. H = Li G   (lower-triangular A operand: k-loop stops at the diagonal block)
    for (int t = warp; t < nbD * nbN; t += nwarps) {
      const int ib = t / nbN, cb = t - ib * nbN;
      double c0 = 0.0, c1 = 0.0;
      const double* la = Li + (8 * ib + gid) * Pd + tig;
      const double* gb = G + tig * Pg + 8 * cb + gid;
      for (int k0 = 0; k0 < 8 * (ib + 1); k0 += 4) dmma(c0, c1, la[k0], gb[k0 * Pg]);
      const int row = 8 * ib + gid, col = 8 * cb + 2 * tig;
      H[row * Pg + col] = c0;
      H[row * Pg + col + 1] = c1;
    }
    __syncthreads();
    // 7. Q = Li^T H = Sigma^-1 G  (written over G)
    for (int t = warp; t < nbD * nbN; t += nwarps) {
      const int mb = t / nbN, cb = t - mb * nbN;
      double c0 = 0.0, c1 = 0.0;
      const double* la = Li + tig * Pd + 8 * mb + gid;
      const double* hb = H + tig * Pg + 8 * cb + gid;
      for (int k0 = 8 * mb; k0 < Dp; k0 += 4) dmma(c0, c1, la[k0 * Pd], hb[k0 * Pg]);
      const int row = 8 * mb + gid, col = 8 * cb + 2 * tig;
      G[row * Pg + col] = c0;
      G[row * Pg + col + 1] = c1;
    }
    __syncthreads();
    // 8. gradients
    for (int r = tid; r < R; r += blockDim.x) {
      const int d = r / nu, k = r - d * nu;
      const double gbar = w * (-G[d * Pg + k] + al[d] * cv[k]);
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
      const double v = sigmoid_d(p.lam[d]) * w * (-0.5 * ds[d] + 0.5 * al[d] * al[d]);
      const long long o = (long long)n * D + d;
      if (s == 0) p.glam_part[o] = v; else p.glam_part[o] += v;
    }
  }
  if (tid == 0) {
    p.ve[n] = bad ? nan("") : ve_acc;
    p.info[n] = bad;
  }
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
  return (long long)sizeof(double) * (2LL * Dp * Pg + 2LL * Dp * Pd + 5 * Dp + Np + 32);
}

extern "C" int fcest_wishart_loglik(const double* fmean, const double* fvar, long long ldf,
                                    const double* eps, const double* y, const double* A, int a_mode,
                                    const double* lam, int S, int N, int D, int nu, double* ve,
                                    double* g_fmean, double* g_fvar, long long ldg, double* gA,
                                    double* glam, double* ws_gA_part, double* ws_glam_part,
                                    int* info, cudaStream_t stream) {
  if (N <= 0) return 0;
  const long long smem = fcest_wishart_loglik_smem_bytes(D, nu);
  if (smem > 227 * 1024) return FCEST_ERR_UNSUPPORTED;
  LikParams p;
  p.fmean = fmean; p.fvar = fvar; p.ldf = ldf; p.eps = eps; p.y = y; p.A = A; p.a_mode = a_mode;
  p.lam = lam; p.S = S; p.N = N; p.D = D; p.nu = nu; p.ve = ve; p.g_fmean = g_fmean;
  p.g_fvar = g_fvar; p.ldg = ldg; p.gA_part = ws_gA_part; p.glam_part = ws_glam_part; p.info = info;
  p.need_grad = (g_fmean != nullptr);
  if (p.need_grad && (!g_fvar || !ws_gA_part || !ws_glam_part || !gA || !glam)) return FCEST_ERR_ARGUMENT;
  cudaError_t e = cudaFuncSetAttribute(wishart_loglik_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  wishart_loglik_kernel<<<N, 256, smem, stream>>>(p);
  FCEST_CHECK_LAUNCH();
  if (p.need_grad) {
    const int R = D * nu;
    const dim3 cb(32, 32);
    if (a_mode == 0) {
      colsum_kernel<<<(R + 31) / 32, cb, 0, stream>>>(ws_gA_part, R, N, R, gA, 0);
    } else {
      colsum_kernel<<<(nu + 31) / 32, cb, 0, stream>>>(ws_gA_part, R, N, R, gA, nu);
    }
    FCEST_CHECK_LAUNCH();
    colsum_kernel<<<(D + 31) / 32, cb, 0, stream>>>(ws_glam_part, D, N, D, glam, 0);
    FCEST_CHECK_LAUNCH();
  }
  return 0;
}
