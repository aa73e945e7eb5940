// Prediction-side kernels: covariance / correlation samples and their streaming moments.
//
// Reference: fcest/models/wishart_process.py:169-212 and :462-504 (_get_cov_samples),
// :120-167 / :392-460 (predict_cov / predict_corr: mean and population std over S* samples),
// fcest/helpers/array_operations.py:56-72 (cov -> corr per sample), :75-103 (PD assertion).
// The reference materialises (S*, N*, D, D) samples (7.2 GB at D=50, N*=1200, S*=300) and reduces
// them afterwards; cov_moments_kernel keeps a Welford (mean, M2) pair per matrix entry in registers
// and streams the samples through shared memory, so nothing of that size ever exists.
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

bool cov_moments_warp_supported(int D, int nu);
int launch_cov_moments_warp(const double* fmean, const double* fvar, long long ldf, const double* eps, const double* A,
                            int a_mode, const double* lam, int Sc, int N, int D, int nu, int corr, long long cnt0,
                            double* mean, double* m2, double* samples, int finalize, double* out_mean,
                            double* out_std, cudaStream_t stream);

__host__ __device__ inline int pitch4p(int x) {
  int p = (x + 7) / 8 * 8 + 4;
  return (p - 8 >= x) ? p - 8 : p;
}

constexpr int kMaxTilesPerWarp = 8;

struct MomParams {
  const double* fmean; const double* fvar; long long ldf;
  const double* eps;  // (Sc, N, R)
  const double* A; int a_mode; const double* lam;
  int Sc, N, D, nu, corr;
  long long cnt0;       // samples already folded into (mean, m2)
  double* mean; double* m2;   // (N, D, D) running state (lower triangle used)
  double* samples;      // optional (Sc, N, D, D) output of the raw samples
  int finalize; double* out_mean; double* out_std;  // (N, D, D)
};

__global__ void __launch_bounds__(256)
cov_moments_kernel(const MomParams p) {
  extern __shared__ __align__(16) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu;
  const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8, Pg = pitch4p(Np);
  double* G = sm;
  double* dg = G + Dp * Pg;   // Dp: diagonal of the current sample
  double* spl = dg + Dp;      // Dp
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nbD = Dp / 8, ntiles = nbD * (nbD + 1) / 2;
  const double* fm = p.fmean + (long long)n * p.ldf;
  const double* fv = p.fvar + (long long)n * p.ldf;
  for (int d = tid; d < Dp; d += blockDim.x) spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;

  double mean[kMaxTilesPerWarp][2], m2[kMaxTilesPerWarp][2];
  const bool track = (p.mean != nullptr);
  if (track) {
#pragma unroll
    for (int q = 0; q < kMaxTilesPerWarp; ++q) {
      const int t = warp + q * nwarps;
      if (t >= ntiles) continue;
      int ib = 0, rem = t;
      while (rem > ib) { rem -= ib + 1; ++ib; }
      const int row = 8 * ib + gid, col = 8 * rem + 2 * tig;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool ok = p.cnt0 > 0 && row < D && col + h <= row;
        const long long o = ((long long)n * D + row) * D + col + h;
        mean[q][h] = ok ? p.mean[o] : 0.0;
        m2[q][h] = ok ? p.m2[o] : 0.0;
      }
    }
  }

  for (int s = 0; s < p.Sc; ++s) {
    const double* ep = p.eps + ((long long)s * p.N + n) * R;
    __syncthreads();
    for (int e = tid; e < Dp * Pg; e += blockDim.x) {
      const int d = e / Pg, k = e - d * Pg;
      double v = 0.0;
      if (d < D && k < nu) {
        const int r = d * nu + k;
        v = (p.a_mode == 0 ? p.A[r] : p.A[k]) * (fm[r] + sqrt(fv[r]) * ep[r]);
      }
      G[e] = v;
    }
    __syncthreads();
    double val[kMaxTilesPerWarp][2];
    {
#pragma unroll
      for (int q = 0; q < kMaxTilesPerWarp; ++q) {
        const int t = warp + q * nwarps;
        if (t >= ntiles) continue;
        int ib = 0, rem = t;
        while (rem > ib) { rem -= ib + 1; ++ib; }
        const int jb = rem;
        double c0 = 0.0, c1 = 0.0;
        const double* ga = G + (8 * ib + gid) * Pg + tig;
        const double* gb = G + (8 * jb + gid) * Pg + tig;
        for (int k0 = 0; k0 < Np; k0 += 4) dmma(c0, c1, ga[k0], gb[k0]);
        const int row = 8 * ib + gid, col = 8 * jb + 2 * tig;
        if (row == col) { c0 += spl[row]; dg[row] = c0; }
        if (row == col + 1) { c1 += spl[row]; dg[row] = c1; }
        val[q][0] = c0;
        val[q][1] = c1;
      }
    }
    if (p.corr) __syncthreads();
    const double cnt = (double)(p.cnt0 + s + 1), icnt = 1.0 / cnt;
#pragma unroll
    for (int q = 0; q < kMaxTilesPerWarp; ++q) {
      const int t = warp + q * nwarps;
      if (t >= ntiles) continue;
      int ib = 0, rem = t;
      while (rem > ib) { rem -= ib + 1; ++ib; }
      const int row = 8 * ib + gid, col = 8 * rem + 2 * tig;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = col + h;
        if (row >= D || c > row) continue;
        double x = val[q][h];
        if (p.corr) x = x / sqrt(dg[row] * dg[c]);
        if (p.samples) {
          double* o = p.samples + (((long long)s * p.N + n) * D) * D;
          o[(long long)row * D + c] = x;
          o[(long long)c * D + row] = x;
        }
        if (track) {
          const double delta = x - mean[q][h];
          mean[q][h] += delta * icnt;
          m2[q][h] += delta * (x - mean[q][h]);
        }
      }
    }
  }
  if (!track) return;
  const double tot = (double)(p.cnt0 + p.Sc);
#pragma unroll
  for (int q = 0; q < kMaxTilesPerWarp; ++q) {
    const int t = warp + q * nwarps;
    if (t >= ntiles) continue;
    int ib = 0, rem = t;
    while (rem > ib) { rem -= ib + 1; ++ib; }
    const int row = 8 * ib + gid, col = 8 * rem + 2 * tig;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = col + h;
      if (row >= D || c > row) continue;
      const long long o = ((long long)n * D + row) * D + c, ot = ((long long)n * D + c) * D + row;
      p.mean[o] = mean[q][h];
      p.m2[o] = m2[q][h];
      if (p.finalize) {
        const double sd = sqrt(fmax(m2[q][h], 0.0) / tot);
        p.out_mean[o] = mean[q][h]; p.out_mean[ot] = mean[q][h];
        p.out_std[o] = sd; p.out_std[ot] = sd;
      }
    }
  }
}

// Large-D variant (D > 80): G lives in a per-CTA slot of a global scratch buffer and the Welford state
// is read-modify-written in the (N, D, D) mean / m2 arrays (each entry is owned by one thread), with a
// persistent grid over time steps.  Same arithmetic as cov_moments_kernel.
__global__ void __launch_bounds__(256)
cov_moments_big_kernel(const MomParams p, double* __restrict__ scratch, long long stride) {
  extern __shared__ __align__(16) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu;
  const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8, Pg = pitch4p(Np);
  double* G = scratch + (long long)blockIdx.x * stride;
  double* dg = sm;
  double* spl = dg + Dp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nbD = Dp / 8, ntiles = nbD * (nbD + 1) / 2;
  const bool track = (p.mean != nullptr);
  for (int d = tid; d < Dp; d += blockDim.x) spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
    const double* fm = p.fmean + (long long)n * p.ldf;
    const double* fv = p.fvar + (long long)n * p.ldf;
    for (int s = 0; s < p.Sc; ++s) {
      const double* ep = p.eps + ((long long)s * p.N + n) * R;
      __syncthreads();
      for (int e = tid; e < Dp * Pg; e += blockDim.x) {
        const int d = e / Pg, k = e - d * Pg;
        double v = 0.0;
        if (d < D && k < nu) {
          const int r = d * nu + k;
          v = (p.a_mode == 0 ? p.A[r] : p.A[k]) * (fm[r] + sqrt(fv[r]) * ep[r]);
        }
        G[e] = v;
      }
      __syncthreads();
      if (p.corr) {  // diagonal first (needed by every entry of the correlation matrix)
        for (int d = tid; d < D; d += blockDim.x) {
          double sq = 0.0;
          for (int k = 0; k < nu; ++k) sq += G[d * Pg + k] * G[d * Pg + k];
          dg[d] = sq + spl[d];
        }
        __syncthreads();
      }
      const double cnt = (double)(p.cnt0 + s + 1), icnt = 1.0 / cnt;
      const bool first = (p.cnt0 + s == 0);
      const bool last = p.finalize && (s == p.Sc - 1);
      for (int t = warp; t < ntiles; t += nwarps) {
        int ib = 0, rem = t;
        while (rem > ib) { rem -= ib + 1; ++ib; }
        const int jb = rem;
        double c[2] = {0.0, 0.0};
        const double* ga = G + (8 * ib + gid) * Pg + tig;
        const double* gb = G + (8 * jb + gid) * Pg + tig;
        for (int k0 = 0; k0 < Np; k0 += 4) dmma(c[0], c[1], ga[k0], gb[k0]);
        const int row = 8 * ib + gid, col = 8 * jb + 2 * tig;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = col + h;
          if (row >= D || cc > row) continue;
          double x = c[h] + (row == cc ? spl[row] : 0.0);
          if (p.corr) x = x / sqrt(dg[row] * dg[cc]);
          const long long o = ((long long)n * D + row) * D + cc, ot = ((long long)n * D + cc) * D + row;
          if (p.samples) {
            double* so = p.samples + (((long long)s * p.N + n) * D) * D;
            so[(long long)row * D + cc] = x;
            so[(long long)cc * D + row] = x;
          }
          if (track) {
            double mu = first ? 0.0 : p.mean[o], m2 = first ? 0.0 : p.m2[o];
            const double delta = x - mu;
            mu += delta * icnt;
            m2 += delta * (x - mu);
            p.mean[o] = mu;
            p.m2[o] = m2;
            if (last) {
              const double sd = sqrt(fmax(m2, 0.0) / cnt);
              p.out_mean[o] = mu; p.out_mean[ot] = mu;
              p.out_std[o] = sd; p.out_std[ot] = sd;
            }
          }
        }
      }
    }
  }
}

// info[b] = -1 if |A - A^T| > atol anywhere (np.allclose(A - A.T, 0, rtol) -> atol 1e-8), else the
// LAPACK-style Cholesky info (0 = positive definite).  One CTA per matrix, smem D x (D+1).
__global__ void __launch_bounds__(128)
batched_chol_info_kernel(const double* __restrict__ A, int B, int D, double atol, int* __restrict__ info,
                         double* scratch) {
  extern __shared__ double smd[];
  __shared__ int flag;
  const int b = blockIdx.x, P = D + 1;
  // working copy: shared memory when it fits, else this matrix's slot of the caller's global scratch
  double* sm = scratch ? scratch + (long long)b * D * P : smd;
  const double* a = A + (long long)b * D * D;
  if (threadIdx.x == 0) flag = 0;
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) sm[(e / D) * P + (e % D)] = a[e];
  __syncthreads();
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    if (!(fabs(sm[i * P + j] - sm[j * P + i]) <= atol)) flag = 1;
  }
  __syncthreads();
  if (flag) { if (threadIdx.x == 0) info[b] = -1; return; }
  int bad = 0;
  for (int j = 0; j < D; ++j) {
    __syncthreads();
    const double piv = sm[j * P + j];
    if (!(piv > 0.0)) { bad = j + 1; break; }
    const double inv = 1.0 / sqrt(piv);
    for (int i = j + 1 + threadIdx.x; i < D; i += blockDim.x) sm[i * P + j] *= inv;
    __syncthreads();
    for (int e = threadIdx.x; e < (D - j - 1) * (D - j - 1); e += blockDim.x) {
      const int i = j + 1 + e / (D - j - 1), k = j + 1 + e % (D - j - 1);
      if (k <= i) sm[i * P + k] -= sm[i * P + j] * sm[k * P + j];
    }
  }
  if (threadIdx.x == 0) info[b] = bad;
}

}  // namespace fcest

using namespace fcest;

static const int kMomBigCtas = 148 * 4;
// bytes of global scratch fcest_cov_moments needs (0 while one D x nu sample fits shared memory)
extern "C" long long fcest_cov_moments_scratch_bytes(int D, int nu) {
  const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8, Pg = pitch4p(Np), nb = Dp / 8;
  const size_t smem = sizeof(double) * ((size_t)Dp * Pg + 2 * Dp);
  const bool bigD = (nb * (nb + 1) / 2 > kMaxTilesPerWarp * 8) || smem > 227 * 1024;
  return bigD ? (long long)sizeof(double) * Dp * Pg * kMomBigCtas : 0;
}

extern "C" int fcest_cov_moments(const double* fmean, const double* fvar, long long ldf,
                                 const double* eps, const double* A, int a_mode, const double* lam,
                                 int Sc, int N, int D, int nu, int corr, long long cnt0, double* mean,
                                 double* m2, double* samples, int finalize, double* out_mean,
                                 double* out_std, double* scratch, cudaStream_t stream) {
  if (N <= 0 || Sc <= 0) return 0;
  const int Dp = (D + 7) / 8 * 8, Np = (nu + 7) / 8 * 8, Pg = pitch4p(Np);
  const int nb = Dp / 8;
  const size_t smem = sizeof(double) * ((size_t)Dp * Pg + 2 * Dp);
  const bool bigD = (nb * (nb + 1) / 2 > kMaxTilesPerWarp * 8) || smem > 227 * 1024;
  if (D > 400 || nu > 400) return FCEST_ERR_UNSUPPORTED;
  if (bigD && scratch == nullptr) return FCEST_ERR_ARGUMENT;
  MomParams p;
  p.fmean = fmean; p.fvar = fvar; p.ldf = ldf; p.eps = eps; p.A = A; p.a_mode = a_mode; p.lam = lam;
  p.Sc = Sc; p.N = N; p.D = D; p.nu = nu; p.corr = corr; p.cnt0 = cnt0; p.mean = mean; p.m2 = m2;
  p.samples = samples; p.finalize = finalize; p.out_mean = out_mean; p.out_std = out_std;
  if (mean && !m2) return FCEST_ERR_ARGUMENT;
  if (finalize && (!mean || !out_mean || !out_std)) return FCEST_ERR_ARGUMENT;
  // D <= 56: register-tiled kernel (predict_warp.cu); FCEST_MOM_IMPL=cta forces the shared-memory kernel below
  static const bool force_cta = [] { const char* v = getenv("FCEST_MOM_IMPL"); return v && !strcmp(v, "cta"); }();
  if (!force_cta && cov_moments_warp_supported(D, nu)) {
    const int rc = launch_cov_moments_warp(fmean, fvar, ldf, eps, A, a_mode, lam, Sc, N, D, nu, corr, cnt0, mean, m2,
                                           samples, finalize, out_mean, out_std, stream);
    if (rc > 0) return rc;
    if (rc == 0) { FCEST_CHECK_LAUNCH(); return 0; }
  }
  if (bigD) {
    const int ctas = N < kMomBigCtas ? N : kMomBigCtas;
    cov_moments_big_kernel<<<ctas, 256, sizeof(double) * 2 * Dp, stream>>>(p, scratch, (long long)Dp * Pg);
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(cov_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  cov_moments_kernel<<<N, 256, smem, stream>>>(p);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_batched_chol_info(const double* A, int B, int D, double atol, int* info,
                                       double* scratch, cudaStream_t stream) {
  if (B <= 0) return 0;
  size_t smem = sizeof(double) * (size_t)D * (D + 1);
  if (smem > 227 * 1024) {  // large D: needs B * D * (D+1) doubles of global scratch
    if (scratch == nullptr) return FCEST_ERR_ARGUMENT;
    smem = 0;
  } else {
    scratch = nullptr;
    cudaError_t e = cudaFuncSetAttribute(batched_chol_info_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  batched_chol_info_kernel<<<B, 128, smem, stream>>>(A, B, D, atol, info, scratch);
  FCEST_CHECK_LAUNCH();
  return 0;
}
