// Streaming covariance / correlation moments for prediction, D <= 56: register-tiled variant.
//
// Reference: fcest/models/wishart_process.py:169-212 / :462-504 (_get_cov_samples) and :120-167 / :392-460
// (predict_cov / predict_corr: mean and population std over S* samples; correlation per sample as in
// fcest/helpers/array_operations.py:56-72).
//
// One CTA (4 warps) per time step.  A o fmean and A o sqrt(fvar) are staged once in shared memory; for every
// MC sample each warp accumulates its quarter of the lower 8 x 8 tiles of Sigma = G G^T in registers, building
// the G fragments on the fly from the staging and the noise streamed from global memory (G is never stored),
// and folds the tile into its Welford (mean, M2) registers.  For correlations the warp also accumulates the
// diagonal of Sigma for the row blocks its fragments cover, so no exchange between warps (and no block
// barrier) is needed per sample.  The CTA-per-time-step kernel in predict.cu (G through shared memory, 2-3
// block barriers per sample) is kept for D > 56.
#include <math.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

struct MomParamsW {
  const double* fmean; const double* fvar; long long ldf;
  const double* eps;  // (Sc, N, R)
  const double* A; int a_mode; const double* lam;
  int Sc, N, D, nu, corr;
  long long cnt0;
  double* mean; double* m2;
  double* samples;
  int finalize; double* out_mean; double* out_std;
};

constexpr int PW_WARPS = 4;

constexpr int pw_tile_ib(int t) { int ib = 0; while ((ib + 1) * (ib + 2) / 2 <= t) ++ib; return ib; }
constexpr int pw_tile_jb(int t) { return t - pw_tile_ib(t) * (pw_tile_ib(t) + 1) / 2; }
template <int NB> constexpr int pw_tpw() { return (NB * (NB + 1) / 2 + PW_WARPS - 1) / PW_WARPS; }
// does warp W (tiles [W * TPW, (W + 1) * TPW)) touch row / column block b ?
template <int NB> constexpr bool pw_needs(int W, int b) {
  const int NT = NB * (NB + 1) / 2, TPW = pw_tpw<NB>();
  for (int t = W * TPW; t < (W + 1) * TPW && t < NT; ++t)
    if (pw_tile_ib(t) == b || pw_tile_jb(t) == b) return true;
  return false;
}

template <int NB, int W>
__device__ __forceinline__ void pw_body(const MomParamsW& p, const double* am, const double* as_, const double* spl,
                                        int PA, int n, int lane) {
  constexpr int NT = NB * (NB + 1) / 2, TPW = pw_tpw<NB>();
  constexpr int T0 = W * TPW, T1 = (T0 + TPW < NT) ? T0 + TPW : NT;
  const int D = p.D, nu = p.nu, R = D * nu;
  const int K4 = (nu + 3) & ~3;
  const int gid = lane >> 2, tig = lane & 3;
  const bool track = (p.mean != nullptr);
  double mean[TPW][2], m2[TPW][2];
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    mean[q][0] = mean[q][1] = m2[q][0] = m2[q][1] = 0.0;
    if (T0 + q < T1 && track && p.cnt0 > 0) {
      const int row = 8 * pw_tile_ib(T0 + q) + gid, col = 8 * pw_tile_jb(T0 + q) + 2 * tig;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (row < D && col + h <= row) {
          const long long o = ((long long)n * D + row) * D + col + h;
          mean[q][h] = p.mean[o];
          m2[q][h] = p.m2[o];
        }
      }
    }
  }
  for (int s = 0; s < p.Sc; ++s) {
    const double* ep = p.eps + ((long long)s * p.N + n) * R;
    double acc[TPW][2], dsum[NB], en[NB];
#pragma unroll
    for (int q = 0; q < TPW; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      dsum[b] = 0.0;
      en[b] = 0.0;
      if (pw_needs<NB>(W, b)) {
        const int row = 8 * b + gid;
        en[b] = (row < D && tig < nu) ? __ldg(ep + row * nu + tig) : 0.0;
      }
    }
    for (int k0 = 0; k0 < K4; k0 += 4) {
      double a[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        a[b] = 0.0;
        if (pw_needs<NB>(W, b)) {
          const int o = (8 * b + gid) * PA + k0 + tig;
          a[b] = am[o] + as_[o] * en[b];
          dsum[b] += a[b] * a[b];
        }
      }
      {  // noise of the next k-step (one step ahead: more pipelining costs a CTA of occupancy)
        const int col = k0 + 4 + tig;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (pw_needs<NB>(W, b)) {
            const int row = 8 * b + gid;
            en[b] = (row < D && col < nu) ? __ldg(ep + row * nu + col) : 0.0;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < TPW; ++q)
        if (T0 + q < T1) dmma(acc[q][0], acc[q][1], a[pw_tile_ib(T0 + q)], a[pw_tile_jb(T0 + q)]);
    }
    // diagonal of Sigma for the blocks this warp covers (row 8 b + gid in every lane of the quad)
    double rinv[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      rinv[b] = 0.0;
      if (pw_needs<NB>(W, b)) {
        double v = dsum[b];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        rinv[b] = rsqrt(v + spl[8 * b + gid]);
      }
    }
    const double icnt = 1.0 / (double)(p.cnt0 + s + 1);
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      if (T0 + q < T1) {
        constexpr int dummy = 0; (void)dummy;
        const int ib = pw_tile_ib(T0 + q), jb = pw_tile_jb(T0 + q);
        const int row = 8 * ib + gid, col = 8 * jb + 2 * tig;
        // 1 / sqrt(Sigma_cc) for the two columns of this lane: held by the lanes with gid' = 2 tig + h
        const double rc0 = __shfl_sync(0xffffffffu, rinv[jb], 4 * (2 * tig)), rc1 = __shfl_sync(0xffffffffu, rinv[jb], 4 * (2 * tig + 1));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = col + h;
          double x = acc[q][h];
          if (row == c) x += spl[row];
          if (p.corr) x = x * rinv[ib] * (h ? rc1 : rc0);
          if (row < D && c <= row) {
            if (p.samples) {
              double* o = p.samples + (((long long)s * p.N + n) * D) * D;
              o[(long long)row * D + c] = x;
              o[(long long)c * D + row] = x;
            }
            const double delta = x - mean[q][h];
            mean[q][h] += delta * icnt;
            m2[q][h] += delta * (x - mean[q][h]);
          }
        }
      }
    }
  }
  if (!track) return;
  const double tot = (double)(p.cnt0 + p.Sc);
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    if (T0 + q < T1) {
      const int row = 8 * pw_tile_ib(T0 + q) + gid, col = 8 * pw_tile_jb(T0 + q) + 2 * tig;
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
}

template <int NB>
__global__ void __launch_bounds__(PW_WARPS * 32)
cov_moments_warp_kernel(const MomParamsW p) {
  extern __shared__ __align__(16) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu;
  constexpr int Dp = 8 * NB;
  const int PA = pitch4((nu + 3) & ~3);
  double* am = sm;
  double* as_ = am + Dp * PA;
  double* spl = as_ + Dp * PA;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x;
  const double* fm = p.fmean + (long long)n * p.ldf;
  const double* fv = p.fvar + (long long)n * p.ldf;
  for (int e = tid; e < 2 * Dp * PA; e += blockDim.x) sm[e] = 0.0;
  for (int d = tid; d < Dp; d += blockDim.x) spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  __syncthreads();
  for (int r = tid; r < R; r += blockDim.x) {
    const int d = r / nu, k = r - d * nu;
    const double a = p.a_mode == 0 ? __ldg(p.A + r) : __ldg(p.A + k);
    am[d * PA + k] = a * fm[r];
    as_[d * PA + k] = a * sqrt(fv[r]);
  }
  __syncthreads();
  if (warp == 0) pw_body<NB, 0>(p, am, as_, spl, PA, n, lane);
  else if (warp == 1) pw_body<NB, 1>(p, am, as_, spl, PA, n, lane);
  else if (warp == 2) pw_body<NB, 2>(p, am, as_, spl, PA, n, lane);
  else pw_body<NB, 3>(p, am, as_, spl, PA, n, lane);
}

bool cov_moments_warp_supported(int D, int nu) { return D >= 1 && D <= 56 && nu >= 1 && nu <= 4096; }

template <int NB>
static int pw_launch(const MomParamsW& p, cudaStream_t stream) {
  const int PA = pitch4((p.nu + 3) & ~3);
  const size_t smem = sizeof(double) * (2 * (size_t)(8 * NB) * PA + 8 * NB);
  if (smem > 227 * 1024) return -1;
  auto kern = cov_moments_warp_kernel<NB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<p.N, PW_WARPS * 32, smem, stream>>>(p);
  return 0;
}

int launch_cov_moments_warp(const double* fmean, const double* fvar, long long ldf, const double* eps, const double* A,
                            int a_mode, const double* lam, int Sc, int N, int D, int nu, int corr, long long cnt0,
                            double* mean, double* m2, double* samples, int finalize, double* out_mean,
                            double* out_std, cudaStream_t stream) {
  MomParamsW p;
  p.fmean = fmean; p.fvar = fvar; p.ldf = ldf; p.eps = eps; p.A = A; p.a_mode = a_mode; p.lam = lam; p.Sc = Sc;
  p.N = N; p.D = D; p.nu = nu; p.corr = corr; p.cnt0 = cnt0; p.mean = mean; p.m2 = m2; p.samples = samples;
  p.finalize = finalize; p.out_mean = out_mean; p.out_std = out_std;
  switch ((D + 7) / 8) {
    case 1: return pw_launch<1>(p, stream);
    case 2: return pw_launch<2>(p, stream);
    case 3: return pw_launch<3>(p, stream);
    case 4: return pw_launch<4>(p, stream);
    case 5: return pw_launch<5>(p, stream);
    case 6: return pw_launch<6>(p, stream);
    case 7: return pw_launch<7>(p, stream);
    default: return -1;
  }
}

}  // namespace fcest
