// Packed-quadratic-form projection kernels for the whitened SVGP / VGP conditional.
//
// GPflow's base_conditional materialises LTA = tril(q_sqrt_r)^T P for every
// latent r, an (R, M, N) tensor.  Here the predictive variance is rewritten as a quadratic form
//     v[n, r] = c[n] + p_n^T S_r p_n ,   S_r = L_r L_r^T ,   p_n = P[:, n]
// and both sides are packed over the lower triangle (K = M(M+1)/2 entries, row-major (i,j), i>=j):
//     Sk[r, (i,j)] = S_r[i,j]                    (tri_syrk_pack_kernel)
//     Pk[n, (i,j)] = w_ij P[i,n] P[j,n], w=1|2   (pack_outer_kernel)
// so that  V = Pk Sk^T,  dELBO/dSk = g^T Pk  and  T = g Sk  are three *plain* fp64 GEMMs (gemm.cu)
// with the algorithmic flop count 3 M^2 N R and no recompute.  The kernels in this file are the
// (cheap) packing / unpacking steps around those GEMMs.
#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

// ---------------------------------------------------------------------------------------------
// Pk[n, idx(i,j)] = w_ij * P[i,n] * P[j,n];  bias[n] = (use_prior ? var - sum_m P[m,n]^2 : 0)
// grid = N, block = 256, dynamic smem = M doubles (+32 for the reduction)
// ---------------------------------------------------------------------------------------------
__global__ void pack_outer_kernel(const double* __restrict__ P, long long ldp, int M, int N,
                                  double* __restrict__ Pk, long long ldk, double* __restrict__ bias,
                                  const double* __restrict__ kvar, int use_prior) {
  extern __shared__ double sm[];
  double* p = sm;
  double* red = sm + M;
  const int n = blockIdx.x;
  double sq = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double v = P[(long long)m * ldp + n];
    p[m] = v;
    sq += v * v;
  }
  sq = block_sum(sq, red);  // contains __syncthreads -> p[] visible
  if (threadIdx.x == 0) bias[n] = use_prior ? (kvar[0] - sq) : 0.0;
  const long long K = (long long)M * (M + 1) / 2;
  double* out = Pk + (long long)n * ldk;
  for (long long idx = threadIdx.x; idx < ldk; idx += blockDim.x) {
    double v = 0.0;
    if (idx < K) {
      int i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
      while ((long long)(i + 1) * (i + 2) / 2 <= idx) ++i;
      while ((long long)i * (i + 1) / 2 > idx) --i;
      const int j = (int)(idx - (long long)i * (i + 1) / 2);
      v = p[i] * p[j];
      if (i != j) v *= 2.0;
    }
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Sk[r, idx(i,j)] = sum_{k<=j} L_r[i,k] L_r[j,k]  (L_r = tril(q_sqrt_r)), DMMA 8x8 tiles read
// straight from global/L1 (L_r is 8*M*M bytes, L1/L2 resident for the CTA that owns r).
// Also emits the two KL reductions per latent: klp[r,0] = ||tril L_r||_F^2 = tr(S_r),
// klp[r,1] = sum_m log(L_r[m,m]^2).                              grid = R, block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tri_syrk_pack_kernel(const double* __restrict__ q_sqrt, int M, double* __restrict__ Sk,
                     long long ldk, double* __restrict__ klp) {
  __shared__ double red[32];
  const int r = blockIdx.x;
  const double* __restrict__ L = q_sqrt + (long long)r * M * M;
  double* __restrict__ out = Sk + (long long)r * ldk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nb = (M + 7) / 8;
  const int ntiles = nb * (nb + 1) / 2;
  double tr = 0.0;
  // heavier tiles (large jb) first so the tail is short; tiles are dealt round-robin to warps
  for (int t = ntiles - 1 - warp; t >= 0; t -= nwarps) {
    int ib = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((ib + 1) * (ib + 2) / 2 <= t) ++ib;
    while (ib * (ib + 1) / 2 > t) --ib;
    const int jb = t - ib * (ib + 1) / 2;
    const int ra = 8 * ib + gid, rb = 8 * jb + gid;
    const double* La = L + (long long)ra * M;
    const double* Lb = L + (long long)rb * M;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < 8 * (jb + 1); k0 += 4) {
      const int k = k0 + tig;
      const double a = (ra < M && k <= ra) ? La[k] : 0.0;
      const double b = (rb < M && k <= rb) ? Lb[k] : 0.0;
      dmma(c0, c1, a, b);
    }
    const int row = ra, col = 8 * jb + 2 * tig;
    if (row < M) {
      const long long base = (long long)row * (row + 1) / 2;
      if (col <= row) { out[base + col] = c0; if (col == row) tr += c0; }
      if (col + 1 <= row) { out[base + col + 1] = c1; if (col + 1 == row) tr += c1; }
    }
  }
  // zero the row padding [K, ldk)
  const long long K = (long long)M * (M + 1) / 2;
  for (long long idx = K + threadIdx.x; idx < ldk; idx += blockDim.x) out[idx] = 0.0;
  double ld = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double d = L[(long long)m * M + m];
    ld += log(d * d);
  }
  tr = block_sum(tr, red);
  ld = block_sum(ld, red);
  if (threadIdx.x == 0) { klp[2 * r] = tr; klp[2 * r + 1] = ld; }
}

// ---------------------------------------------------------------------------------------------
// dq_sqrt_r = tril( 2 W_r L_r - kl * L_r ) + kl * diag(1 / L_r[m,m]),
//   W_r symmetric, unpacked from Gk[r, idx(i,j)] = w_ij W_r[i,j]   (Gk = g_var^T Pk).
// This is d ELBO / d q_sqrt including the KL term (kl = 1) or the bare projection adjoint (kl = 0).
// grid = R, block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tri_grad_kernel(const double* __restrict__ Gk, long long ldk, const double* __restrict__ q_sqrt,
                int M, double* __restrict__ dq_sqrt, double kl) {
  const int r = blockIdx.x;
  const double* __restrict__ L = q_sqrt + (long long)r * M * M;
  const double* __restrict__ G = Gk + (long long)r * ldk;
  double* __restrict__ out = dq_sqrt + (long long)r * M * M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nb = (M + 7) / 8;
  const int ntiles = nb * (nb + 1) / 2;
  // tile (ib, jb): cost ~ (nb - jb); enumerate column-major so that heavy (small jb) tiles come first
  for (int t = warp; t < ntiles; t += nwarps) {
    // column-major enumeration of the lower triangle: column jb holds (nb - jb) tiles
    int jb = 0, rem = t;
    while (rem >= nb - jb) { rem -= nb - jb; ++jb; }
    const int ib = jb + rem;
    const int ra = 8 * ib + gid;       // row of W held by this lane (A fragment)
    const int cb = 8 * jb + gid;       // column of L held by this lane (B fragment)
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 8 * jb; k0 < 8 * nb; k0 += 4) {
      const int k = k0 + tig;
      double a = 0.0, b = 0.0;
      if (ra < M && k < M) {
        a = (k <= ra) ? G[(long long)ra * (ra + 1) / 2 + k] : G[(long long)k * (k + 1) / 2 + ra];
        if (k != ra) a *= 0.5;
      }
      if (k < M && cb < M && cb <= k) b = L[(long long)k * M + cb];
      dmma(c0, c1, a, b);
    }
    const int row = ra, col = 8 * jb + 2 * tig;
    if (row < M) {
      if (col < M) {
        double v = 0.0;
        if (col <= row) {
          const double l = L[(long long)row * M + col];
          v = 2.0 * c0 - kl * l;
          if (col == row) v += kl / l;
        }
        out[(long long)row * M + col] = v;
      }
      if (col + 1 < M) {
        double v = 0.0;
        if (col + 1 <= row) {
          const double l = L[(long long)row * M + col + 1];
          v = 2.0 * c1 - kl * l;
          if (col + 1 == row) v += kl / l;
        }
        out[(long long)row * M + col + 1] = v;
      }
    }
  }
  // strictly-upper tiles are never visited: zero them
  for (long long e = threadIdx.x; e < (long long)M * M; e += blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    if ((j >> 3) > (i >> 3)) out[e] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// dP[:, n] (+)= 2 T_n p_n - 2 gs[n] p_n * use_prior,   T_n symmetric from packed Tk[n, idx(i,j)].
// gs[n] = sum_r g_var[n, r] (row sums, computed here).   grid = N, block = 256, smem = 2M + 32
// ---------------------------------------------------------------------------------------------
__global__ void sym_matvec_kernel(const double* __restrict__ Tk, long long ldk,
                                  const double* __restrict__ P, long long ldp, int M, int N,
                                  const double* __restrict__ g_var, long long ldg, int R,
                                  int use_prior, double* __restrict__ dP, long long lddp,
                                  double* __restrict__ gs_out, int accumulate) {
  extern __shared__ double sm[];
  double* p = sm;
  double* acc = sm + M;
  double* red = sm + 2 * M;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int m = threadIdx.x; m < M; m += blockDim.x) p[m] = P[(long long)m * ldp + n];
  double gs = 0.0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) gs += g_var[(long long)n * ldg + r];
  gs = block_sum(gs, red);
  if (threadIdx.x == 0 && gs_out) gs_out[n] = gs;
  const double* T = Tk + (long long)n * ldk;
  // row part: acc[i] = sum_{j<=i} T[i,j] p[j]   (one warp per row, coalesced)
  for (int i = warp; i < M; i += nwarps) {
    const double* Ti = T + (long long)i * (i + 1) / 2;
    double s = 0.0;
    for (int j = lane; j <= i; j += 32) s += Ti[j] * p[j];
    s = warp_sum(s);
    if (lane == 0) acc[i] = s;
  }
  __syncthreads();
  // column part: out[j] += sum_{i>j} T[i,j] p[i]   (thread per column, coalesced across threads)
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    double s = acc[j];
    for (int i = j + 1; i < M; ++i) s += T[(long long)i * (i + 1) / 2 + j] * p[i];
    double v = 2.0 * s;
    if (use_prior) v -= 2.0 * gs * p[j];
    double* dst = dP + (long long)j * lddp + n;
    *dst = accumulate ? (*dst + v) : v;
  }
}

}  // namespace fcest

using namespace fcest;

extern "C" int fcest_pack_outer(const double* P, long long ldp, int M, int N, double* Pk,
                                long long ldk, double* bias, const double* kvar, int use_prior,
                                cudaStream_t stream) {
  if (N <= 0) return 0;
  const size_t smem = sizeof(double) * (M + 32);
  if (smem > 200 * 1024) return FCEST_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(pack_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pack_outer_kernel<<<N, 256, smem, stream>>>(P, ldp, M, N, Pk, ldk, bias, kvar, use_prior);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_tri_syrk_pack(const double* q_sqrt, int R, int M, double* Sk, long long ldk,
                                   double* kl_parts, cudaStream_t stream) {
  if (R <= 0) return 0;
  tri_syrk_pack_kernel<<<R, 256, 0, stream>>>(q_sqrt, M, Sk, ldk, kl_parts);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_tri_grad(const double* Gk, long long ldk, const double* q_sqrt, int R, int M,
                              double* dq_sqrt, double kl, cudaStream_t stream) {
  if (R <= 0) return 0;
  tri_grad_kernel<<<R, 256, 0, stream>>>(Gk, ldk, q_sqrt, M, dq_sqrt, kl);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_sym_matvec(const double* Tk, long long ldk, const double* P, long long ldp,
                                int M, int N, const double* g_var, long long ldg, int R,
                                int use_prior, double* dP, long long lddp, double* gs_out,
                                int accumulate, cudaStream_t stream) {
  if (N <= 0) return 0;
  const size_t smem = sizeof(double) * (2 * M + 32);
  if (smem > 200 * 1024) return FCEST_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sym_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sym_matvec_kernel<<<N, 256, smem, stream>>>(Tk, ldk, P, ldp, M, N, g_var, ldg, R, use_prior, dP,
                                              lddp, gs_out, accumulate);
  FCEST_CHECK_LAUNCH();
  return 0;
}
