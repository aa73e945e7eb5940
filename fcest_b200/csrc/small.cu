// Small-matrix and elementwise kernels around the GP conditional:
// Matern-5/2 kernel matrices (+ adjoint), single-matrix Cholesky / triangular inverse,
// masks / transposes / reductions, softplus parameter transforms, KL assembly, Keras-Adam.
// None of these is a throughput kernel (O(M^2) .. O(M^3) on one M x M matrix, M ~ 200); they exist
// so that the whole ELBO + gradient evaluation stays on the device with no host round trips.
#include <math.h>

#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

// ---------------------------------------------------------------------------------------------
// Matern-5/2, restating gpflow.kernels.Matern52 / square_distance op for op:
//   xs = x / ls;  r2 = -2 <xs, xs'> + (|xs|^2 + |xs'|^2);  r = sqrt(max(r2, 1e-36))
//   k = var (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)            (+ jitter on the diagonal if asked)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double scaled_r2(const double* x1, const double* x2, int dim, double ls) {
  double cross = 0.0, a = 0.0, b = 0.0;
  for (int d = 0; d < dim; ++d) {
    const double u = x1[d] / ls, v = x2[d] / ls;
    cross = __dadd_rn(cross, __dmul_rn(u, v));
    a = __dadd_rn(a, __dmul_rn(u, u));
    b = __dadd_rn(b, __dmul_rn(v, v));
  }
  return __dadd_rn(__dmul_rn(-2.0, cross), __dadd_rn(a, b));
}

__global__ void matern52_kernel(const double* __restrict__ X1, int n1, const double* __restrict__ X2,
                                int n2, int dim, const double* __restrict__ kvar,
                                const double* __restrict__ kls, double jitter,
                                double* __restrict__ K, long long ldk) {
  const long long total = (long long)n1 * n2;
  const double var = kvar[0], ls = kls[0];
  const double s5 = 2.23606797749979;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / n2), j = (int)(e % n2);
    const double r2 = scaled_r2(X1 + (long long)i * dim, X2 + (long long)j * dim, dim, ls);
    const double r = sqrt(fmax(r2, 1e-36));
    double k = var * (1.0 + s5 * r + 5.0 / 3.0 * (r * r)) * exp(-s5 * r);
    if (i == j) k += jitter;
    K[(long long)i * ldk + j] = k;
  }
}

// Adjoint of matern52_kernel.  gK is (n1 x n2); if `sym` the effective cotangent is
// (gK + gK^T)/2 (n1 == n2, X1 == X2) and gX1 receives both argument slots (factor 2).
// One CTA per row i: part[i] = {d/dvar, d/dls}, gX1[i, :] (optional).
__global__ void matern52_bwd_kernel(const double* __restrict__ X1, int n1,
                                    const double* __restrict__ X2, int n2, int dim,
                                    const double* __restrict__ kvar, const double* __restrict__ kls,
                                    const double* __restrict__ gK, long long ldg, int sym,
                                    double* __restrict__ part, double* __restrict__ gX1) {
  __shared__ double red[32];
  const int i = blockIdx.x;
  const double var = kvar[0], ls = kls[0];
  const double s5 = 2.23606797749979;
  double a_var = 0.0, a_ls = 0.0;
  double a_x[4] = {0.0, 0.0, 0.0, 0.0};  // dim <= 4 supported for input gradients
  const double* x1 = X1 + (long long)i * dim;
  for (int j = threadIdx.x; j < n2; j += blockDim.x) {
    double g = gK[(long long)i * ldg + j];
    if (sym) g = 0.5 * (g + gK[(long long)j * ldg + i]);
    const double* x2 = X2 + (long long)j * dim;
    const double r2 = scaled_r2(x1, x2, dim, ls);
    const double r = sqrt(fmax(r2, 1e-36));
    const double ex = exp(-s5 * r);
    a_var += g * (1.0 + s5 * r + 5.0 / 3.0 * (r * r)) * ex;
    if (r2 > 1e-36) {
      // dk/dr2 = var * f'(r) / (2 r),  f'(r) = -(5/3) r (1 + sqrt5 r) exp(-sqrt5 r)
      const double dk_dr2 = -var * (5.0 / 6.0) * (1.0 + s5 * r) * ex;
      a_ls += g * dk_dr2 * (-2.0 * r2 / ls);
      if (gX1) {
        for (int d = 0; d < dim && d < 4; ++d)
          a_x[d] += g * dk_dr2 * 2.0 * (x1[d] - x2[d]) / (ls * ls);
      }
    }
  }
  a_var = block_sum(a_var, red);
  a_ls = block_sum(a_ls, red);
  if (threadIdx.x == 0) { part[2 * i] = a_var; part[2 * i + 1] = a_ls; }
  if (gX1) {
    for (int d = 0; d < dim && d < 4; ++d) {
      const double v = block_sum(a_x[d], red);
      if (threadIdx.x == 0) gX1[(long long)i * dim + d] = sym ? 2.0 * v : v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Single-matrix Cholesky (lower, in place, strict upper zeroed) - one CTA, global/L2 resident.
// info[0] = 0 or the 1-based index of the first non-positive pivot (LAPACK potrf convention).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
chol_single_kernel(double* __restrict__ A, int M, long long lda, int* __restrict__ info) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  int bad = 0;
  for (int j = 0; j < M; ++j) {
    __syncthreads();
    const double piv = A[(long long)j * lda + j];
    if (!(piv > 0.0)) { bad = j + 1; break; }
    const double dj = sqrt(piv), inv = 1.0 / dj;
    for (int i = j + 1 + tid; i < M; i += blockDim.x) A[(long long)i * lda + j] *= inv;
    __syncthreads();
    if (tid == 0) A[(long long)j * lda + j] = dj;
    for (int i = j + 1 + warp; i < M; i += nwarps) {
      const double lij = A[(long long)i * lda + j];
      for (int k = j + 1 + lane; k <= i; k += 32)
        A[(long long)i * lda + k] -= lij * A[(long long)k * lda + j];
    }
  }
  __syncthreads();
  for (long long e = tid; e < (long long)M * M; e += blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    if (j > i) A[(long long)i * lda + j] = 0.0;
  }
  if (tid == 0) info[0] = bad;
}

// Li = L^-1 for lower-triangular L (one CTA; thread c owns column c, forward substitution).
__global__ void __launch_bounds__(1024)
tri_inv_single_kernel(const double* __restrict__ L, int M, long long ldl, double* __restrict__ Li,
                      long long ldi) {
  for (int c = threadIdx.x; c < M; c += blockDim.x) {
    for (int i = 0; i < c; ++i) Li[(long long)i * ldi + c] = 0.0;
    for (int i = c; i < M; ++i) {
      const double* Lrow = L + (long long)i * ldl;
      double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = c;
      for (; k + 3 < i; k += 4) {
        s0 -= Lrow[k] * Li[(long long)k * ldi + c];
        s1 -= Lrow[k + 1] * Li[(long long)(k + 1) * ldi + c];
        s2 -= Lrow[k + 2] * Li[(long long)(k + 2) * ldi + c];
        s3 -= Lrow[k + 3] * Li[(long long)(k + 3) * ldi + c];
      }
      for (; k < i; ++k) s0 -= Lrow[k] * Li[(long long)k * ldi + c];
      Li[(long long)i * ldi + c] = ((s0 + s1) + (s2 + s3)) / Lrow[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Blocked single-matrix Cholesky + triangular inverse in ONE CTA (M <= 400): block size 32,
// left-looking.  Per block column J: (1) panel update A[J:,J] -= L[J:,:J] L[J,:J]^T with the 32 rows
// of L[J,:J] staged transposed in shared memory, (2) unblocked factorisation + inversion of the 32x32
// diagonal block in shared memory, (3) panel solve against the inverted diagonal block.
// Then Li = L^-1 block row by block row:  Li[I,J] = -Li[I,I] sum_K L[I,K] Li[K,J].
// The long dot products run on all 1024 threads; only the 32x32 diagonal work is latency bound
// (the unblocked kernels above spend ~0.7 ms + 1.6 ms at M = 200; this one ~0.1 ms).
// ---------------------------------------------------------------------------------------------
constexpr int CNB = 32;

__global__ void __launch_bounds__(1024)
chol_inv_blocked_kernel(double* __restrict__ A, int M, long long lda, double* __restrict__ Li,
                        long long ldi, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* Dg = sm;                    // CNB x (CNB+1)  diagonal block -> its Cholesky factor
  double* Di = Dg + CNB * (CNB + 1);  // CNB x (CNB+1)  inverse of that factor
  double* Bt = Di + CNB * (CNB + 1);  // M x (CNB+1)    staged rows (transposed) / T block
  double* Tb = Bt + (size_t)M * (CNB + 1);  // CNB x (M+1)  T = L[I,:] Li[:, :]
  const int tid = threadIdx.x, nthr = blockDim.x;
  int bad = 0;

  for (int J0 = 0; J0 < M; J0 += CNB) {
    const int nb = min(CNB, M - J0);
    // stage L[J0:J0+nb, 0:J0] transposed: Bt[k][j]
    for (int e = tid; e < nb * J0; e += nthr) {
      const int j = e / J0, k = e - j * J0;
      Bt[k * (CNB + 1) + j] = A[(long long)(J0 + j) * lda + k];
    }
    __syncthreads();
    // (1) panel update for rows i >= J0, columns j in the block (lower part only)
    for (int o = tid; o < (M - J0) * nb; o += nthr) {
      const int i = J0 + o / nb, j = o % nb;
      if (J0 + j > i) continue;
      const double* Lrow = A + (long long)i * lda;
      double s0 = 0.0, s1 = 0.0;
      int k = 0;
      for (; k + 1 < J0; k += 2) {
        s0 += Lrow[k] * Bt[k * (CNB + 1) + j];
        s1 += Lrow[k + 1] * Bt[(k + 1) * (CNB + 1) + j];
      }
      if (k < J0) s0 += Lrow[k] * Bt[k * (CNB + 1) + j];
      A[(long long)i * lda + J0 + j] -= (s0 + s1);
    }
    __syncthreads();
    // (2) diagonal block -> smem, unblocked Cholesky, then its inverse
    for (int e = tid; e < nb * nb; e += nthr) {
      const int i = e / nb, j = e % nb;
      Dg[i * (CNB + 1) + j] = (j <= i) ? A[(long long)(J0 + i) * lda + J0 + j] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
      const double piv = Dg[j * (CNB + 1) + j];
      if (!(piv > 0.0)) { bad = J0 + j + 1; break; }
      const double dj = sqrt(piv), inv = 1.0 / dj;
      __syncthreads();
      if (tid > j && tid < nb) Dg[tid * (CNB + 1) + j] *= inv;
      if (tid == 0) Dg[j * (CNB + 1) + j] = dj;
      __syncthreads();
      for (int o = tid; o < (nb - j - 1) * (nb - j - 1); o += nthr) {
        const int i = j + 1 + o / (nb - j - 1), k = j + 1 + o % (nb - j - 1);
        if (k <= i) Dg[i * (CNB + 1) + k] -= Dg[i * (CNB + 1) + j] * Dg[k * (CNB + 1) + j];
      }
      __syncthreads();
    }
    if (bad) break;
    if (tid < nb) {  // Di = Dg^-1, thread c owns column c
      const int c = tid;
      for (int i = 0; i < nb; ++i) {
        double s = (i == c) ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) s -= Dg[i * (CNB + 1) + k] * Di[k * (CNB + 1) + c];
        Di[i * (CNB + 1) + c] = (i < c) ? 0.0 : s / Dg[i * (CNB + 1) + i];
      }
    }
    __syncthreads();
    // write the factored diagonal block; (3) panel solve L[i, J] = A[i, J] Dg^-T for rows below
    for (int e = tid; e < nb * nb; e += nthr) {
      const int i = e / nb, j = e % nb;
      A[(long long)(J0 + i) * lda + J0 + j] = Dg[i * (CNB + 1) + j];
    }
    for (int o = tid; o < (M - J0 - nb) * nb; o += nthr) {
      const int i = J0 + nb + o / nb, j = o % nb;
      // X = A_row * Dg^-T  ->  X[j] = sum_{k<=j} A_row[k] * Di[j][k]
      const double* arow = A + (long long)i * lda + J0;
      double s = 0.0;
      for (int k = 0; k <= j; ++k) s += arow[k] * Di[j * (CNB + 1) + k];
      Tb[(o / nb) * (CNB + 1) + j] = s;  // (M - J0 - nb) x nb staging (fits: rows <= M)
    }
    __syncthreads();
    for (int o = tid; o < (M - J0 - nb) * nb; o += nthr) {
      const int i = J0 + nb + o / nb, j = o % nb;
      A[(long long)i * lda + J0 + j] = Tb[(o / nb) * (CNB + 1) + j];
    }
    // stash the inverse of the diagonal block into Li now (needed by the inversion sweep)
    for (int e = tid; e < nb * nb; e += nthr) {
      const int i = e / nb, j = e % nb;
      Li[(long long)(J0 + i) * ldi + J0 + j] = Di[i * (CNB + 1) + j];
    }
    __syncthreads();
  }
  // zero the strict upper triangle of L and of Li (block-upper part)
  for (long long e = tid; e < (long long)M * M; e += nthr) {
    const int i = (int)(e / M), j = (int)(e % M);
    if (j > i) A[(long long)i * lda + j] = 0.0;
    if ((j / CNB) > (i / CNB)) Li[(long long)i * ldi + j] = 0.0;
  }
  if (tid == 0) info[0] = bad;
  if (bad) return;
  __syncthreads();
  // inversion sweep, block row I (I0 > 0):  T = L[I, 0:I0] Li[0:I0, 0:I0] ; Li[I, 0:I0] = -Li[I,I] T
  for (int I0 = CNB; I0 < M; I0 += CNB) {
    const int nb = min(CNB, M - I0);
    for (int e = tid; e < nb * I0; e += nthr) {  // stage L[I rows, 0:I0] : Bt[i][k] (pitch M+1)
      const int i = e / I0, k = e - i * I0;
      Tb[i * (M + 1) + k] = A[(long long)(I0 + i) * lda + k];
    }
    for (int e = tid; e < nb * nb; e += nthr) {
      const int i = e / nb, j = e % nb;
      Di[i * (CNB + 1) + j] = Li[(long long)(I0 + i) * ldi + I0 + j];
    }
    __syncthreads();
    for (int o = tid; o < nb * I0; o += nthr) {
      const int i = o / I0, c = o - i * I0;
      const double* Lrow = Tb + i * (M + 1);
      double s0 = 0.0, s1 = 0.0;
      int k = c;
      for (; k + 1 < I0; k += 2) {
        s0 += Lrow[k] * Li[(long long)k * ldi + c];
        s1 += Lrow[k + 1] * Li[(long long)(k + 1) * ldi + c];
      }
      if (k < I0) s0 += Lrow[k] * Li[(long long)k * ldi + c];
      Bt[c * (CNB + 1) + i] = s0 + s1;  // T[i][c] stored as Bt[c][i]
    }
    __syncthreads();
    for (int o = tid; o < nb * I0; o += nthr) {
      const int i = o / I0, c = o - i * I0;
      double s = 0.0;
      for (int q = 0; q <= i; ++q) s += Di[i * (CNB + 1) + q] * Bt[c * (CNB + 1) + q];
      Li[(long long)(I0 + i) * ldi + c] = -s;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// masks / transposes / axpby / reductions
// ---------------------------------------------------------------------------------------------
// mode 0: keep lower triangle (incl. diagonal); mode 1: Phi = lower triangle with halved diagonal
__global__ void tril_kernel(double* __restrict__ X, int M, long long ld, int mode) {
  const long long total = (long long)M * M;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    double* p = X + (long long)i * ld + j;
    if (j > i) *p = 0.0;
    else if (mode == 1 && i == j) *p *= 0.5;
  }
}

__global__ void transpose_kernel(const double* __restrict__ X, int rows, int cols, long long ldx,
                                 double* __restrict__ Y, long long ldy) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < rows && j < cols) ? X[(long long)i * ldx + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = bx + r, i = by + threadIdx.x;  // Y[j][i] = X[i][j]
    if (i < rows && j < cols) Y[(long long)j * ldy + i] = tile[threadIdx.x][r];
  }
}

__global__ void axpby_kernel(long long n, double a, const double* __restrict__ x, double b,
                             double* __restrict__ y) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x)
    y[e] = a * x[e] + (b == 0.0 ? 0.0 : b * y[e]);
}

// Deterministic single-CTA reduction: out[0] = (accumulate ? out[0] : 0) + scale * sum_i x[i*stride]
__global__ void __launch_bounds__(1024)
sum_kernel(const double* __restrict__ x, long long n, long long stride, double scale,
           double* __restrict__ out, int accumulate) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += x[i * stride];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.0) + scale * s;
}

// sum of squares over a pitched 2-D array (used for |q_mu|^2): per-CTA partials (fixed grid ->
// deterministic), folded by sum_kernel.
constexpr int kSumsqBlocks = 296;
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const double* __restrict__ x, int rows, int cols, long long ld,
                     double* __restrict__ part) {
  __shared__ double red[32];
  double s = 0.0;
  const long long total = (long long)rows * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const double v = x[(e / cols) * ld + (e % cols)];
    s += v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// constrained = softplus(unconstrained) ; g_unconstrained = g_constrained * sigmoid(unconstrained)
__global__ void softplus_fwd_kernel(const double* u, double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = softplus_d(u[i]);
}
__global__ void softplus_bwd_kernel(const double* u, const double* g_c, double* g_u, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g_u[i] = g_c[i] * sigmoid_d(u[i]);
}

// ---------------------------------------------------------------------------------------------
// Keras-flavoured Adam (tf.optimizers.Adam; reference fcest/helpers/inference.py:72-74):
//   m += (g - m)(1-b1);  v += (g^2 - v)(1-b2);  theta -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps)
// `sign` = -1 turns an ELBO gradient into a loss gradient.
// ---------------------------------------------------------------------------------------------
// alpha_dev != nullptr: the bias-corrected step size is read from device memory (CUDA-graph replays keep the
// same kernel arguments while the step counter advances)
__global__ void adam_kernel(double* __restrict__ theta, const double* __restrict__ grad,
                            double* __restrict__ m, double* __restrict__ v, long long n, double sign,
                            double alpha_t, double b1, double b2, double eps,
                            const double* __restrict__ alpha_dev = nullptr) {
  if (alpha_dev != nullptr) alpha_t = alpha_dev[0];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    double th = theta[e], mi = m[e], vi = v[e];
    adam_update(th, mi, vi, sign * grad[e], alpha_t, b1, b2, eps);
    m[e] = mi;
    v[e] = vi;
    theta[e] = th;
  }
}

// ---------------------------------------------------------------------------------------------
// Counter-based N(0,1) generator: Philox4x32-10 + Box-Muller on 53-bit uniforms.
// element e uses counter (e/2, offset) and key (seed): reproducible and order independent.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                             uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__global__ void randn_philox_kernel(double* __restrict__ out, long long n, unsigned long long seed,
                                    unsigned long long offset) {
  const long long pairs = (n + 1) / 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < pairs;
       q += (long long)gridDim.x * blockDim.x) {
    uint32_t c0 = (uint32_t)q, c1 = (uint32_t)((unsigned long long)q >> 32);
    uint32_t c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c0, c1, c2, c3, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const unsigned long long a = ((unsigned long long)c0 << 32) | c1;
    const unsigned long long b = ((unsigned long long)c2 << 32) | c3;
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    out[2 * q] = rad * cs;
    if (2 * q + 1 < n) out[2 * q + 1] = rad * sn;
  }
}

static inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  if (g > 148LL * 32) g = 148LL * 32;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace fcest

using namespace fcest;

static long long g_launch_count = 0;
extern "C" long long fcest_launch_count_add(long long n) { g_launch_count += n; return g_launch_count; }
extern "C" long long fcest_launch_count(void) { return g_launch_count; }

extern "C" int fcest_matern52(const double* X1, int n1, const double* X2, int n2, int dim,
                              const double* kvar, const double* kls, double jitter, double* K,
                              long long ldk, cudaStream_t stream) {
  if (n1 <= 0 || n2 <= 0) return 0;
  matern52_kernel<<<grid_for((long long)n1 * n2, 256), 256, 0, stream>>>(X1, n1, X2, n2, dim, kvar,
                                                                         kls, jitter, K, ldk);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_matern52_bwd(const double* X1, int n1, const double* X2, int n2, int dim,
                                  const double* kvar, const double* kls, const double* gK,
                                  long long ldg, int sym, double* ws_part, double* g_var,
                                  double* g_ls, double* gX1, int accumulate, cudaStream_t stream) {
  if (n1 <= 0 || n2 <= 0) return 0;
  if (dim > 4 && gX1) return FCEST_ERR_UNSUPPORTED;
  matern52_bwd_kernel<<<n1, 256, 0, stream>>>(X1, n1, X2, n2, dim, kvar, kls, gK, ldg, sym, ws_part,
                                              gX1);
  FCEST_CHECK_LAUNCH();
  sum_kernel<<<1, 1024, 0, stream>>>(ws_part, n1, 2, 1.0, g_var, accumulate);
  FCEST_CHECK_LAUNCH();
  sum_kernel<<<1, 1024, 0, stream>>>(ws_part + 1, n1, 2, 1.0, g_ls, accumulate);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_chol_single(double* A, int M, long long lda, int* info, cudaStream_t stream) {
  chol_single_kernel<<<1, 1024, 0, stream>>>(A, M, lda, info);
  FCEST_CHECK_LAUNCH();
  return 0;
}

namespace fcest {
bool chol_inv_tiles_supported(int M);
int launch_chol_inv_tiles(double* A, int M, long long lda, double* Li, long long ldi, int* info, cudaStream_t stream);
}

extern "C" int fcest_chol_inv_single(double* A, int M, long long lda, double* Li, long long ldi, int* info,
                                     cudaStream_t stream) {
  if (chol_inv_tiles_supported(M)) {  // tensor-pipe tile kernel (chol_single_tiles.cu)
    const int rc = launch_chol_inv_tiles(A, M, lda, Li, ldi, info, stream);
    if (rc != 0) return rc;
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  if (M > 400) {  // staging buffers no longer fit one CTA's shared memory: unblocked path
    chol_single_kernel<<<1, 1024, 0, stream>>>(A, M, lda, info);
    FCEST_CHECK_LAUNCH();
    tri_inv_single_kernel<<<1, 1024, 0, stream>>>(A, M, lda, Li, ldi);
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  const size_t smem = sizeof(double) * (2 * CNB * (CNB + 1) + (size_t)M * (CNB + 1) + (size_t)CNB * (M + 1));
  cudaError_t e = cudaFuncSetAttribute(chol_inv_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  chol_inv_blocked_kernel<<<1, 1024, smem, stream>>>(A, M, lda, Li, ldi, info);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_tri_inv_single(const double* L, int M, long long ldl, double* Li, long long ldi,
                                    cudaStream_t stream) {
  tri_inv_single_kernel<<<1, 1024, 0, stream>>>(L, M, ldl, Li, ldi);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_tril(double* X, int M, long long ld, int mode, cudaStream_t stream) {
  tril_kernel<<<grid_for((long long)M * M, 256), 256, 0, stream>>>(X, M, ld, mode);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_transpose(const double* X, int rows, int cols, long long ldx, double* Y,
                               long long ldy, cudaStream_t stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, stream>>>(X, rows, cols, ldx, Y, ldy);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_axpby(long long n, double a, const double* x, double b, double* y,
                           cudaStream_t stream) {
  if (n <= 0) return 0;
  axpby_kernel<<<grid_for(n, 256), 256, 0, stream>>>(n, a, x, b, y);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_sum(const double* x, long long n, long long stride, double scale, double* out,
                         int accumulate, cudaStream_t stream) {
  sum_kernel<<<1, 1024, 0, stream>>>(x, n, stride, scale, out, accumulate);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_sumsq(const double* x, int rows, int cols, long long ld, double scale,
                           double* out, int accumulate, double* ws_296, cudaStream_t stream) {
  sumsq_partial_kernel<<<kSumsqBlocks, 256, 0, stream>>>(x, rows, cols, ld, ws_296);
  FCEST_CHECK_LAUNCH();
  sum_kernel<<<1, 1024, 0, stream>>>(ws_296, kSumsqBlocks, 1, scale, out, accumulate);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_softplus_fwd(const double* u, double* out, int n, cudaStream_t stream) {
  softplus_fwd_kernel<<<(n + 127) / 128, 128, 0, stream>>>(u, out, n);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_softplus_bwd(const double* u, const double* g_c, double* g_u, int n,
                                  cudaStream_t stream) {
  softplus_bwd_kernel<<<(n + 127) / 128, 128, 0, stream>>>(u, g_c, g_u, n);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_adam_step(double* theta, const double* grad, double* m, double* v, long long n,
                               double grad_sign, long long t, double lr, double b1, double b2,
                               double eps, cudaStream_t stream) {
  if (n <= 0) return 0;
  const double alpha_t = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
  adam_kernel<<<grid_for(n, 256), 256, 0, stream>>>(theta, grad, m, v, n, grad_sign, alpha_t, b1, b2,
                                                    eps);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_adam_step_dev(double* theta, const double* grad, double* m, double* v, long long n,
                                   double grad_sign, const double* alpha_t, double b1, double b2, double eps,
                                   cudaStream_t stream) {
  if (n <= 0) return 0;
  if (alpha_t == nullptr) return FCEST_ERR_ARGUMENT;
  adam_kernel<<<grid_for(n, 256), 256, 0, stream>>>(theta, grad, m, v, n, grad_sign, 0.0, b1, b2, eps, alpha_t);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_randn(double* out, long long n, unsigned long long seed,
                           unsigned long long offset, cudaStream_t stream) {
  if (n <= 0) return 0;
  randn_philox_kernel<<<grid_for((n + 1) / 2, 256), 256, 0, stream>>>(out, n, seed, offset);
  FCEST_CHECK_LAUNCH();
  return 0;
}
