// Single-matrix Cholesky + triangular inverse on the fp64 tensor pipe (M <= 207): the M x M kernel matrix of
// the whitened conditional (tf.linalg.cholesky / triangular_solve inside gpflow's base_conditional, reached
// from fcest/models/wishart_process.py:109-114,381-386).  One CTA, 16 warps; the lower triangle lives as packed
// swizzled 8 x 8 tiles in shared memory (lik_tiles.cuh): left-looking blocked Cholesky (DMMA updates, register /
// shuffle diagonal tiles), then L^-1 block row by block row in place.  ~10x faster than the scalar blocked
// kernel in small.cu, which stays as the fallback for 208 <= M <= 400.
#include <math.h>

#include "common.cuh"
#include "fcest_b200.h"
#include "lik_tiles.cuh"

namespace fcest {

constexpr int CS_WARPS = 16;

__device__ __forceinline__ void cs_decode(int t, int& ib, int& jb) {
  ib = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  if ((ib + 1) * (ib + 2) / 2 <= t) ++ib;
  if (ib * (ib + 1) / 2 > t) --ib;
  jb = t - ib * (ib + 1) / 2;
}

__global__ void __launch_bounds__(CS_WARPS * 32, 1)
chol_inv_tiles_kernel(double* __restrict__ A, int M, long long lda, double* __restrict__ Li, long long ldi,
                      int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int NB = (M + 7) / 8, NT = NB * (NB + 1) / 2;
  double* tiles = sm;                 // NT x 64 (swizzled)
  double* ldiag = tiles + NT * 64;    // NB x 64: the diagonal blocks of L (row-major)
  double* dinv = ldiag + NB * 64;     // 8 NB
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int oA0 = swz(gid, tig), oA1 = swz(gid, 4 + tig);
  const int oT0 = swz(tig, gid), oT1 = swz(4 + tig, gid);
  const int oC = swz(gid, 2 * tig);

  // 1. lower triangle -> tiles (identity padding past M)
  for (int e = tid; e < NT * 64; e += blockDim.x) {
    int ib, jb;
    cs_decode(e >> 6, ib, jb);
    const int r = (e >> 3) & 7, c = e & 7;
    const int i = 8 * ib + r, j = 8 * jb + c;
    double v = 0.0;
    if (i < M && j <= i) v = A[(long long)i * lda + j];
    else if (i >= M && j == i) v = 1.0;
    tiles[(e >> 6) * 64 + swz(r, c)] = v;
  }
  __syncthreads();

  // 2. left-looking blocked Cholesky; diagonal tile -> I_jj (its inverse), the factor block itself -> ldiag
  int bad = 0;
  for (int jb = 0; jb < NB; ++jb) {
    for (int ib = jb + warp; ib < NB; ib += CS_WARPS) {
      double2 c = *reinterpret_cast<const double2*>(tiles + TI(ib, jb) * 64 + oC);
      double e0 = 0.0, e1 = 0.0;  // second accumulator chain (odd kb)
      int kb = 0;
      for (; kb + 1 < jb; kb += 2) {
        const double* Tj = tiles + TI(jb, kb) * 64;
        const double* Ti = tiles + TI(ib, kb) * 64;
        const double* Tj2 = tiles + TI(jb, kb + 1) * 64;
        const double* Ti2 = tiles + TI(ib, kb + 1) * 64;
        dmma(c.x, c.y, Ti[oA0], -Tj[oA0]);
        dmma(e0, e1, Ti2[oA0], -Tj2[oA0]);
        dmma(c.x, c.y, Ti[oA1], -Tj[oA1]);
        dmma(e0, e1, Ti2[oA1], -Tj2[oA1]);
      }
      if (kb < jb) {
        const double* Tj = tiles + TI(jb, kb) * 64;
        const double* Ti = tiles + TI(ib, kb) * 64;
        dmma(c.x, c.y, Ti[oA0], -Tj[oA0]);
        dmma(c.x, c.y, Ti[oA1], -Tj[oA1]);
      }
      *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(c.x + e0, c.y + e1);
    }
    __syncthreads();
    if (warp == 0) bad = factor_diag_swz(tiles + TI(jb, jb) * 64, dinv + 8 * jb, lane, -1, 8 * jb, M, bad, ldiag + jb * 64);
    __syncthreads();
    {
      const double* Ij = tiles + TI(jb, jb) * 64;
      const double b0 = Ij[oA0], b1 = Ij[oA1];
      for (int ib = jb + 1 + warp; ib < NB; ib += CS_WARPS) {
        const double* X = tiles + TI(ib, jb) * 64;
        double r0 = 0.0, r1 = 0.0;
        const double x0 = X[oA0], x1 = X[oA1];
        dmma(r0, r1, x0, b0);
        dmma(r0, r1, x1, b1);
        __syncwarp();
        *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(r0, r1);
      }
    }
    __syncthreads();
  }
  if (tid == 0) info[0] = bad;

  // 3. write L (dense lower, strict upper zero) before the tiles are overwritten by L^-1
  for (int e = tid; e < M * M; e += blockDim.x) {
    const int i = e / M, j = e - i * M;
    double v = 0.0;
    if (j <= i) {
      const int ib = i >> 3, jb = j >> 3;
      v = (ib == jb) ? ldiag[ib * 64 + (i & 7) * 8 + (j & 7)] : tiles[TI(ib, jb) * 64 + swz(i & 7, j & 7)];
    }
    A[(long long)i * lda + j] = v;
  }
  __syncthreads();

  // 4. Li = L^-1 block row by block row, in place (diagonal tiles already hold I_jj)
  for (int ib = 1; ib < NB; ++ib) {
    double x[2][2];
    int jbs[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      jbs[t] = warp + t * CS_WARPS;
      x[t][0] = x[t][1] = 0.0;
      if (jbs[t] < ib) {
        for (int kb = jbs[t]; kb < ib; ++kb) {
          const double* Lk = tiles + TI(ib, kb) * 64;
          const double* Xk = tiles + TI(kb, jbs[t]) * 64;
          dmma(x[t][0], x[t][1], Lk[oA0], Xk[oT0]);
          dmma(x[t][0], x[t][1], Lk[oA1], Xk[oT1]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (jbs[t] < ib) *reinterpret_cast<double2*>(tiles + TI(ib, jbs[t]) * 64 + oC) = make_double2(x[t][0], x[t][1]);
    __syncthreads();
    const double i0 = tiles[TI(ib, ib) * 64 + oA0], i1 = tiles[TI(ib, ib) * 64 + oA1];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (jbs[t] < ib) {
        const double* X = tiles + TI(ib, jbs[t]) * 64;
        double r0 = 0.0, r1 = 0.0;
        const double c0 = X[oT0], c1 = X[oT1];
        dmma(r0, r1, i0, c0);
        dmma(r0, r1, i1, c1);
        __syncwarp();
        *reinterpret_cast<double2*>(tiles + TI(ib, jbs[t]) * 64 + oC) = make_double2(-r0, -r1);
      }
    }
    __syncthreads();
  }
  // 5. write Li (dense lower, strict upper zero)
  for (int e = tid; e < M * M; e += blockDim.x) {
    const int i = e / M, j = e - i * M;
    Li[(long long)i * ldi + j] = (j <= i) ? tiles[TI(i >> 3, j >> 3) * 64 + swz(i & 7, j & 7)] : 0.0;
  }
}

bool chol_inv_tiles_supported(int M) {
  const int NB = (M + 7) / 8;
  return M >= 1 && NB <= 26 && NB <= 2 * CS_WARPS;
}

int launch_chol_inv_tiles(double* A, int M, long long lda, double* Li, long long ldi, int* info, cudaStream_t stream) {
  const int NB = (M + 7) / 8, NT = NB * (NB + 1) / 2;
  const size_t smem = sizeof(double) * ((size_t)NT * 64 + (size_t)NB * 64 + 8 * (size_t)NB);
  cudaError_t e = cudaFuncSetAttribute(chol_inv_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  chol_inv_tiles_kernel<<<1, CS_WARPS * 32, smem, stream>>>(A, M, lda, Li, ldi, info);
  return 0;
}

}  // namespace fcest
