// Shared-memory tile helpers of the Wishart likelihood kernels (likelihood_warp.cu, likelihood_tiles.cu):
// swizzled 8x8 fp64 tiles, the register / shuffle Cholesky + inverse of a diagonal tile, L2 prefetch.
#pragma once
#include "common.cuh"

namespace fcest {

// element (r, c) of a swizzled 8x8 tile (64 doubles): conflict-free for the DMMA A-type ([gid][tig]),
// transposed ([tig][gid]) and accumulator ([gid][2 tig .. 2 tig + 1], 16-byte) access patterns.
__host__ __device__ __forceinline__ int swz(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }

// Cholesky + inverse of one swizzled 8x8 diagonal tile (rows in lanes, replicated over the four lane
// groups; registers + shuffles).  On exit the tile holds I = L_bb^-1 (lower, strict upper zeroed) and
// dinv8[i] = 1 / L_bb[i][i].  Row `yr` (if >= 0) is the bordered y row: its pivot is forced to 1.
static __device__ __noinline__ int factor_diag_swz(double* tile, double* dinv8, int lane, int yr, int rowbase,
                                               int D, int bad, double* ltile = nullptr) {
  const int i = lane & 7;
  const bool sw = (i & 2) != 0;
  double ph[8];
  {
    const double2* pr = reinterpret_cast<const double2*>(tile + i * 8);
#pragma unroll
    for (int c = 0; c < 4; ++c) { const double2 v = pr[c]; ph[2 * c] = v.x; ph[2 * c + 1] = v.y; }
  }
  double row[8], x[8], my_rinv = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double v = sw ? ph[c ^ 4] : ph[c];
    row[c] = (c <= i) ? v : 0.0;
    x[c] = (c == i) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double piv = __shfl_sync(0xffffffffu, row[j], j);
    if (j == yr) piv = 1.0;
    if (!(piv > 0.0) && bad == 0 && rowbase + j < D) bad = rowbase + j + 1;
    const double rinv = rsqrt(piv);
    if (i == j) my_rinv = rinv;
    row[j] = (i == j) ? piv * rinv : row[j] * rinv;
#pragma unroll
    for (int k = j + 1; k < 8; ++k) {
      const double lkj = __shfl_sync(0xffffffffu, row[j], k);
      if (i >= k) row[k] -= row[j] * lkj;
    }
  }
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
  if (lane < 8 && ltile != nullptr) {  // optional copy of the factor itself (row-major 8 x 8, strict upper zeroed)
#pragma unroll
    for (int c = 0; c < 8; ++c) ltile[i * 8 + c] = (c <= i) ? row[c] : 0.0;
  }
  if (lane < 8) {
    double out[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] = (c <= i) ? x[c] : 0.0;
    double2* pw = reinterpret_cast<double2*>(tile + i * 8);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double v0 = sw ? out[(2 * c) ^ 4] : out[2 * c];
      const double v1 = sw ? out[(2 * c + 1) ^ 4] : out[2 * c + 1];
      pw[c] = make_double2(v0, v1);
    }
    dinv8[i] = my_rinv;
  }
  return bad;
}

// L2 prefetch of `bytes` starting at p (128-byte lines dealt to the lanes of one warp)
__device__ __forceinline__ void prefetch_l2_warp(const void* p, long long bytes, int lane) {
  const char* c = reinterpret_cast<const char*>(p);
  for (long long off = (long long)lane * 128; off < bytes; off += 32 * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}

// packed index of lower tile (ib, jb), ib >= jb
#define TI(ib, jb) ((ib) * ((ib) + 1) / 2 + (jb))

}  // namespace fcest
