// Shared-memory tile helpers of the Wishart likelihood kernels (likelihood_warp.cu, likelihood_tiles.cu):
// swizzled 8x8 fp64 tiles, the register / shuffle Cholesky + inverse of a diagonal tile, L2 prefetch.
#pragma once
#include "common.cuh"

namespace fcest {

// element (r, c) of a swizzled 8x8 tile (64 doubles): conflict-free for the DMMA A-type ([gid][tig]),
// transposed ([tig][gid]) and accumulator ([gid][2 tig .. 2 tig + 1], 16-byte) access patterns.
__host__ __device__ __forceinline__ int swz(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }

// Cholesky + inverse of one swizzled 8x8 diagonal tile, all in registers + shuffles.  Lane (i = lane & 7,
// g = lane >> 3) holds the two entries [i][2g], [i][2g+1] of the tile and of the running inverse X = L^-1.
// Step j of the right-looking factorisation broadcasts the unscaled column j (pivot, own-row entry, the
// two column-partner entries) with four shuffles issued BEFORE 1/sqrt(pivot) is known, so the dependent
// chain per step is  shuffle -> rsqrt -> scale -> fused update;  the forward substitution for X is merged
// into the same step (row j of X is final once column j of L is), off the critical path.
// On exit the tile holds I = L_bb^-1 (lower, strict upper zeroed), dinv8[i] = 1 / L_bb[i][i] and, if asked,
// ltile the factor block itself (row-major 8 x 8).  Row `yr` (if >= 0) is the bordered y row: unit pivot.
static __device__ __noinline__ int factor_diag_swz(double* tile, double* dinv8, int lane, int yr, int rowbase,
                                               int D, int bad, double* ltile = nullptr) {
  const int i = lane & 7, g = lane >> 3;
  const int pidx = g ^ (i & 2);  // physical 16-byte pair of row i that holds logical columns 2g, 2g+1
  double r[2], x[2], my_rinv = 0.0;
  {
    const double2 v = reinterpret_cast<const double2*>(tile + i * 8)[pidx];
    r[0] = (2 * g <= i) ? v.x : 0.0;
    r[1] = (2 * g + 1 <= i) ? v.y : 0.0;
    x[0] = (2 * g == i) ? 1.0 : 0.0;
    x[1] = (2 * g + 1 == i) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gj = j >> 1, ej = j & 1, base = 8 * gj;
    const double raw_p = __shfl_sync(0xffffffffu, r[ej], base + j);
    const double raw_i = __shfl_sync(0xffffffffu, r[ej], base + i);
    const double raw_k0 = __shfl_sync(0xffffffffu, r[ej], base + 2 * g);
    const double raw_k1 = __shfl_sync(0xffffffffu, r[ej], base + 2 * g + 1);
    const double piv = (j == yr) ? 1.0 : raw_p;
    if (!(piv > 0.0) && bad == 0 && rowbase + j < D) bad = rowbase + j + 1;
    const double rinv = rsqrt(piv);
    const double lij = raw_i * rinv;  // L[i][j] for i > j
    if (i == j) my_rinv = rinv;
    // trailing update of this lane's two columns k = 2g + e > j (lower part: i >= k)
    if (2 * g > j && i >= 2 * g) r[0] -= lij * (raw_k0 * rinv);
    if (2 * g + 1 > j && i >= 2 * g + 1) r[1] -= lij * (raw_k1 * rinv);
    // column j itself becomes L[:, j]
    if (g == gj) r[ej] = (i == j) ? piv * rinv : (i > j ? lij : r[ej]);
    // forward substitution for X = L^-1: row j is final after scaling; eliminate it from the rows below
    if (i == j) { x[0] *= rinv; x[1] *= rinv; }
    const double xj0 = __shfl_sync(0xffffffffu, x[0], 8 * g + j);
    const double xj1 = __shfl_sync(0xffffffffu, x[1], 8 * g + j);
    if (i > j) { x[0] -= lij * xj0; x[1] -= lij * xj1; }
  }
  if (ltile != nullptr) {  // optional copy of the factor itself (row-major 8 x 8, strict upper zeroed)
    ltile[i * 8 + 2 * g] = (2 * g <= i) ? r[0] : 0.0;
    ltile[i * 8 + 2 * g + 1] = (2 * g + 1 <= i) ? r[1] : 0.0;
  }
  reinterpret_cast<double2*>(tile + i * 8)[pidx] = make_double2((2 * g <= i) ? x[0] : 0.0, (2 * g + 1 <= i) ? x[1] : 0.0);
  if (lane < 8) dinv8[i] = my_rinv;
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
