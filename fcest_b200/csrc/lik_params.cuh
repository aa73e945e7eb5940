// Argument block shared by the Wishart likelihood kernels (likelihood.cu: CTA-per-time-step shared-memory /
// L2-scratch variants; likelihood_warp.cu: warp-per-matrix variant for D <= 55).
#pragma once
#include <cuda_runtime.h>

namespace fcest {

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
  double* ws; long long ws_stride;  // BIG variant: per-CTA global scratch for G, T, Yb
};

// Warp-per-matrix variant.  Returns 0 if launched, a cudaError_t (> 0) on failure, or -1 if the shape
// is outside what the variant supports (caller falls through to the CTA-per-time-step kernels).
// *part_rows receives the number of rows of gA_part / glam_part it filled (one per persistent CTA).
int launch_loglik_warp(const LikParams& p, cudaStream_t stream, int* part_rows);
bool loglik_warp_supported(int D, int nu);

// Sub-warp variant (likelihood_subwarp.cu), D <= 7: one 8-lane group per time step, registers + shuffles only.
int launch_loglik_subwarp(const LikParams& p, cudaStream_t stream, int* part_rows);
bool loglik_subwarp_supported(int D, int nu);

// Shared-memory-tiled CTA-per-matrix variant (likelihood_tiles.cu), 8 <= D <= 207.  p.ws must point to
// loglik_tiles_scratch_bytes(D, nu) bytes of global scratch.  Same return convention as above.
int launch_loglik_tiles(LikParams p, cudaStream_t stream, int* part_rows);
bool loglik_tiles_supported(int D, int nu);
long long loglik_tiles_scratch_bytes(int D, int nu);

// G = A o (fmean + sqrt(fvar) o eps) for all samples of the time steps [n0, n0 + nc): slab z = s nc + (n - n0) of Gc, D rows of
// pitch ldgs (likelihood_tiles.cu; shared by the tiles and the cluster-tiled kernels).  Returns 0 or a cudaError_t.
int launch_loglik_g_fill(const LikParams& p, int n0, int nc, int ldgs, double* Gc, cudaStream_t stream);

// Cluster-tiled variant (likelihood_cluster.cu), 8 <= D <= 407: one thread-block cluster per matrix, the factor's
// tiles in the distributed shared memory of the cluster.  p.ws must point to loglik_cluster_scratch_bytes(D, nu) bytes.
int launch_loglik_cluster(LikParams p, cudaStream_t stream, int* part_rows);
bool loglik_cluster_supported(int D, int nu);
long long loglik_cluster_scratch_bytes(int D, int nu);

}  // namespace fcest
