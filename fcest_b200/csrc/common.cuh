// Shared device helpers for the fcest_b200 sm_100a kernels (fp64 everywhere).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Called once after every kernel launch: bumps the library's launch counter (bench.py reports it as
// "gpu_launches") and surfaces launch errors as the C-ABI return code.
extern "C" long long fcest_launch_count_add(long long n);
#define FCEST_CHECK_LAUNCH()                                   \
  do {                                                         \
    fcest_launch_count_add(1);                                 \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) return (int)e__;                   \
  } while (0)

namespace fcest {

// Keras-flavoured Adam on one element (tf.optimizers.Adam as run_adam applies it): alpha_t = lr sqrt(1 - b2^t) /
// (1 - b1^t), epsilon outside the square root.  One definition for the stand-alone kernel and for the update fused
// into tri_grad's epilogue, so that both produce the same bits.
__device__ __forceinline__ void adam_update(double& theta, double& m, double& v, double g, double alpha_t, double b1,
                                            double b2, double eps) {
  const double mi = m + (g - m) * (1.0 - b1);
  const double vi = v + (g * g - v) * (1.0 - b2);
  m = mi;
  v = vi;
  theta -= alpha_t * mi / (sqrt(vi) + eps);
}


// fp64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// Fragment ownership (lane = 4*gid + tig): a = A[gid][tig], b = B[tig][gid],
// c0/c1 = C[gid][2*tig], C[gid][2*tig+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte async copy global->shared; bytes beyond `src_bytes` (0..16) are zero-filled.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; `red` is >= 32 doubles of shared memory. Result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double t = (lane < nw) ? red[lane] : 0.0;
  t = warp_sum(t);
  return t;
}

__device__ __forceinline__ double softplus_d(double u) {
  // log(1 + exp(u)), stable on both tails (matches tf.math.softplus to rounding)
  return u > 0 ? u + log1p(exp(-u)) : log1p(exp(u));
}
__device__ __forceinline__ double sigmoid_d(double u) {
  return u >= 0 ? 1.0 / (1.0 + exp(-u)) : exp(u) / (1.0 + exp(u));
}

// packed lower-triangular index (row-major): (i,j), i >= j  ->  i(i+1)/2 + j
__host__ __device__ __forceinline__ long long tri_index(int i, int j) {
  return (long long)i * (i + 1) / 2 + j;
}

}  // namespace fcest
