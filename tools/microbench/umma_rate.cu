// Issue-to-completion rate of tcgen05.mma for several kinds / shapes / smem layouts (operands: zeros in shared memory).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t addr, int sbo, int layout) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int KIND>  // 0 = i8, 1 = f8f6f4, 2 = f16
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else if (KIND == 1)
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int sbo, int layout, int kstep_bytes, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  if (threadIdx.x == 0) {
    // c_format: i8 -> S32 (2), others F32 (1); a/b format: i8 -> 1 (signed), f8 -> 0 (e4m3), f16 -> 1 (bf16)
    uint32_t idesc = ((KIND == 0 ? 2u : 1u) << 4) | ((KIND == 1 ? 0u : 1u) << 7) | ((KIND == 1 ? 0u : 1u) << 10) |
                     ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = desc(base, sbo, layout), db = desc(base + 32768, sbo, layout);
    const int ksteps = 4;  // cycle through 4 K offsets of the tile like a real main loop
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i % ksteps;
      const int off = (k * kstep_bytes) >> 4;
      mma<KIND>(tmem + (uint32_t)((i & 1) * 256), da + off, db + off, idesc, 1u);
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND>
void run(const char* name, int N, int sbo, int layout, int kstep_bytes, int grid) {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 4096, smem = 100 * 1024;
  cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<KIND><<<grid, 128, smem>>>(N, sbo, layout, kstep_bytes, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s grid %3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", name, grid, (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<0>("i8  M128 N128 K32  SW64  (sbo 512)", 128, 512, 4, 32, grid);
    run<0>("i8  M128 N128 K32  SW128 (sbo 1024)", 128, 1024, 2, 32, grid);
    run<0>("i8  M128 N256 K32  SW128", 256, 1024, 2, 32, grid);
    run<0>("i8  M128 N64  K32  SW128", 64, 1024, 2, 32, grid);
    run<1>("e4m3 M128 N128 K32 SW128", 128, 1024, 2, 32, grid);
    run<1>("e4m3 M128 N256 K32 SW128", 256, 1024, 2, 32, grid);
    run<2>("bf16 M128 N128 K16 SW128", 128, 1024, 2, 32, grid);
    run<2>("bf16 M128 N256 K16 SW128", 256, 1024, 2, 32, grid);
  }
  return 0;
}
