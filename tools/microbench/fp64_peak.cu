// Microbenchmark: fp64 DMMA (mma.sync m8n8k4) vs DFMA issue throughput on sm_100a.
// Used to establish the fp64 roofline denominator (MEASURED_PEAKS.json has no fp64 figure).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_loop(double* out, int iters, double seed) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = seed * i; c[i][1] = seed + i; }
  double a = seed * threadIdx.x, b = seed + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_loop(double* out, int iters, double seed) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = seed * i;
  double a = seed * threadIdx.x, b = seed + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s sms %d\n", p.name, sms);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int cps : {1, 2}) {
      int threads = warps * 32 / cps; if (threads < 32) continue;
      int grid = sms * cps;
      float ms = time_it([&] { dmma_loop<16><<<grid, threads>>>(out, iters, 1e-9); });
      double flops = 2.0 * 256 * 16 * (double)iters * (threads / 32) * grid;
      printf("DMMA ilp16 warps/SM %2d (ctas/SM %d): %.3f ms  %.2f TFLOP/s\n", warps, cps, ms, flops / ms * 1e-9);
    }
  }
  for (int warps : {4, 8}) {
    float ms = time_it([&] { dmma_loop<4><<<sms, warps * 32>>>(out, iters, 1e-9); });
    double flops = 2.0 * 256 * 4 * (double)iters * warps * sms;
    printf("DMMA ilp4  warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
    ms = time_it([&] { dmma_loop<32><<<sms, warps * 32>>>(out, iters / 2, 1e-9); });
    flops = 2.0 * 256 * 32 * (double)(iters / 2) * warps * sms;
    printf("DMMA ilp32 warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_it([&] { dfma_loop<16><<<sms, warps * 32>>>(out, iters, 1e-9); });
    double flops = 2.0 * 32 * 16 * (double)iters * warps * sms;
    printf("DFMA ilp16 warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  return 0;
}
