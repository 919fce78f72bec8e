// fp64 FMA throughput microbenchmark (developer tool, not product code)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  k<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s); k<<<blocks, threads>>>(out, iters); cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double fmas = (double)blocks * threads * iters * 8;
  printf("%s SMs=%d clock=%d kHz: %.2f TFLOP/s fp64 (FMA=2), %.1f DFMA/clk/SM at nominal clock\n", p.name, p.multiProcessorCount,
         p.clockRate, 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
  return 0;
}
