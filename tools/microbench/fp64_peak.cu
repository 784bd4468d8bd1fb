// fp64_peak.cu -- measured non-tensor FP64 peak of the device: independent DFMA chains, all SMs, 8 warps/scheduler.
// Prints one JSON line {"dfma_tflops": ...}.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll 8
    for (int j = 0; j < 8; ++j) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 4096;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 1.0 + r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
  printf("{\"dfma_tflops\": %.3f, \"sms\": %d, \"kernel_ms\": %.4f, \"how\": \"8 independent DFMA chains per thread, 2 CTAs x 1024 threads per SM, best of 5\"}\n",
         flops / (best * 1e-3) / 1e12, p.multiProcessorCount, best);
  return 0;
}
