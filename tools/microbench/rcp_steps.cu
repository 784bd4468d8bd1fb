// How many Newton steps does the barrier's reciprocal need?  fast_rcp (cilqr_kernel.cuh) refines the hardware seed
// rcp.approx.ftz.f64 (MUFU.RCP64H) with r <- r + r (1 - g r).  This measures the error in ulps against the IEEE
// division for 1, 2 and 3 steps over log-uniform |g| in [1e-12, 1e12], both signs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rcp_steps tools/microbench/rcp_steps.cu && /tmp/rcp_steps
#include <cstdio>
#include <cstdint>
#include <cmath>
__device__ __forceinline__ double seed(double g) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(g)); return r; }
__device__ __forceinline__ double step(double g, double r) { const double e = fma(-g, r, 1.0); return fma(r, e, r); }
__device__ unsigned long long rng(unsigned long long& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
__global__ void k(unsigned long long* worst, double* rel0, int per_thread) {
  unsigned long long s = 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long w[4] = {0, 0, 0, 0};
  double r0max = 0.0;
  for (int i = 0; i < per_thread; ++i) {
    const double u = (rng(s) >> 11) * (1.0 / 9007199254740992.0);
    const double mant = 1.0 + (rng(s) >> 12) * (1.0 / 4503599627370496.0);
    double g = mant * exp2(floor((u - 0.5) * 80.0));
    if (rng(s) & 1) g = -g;
    const double ex = 1.0 / g;
    double r = seed(g);
    r0max = fmax(r0max, fabs(r * g - 1.0));
    for (int n = 1; n <= 3; ++n) {
      r = step(g, r);
      const long long d = __double_as_longlong(r) - __double_as_longlong(ex);
      const unsigned long long a = d < 0 ? -d : d;
      if (a > w[n]) w[n] = a;
    }
  }
  for (int n = 1; n <= 3; ++n) atomicMax(worst + n, w[n]);
  atomicMax((unsigned long long*)rel0, (unsigned long long)__double_as_longlong(r0max));
}
int main() {
  unsigned long long* d; double* r0;
  cudaMalloc(&d, 32); cudaMemset(d, 0, 32); cudaMalloc(&r0, 8); cudaMemset(r0, 0, 8);
  k<<<592, 256>>>(d, r0, 4096);
  unsigned long long h[4]; double hr;
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost); cudaMemcpy(&hr, r0, 8, cudaMemcpyDeviceToHost);
  printf("{\"samples\": %lld, \"seed_max_rel_err\": %.3e, \"max_ulp_error_vs_ieee_division\": {\"1_step\": %llu, \"2_steps\": %llu, \"3_steps\": %llu}}\n",
         592ll * 256 * 4096, hr, h[1], h[2], h[3]);
  return cudaGetLastError() != cudaSuccess;
}
