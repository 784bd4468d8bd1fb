// Instruction-cache capacity probe (development tool; results in profiles/).
// W warps per SM loop over S bytes of straight-line code, unaligned (staggered start delays).
// Prints instructions per clock per SM for each (S, W): the knee in S is the capacity of the last
// instruction-cache level that can feed several unaligned instruction streams.
#include <cstdio>
#include <cuda_runtime.h>

#define I4(a, b, c, d)                                              \
  asm volatile("fma.rn.f32 %0, %0, %4, %5;\n\tfma.rn.f32 %1, %1, %4, %5;\n\t" \
               "fma.rn.f32 %2, %2, %4, %5;\n\tfma.rn.f32 %3, %3, %4, %5;"     \
               : "+f"(a), "+f"(b), "+f"(c), "+f"(d) : "f"(m), "f"(n));
#define I16 I4(a0, a1, a2, a3) I4(a4, a5, a6, a7) I4(a0, a1, a2, a3) I4(a4, a5, a6, a7)
#define I64 I16 I16 I16 I16
#define I256 I64 I64 I64 I64  // 4 KB of code

template <int NB>
struct Body {
  static __device__ __forceinline__ void run(float& a0, float& a1, float& a2, float& a3, float& a4, float& a5,
                                             float& a6, float& a7, float m, float n) {
    I256 Body<NB - 1>::run(a0, a1, a2, a3, a4, a5, a6, a7, m, n);
  }
};
template <>
struct Body<0> {
  static __device__ __forceinline__ void run(float&, float&, float&, float&, float&, float&, float&, float&, float,
                                             float) {}
};

template <int NB>
__global__ void probe(float* out, int iters, float m, float n) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  // stagger: each warp waits a different number of clocks so the instruction streams are unaligned
  const int warp = (threadIdx.x >> 5) + blockIdx.x * (blockDim.x >> 5);
  const long long t0 = clock64();
  while (clock64() - t0 < (long long)(warp % 37) * 977) {}
#pragma unroll 1
  for (int it = 0; it < iters; ++it) Body<NB>::run(a0, a1, a2, a3, a4, a5, a6, a7, m, n);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int NB>
void run(int sms, float* out, float clk_ghz) {
  for (int W : {4, 8, 12, 16}) {
    const int total_instr = 1 << 22;  // per warp
    const int iters = total_instr / (NB * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    // one warp per CTA so that warps are independent streams; W CTAs per SM
    probe<NB><<<sms * W, 32>>>(out, 4, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<NB><<<sms * W, 32>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr_per_sm = (double)W * iters * NB * 256;
    printf("S=%4d KB  W=%2d  ms=%8.3f  ipc_per_sm=%.3f\n", NB * 4, W, ms, instr_per_sm / (ms * 1e-3 * clk_ghz * 1e9));
  }
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* out;
  cudaMalloc(&out, 148 * 16 * 32 * 4 * 4);
  const float ghz = p.clockRate * 1e-6f;
  printf("%s sms=%d clock=%.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<2>(p.multiProcessorCount, out, ghz);
  run<4>(p.multiProcessorCount, out, ghz);
  run<6>(p.multiProcessorCount, out, ghz);
  run<8>(p.multiProcessorCount, out, ghz);
  run<12>(p.multiProcessorCount, out, ghz);
  run<16>(p.multiProcessorCount, out, ghz);
  run<24>(p.multiProcessorCount, out, ghz);
  run<32>(p.multiProcessorCount, out, ghz);
  run<48>(p.multiProcessorCount, out, ghz);
  run<64>(p.multiProcessorCount, out, ghz);
  return 0;
}
