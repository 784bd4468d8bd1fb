// Development aid: compiles the DEVICE code of cilqr_b200/csrc/corridor_kernel.cuh for the host (one
// "thread" per call, shared-memory stride 1) so that the build logic can be checked against the
// oracle in a container without a GPU.  Not part of the product, the tests or the bench.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o /tmp/libcorr_emul.so tools/corridor_host_emul.cc
#define CORRIDOR_HOST_EMUL
#include <cmath>
#include <cstddef>
#include <cstdint>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__
struct emul_dim3 { int x; };
static emul_dim3 threadIdx{0}, blockIdx{0}, blockDim{1}, gridDim{1};
struct double2 { double x, y; };
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline float __double2float_rn(double a) { return (float)a; }
namespace corridor { unsigned char corr_smem[1 << 16]; }
#include "../cilqr_b200/csrc/corridor_kernel.cuh"

extern "C" void emul_corridor(int B, int K, int P_max, int M_max, int cap, const double* traj,
                              const double* pts, const int* cnt, double* corridor, int* ccnt, double* poly,
                              int* code) {
  corridor::Args a;
  a.B = B; a.K = K; a.P_max = P_max; a.M_max = M_max; a.cap = cap;
  a.max_diff_x = 25; a.max_diff_y = 25; a.radius = 150; a.max_axis_x = 10; a.max_axis_y = 10;
  a.traj = traj; a.obs_points = pts; a.obs_cnt = cnt; a.corridor = corridor; a.corridor_cnt = ccnt;
  a.polygon = poly; a.code = code;
  corridor::corridor_build_kernel(a);
}
