// Development aid: compiles the DEVICE code of cilqr_b200/csrc/dp_kernel.cuh for the host (one "thread" per
// CTA) so that the planner logic can be checked against the oracle in a container without a GPU.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o /tmp/libdp_emul.so tools/dp_host_emul.cc
#define DP_HOST_EMUL
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __grid_constant__
#define __noinline__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__
static inline void __syncthreads() {}
static inline int atomicAdd(int* p, int v) { const int o = *p; *p += v; return o; }
struct emul_dim3 { int x; };
static emul_dim3 threadIdx{0}, blockIdx{0}, blockDim{1}, gridDim{1};
namespace dp { unsigned char dp_smem[1 << 18]; }
#include "../cilqr_b200/csrc/dp_kernel.cuh"

// cfg: the 14 doubles of CilqrDpConfig; dims: B, R, NB, V, n_static, n_dyn, T
extern "C" int emul_dp(const double* cfg, const int* dims, const double* ref, const double* barrier, const double* start,
                       const double* static_poly, const int* static_nv, const double* dyn_time, const int* dyn_samples,
                       const double* dyn_poly, const int* dyn_nv, double* trajectory, double* coarse, double* xytheta,
                       int* ok, double* cost, double* waypoints) {
  dp::Args a;
  memset(&a, 0, sizeof(a));
  dp::make_lattice(cfg[0], cfg[1], cfg[9], cfg[10], cfg[11], cfg[12], cfg[13], &a.lat);
  a.B = dims[0]; a.R = dims[1]; a.NB = dims[2]; a.V = dims[3]; a.n_static = dims[4]; a.n_dyn = dims[5]; a.T = dims[6];
  a.tf = cfg[0]; a.delta_t = cfg[1]; a.nominal_velocity = cfg[2]; a.w_obstacle = cfg[3]; a.w_lateral = cfg[4];
  a.w_lateral_change = cfg[5]; a.w_lateral_velocity_change = cfg[6]; a.w_longitudinal_velocity_bias = cfg[7];
  a.w_longitudinal_velocity_change = cfg[8]; a.wheel_base = cfg[11];
  a.ref_s0 = ref[0];
  a.ref_inv_ds = (double)(a.R - 1) / (ref[(size_t)(a.R - 1) * 7] - ref[0]);
  a.ref = ref; a.barrier = barrier; a.start = start; a.static_poly = static_poly; a.static_nv = static_nv;
  a.dyn_time = dyn_time; a.dyn_samples = dyn_samples; a.dyn_poly = dyn_poly; a.dyn_nv = dyn_nv;
  a.trajectory = trajectory; a.coarse = coarse; a.xytheta = xytheta; a.ok = ok; a.cost = cost; a.waypoints = waypoints;
  std::vector<int> gs, gi;
  std::vector<double> gxy;
  dp::build_grid(barrier, a.NB, a.lat.radius, &a, &gs, &gi, &gxy);
  a.grid_start = gs.data();
  a.grid_idx = gi.data();
  a.grid_xy = gxy.data();
  a.use_sample_bounds = dp::smem_bytes(a.lat.K, a.n_static + a.n_dyn, a.n_dyn, a.T, true) <= sizeof(dp::dp_smem) ? 1 : 0;
  gridDim.x = 1;
  dp::dp_plan_kernel(a);
  return a.lat.K;
}
