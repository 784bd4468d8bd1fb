// Development aid: compiles the DEVICE code of cilqr_b200/csrc/tracker_kernel.cuh for the host (one "thread" per CTA)
// so that the tracker logic can be checked against the oracle in a container without a GPU.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o /tmp/libtracker_emul.so tools/tracker_host_emul.cc
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __align__(x)
#define __shared__
struct emul_dim3 { int x; };
static emul_dim3 threadIdx{0}, blockIdx{0}, blockDim{1}, gridDim{1};
namespace trk { double trk_smem[1 << 14]; }
#define TRACKER_HOST_EMUL
#include "../cilqr_b200/csrc/tracker_kernel.cuh"

// cfg: the 21 doubles of CilqrTrackerConfig followed by max_num_iteration
extern "C" void emul_tracker(const double* cfg, int max_iter, int B, int K, const double* start, const double* coarse,
                             double* traj, double* gx, double* gu, int* ok) {
  trk::Args a;
  memset(&a, 0, sizeof(a));
  memcpy(&a.c, cfg, sizeof(double) * 21);
  a.c.max_num_iteration = max_iter;
  a.B = B; a.K = K; a.start = start; a.coarse = coarse; a.traj = traj; a.guess_states = gx; a.guess_controls = gu; a.ok = ok;
  gridDim.x = 1;
  trk::tracker_kernel(a);
}
