// dp_capi.cu -- extern "C" entry points of the batched DP planner (include/cilqr_b200.h, CilqrDp*).
// Compiled with -fmad=false (see dp_kernel.cuh); linked into libcilqr_b200.so.
#include "dp_kernel.cuh"

#include <cstdio>
#include <cstring>
#include <vector>

#include "cilqr_internal.h"

namespace {

#define CKH(call)                                                   \
  do {                                                              \
    cudaError_t e_ = (call);                                        \
    if (e_ != cudaSuccess) return cilqr_internal_fail(h, e_, #call); \
  } while (0)

int validate(const cilqr_handle* h, const CilqrDpConfig* cfg, const CilqrDpIn* in, const CilqrDpOut* out) {
  if (!h || !cfg || !in || !out) return CILQR_E_INVALID;
  if (in->B < 0 || in->R < 2 || in->NB < 0 || in->V < 1 || in->n_static < 0 || in->n_dyn < 0 || in->T < 0)
    return CILQR_E_INVALID;
  if (!(cfg->tf > 0) || !(cfg->delta_t > 0)) return CILQR_E_INVALID;
  if (in->B == 0) return CILQR_OK;
  if (!in->ref || !in->start || !out->ok || (in->NB > 0 && !in->barrier)) return CILQR_E_INVALID;
  if (in->n_static > 0 && (!in->static_poly || !in->static_nv)) return CILQR_E_INVALID;
  if (in->n_dyn > 0 && (!in->dyn_time || !in->dyn_samples || !in->dyn_poly || !in->dyn_nv)) return CILQR_E_INVALID;
  return CILQR_OK;
}

// host_barrier: the barrier points in host memory (the host path has them; the device path copies them back once
// per call -- 16 bytes per point)
int launch(cilqr_handle* h, const CilqrDpConfig* cfg, const CilqrDpIn* in, const CilqrDpOut* out, double ref_s0,
           double ref_s1, const double* host_barrier, cudaStream_t st) {
  dp::Args a;
  memset(&a, 0, sizeof(a));
  dp::make_lattice(cfg->tf, cfg->delta_t, cfg->max_velocity, cfg->width, cfg->wheel_base, cfg->front_hang_length,
                   cfg->rear_hang_length, &a.lat);
  if (a.lat.K < 2 || a.lat.K > dp::kMaxKnots) return CILQR_E_CAPACITY;
  for (int k = 0; k < dp::NT; ++k)
    if (a.lat.nseg[k] < 1) return CILQR_E_INVALID;
  a.B = in->B; a.R = in->R; a.NB = in->NB; a.V = in->V; a.n_static = in->n_static; a.n_dyn = in->n_dyn; a.T = in->T;
  a.tf = cfg->tf; a.delta_t = cfg->delta_t; a.nominal_velocity = cfg->dp_nominal_velocity;
  a.w_obstacle = cfg->dp_w_obstacle; a.w_lateral = cfg->dp_w_lateral; a.w_lateral_change = cfg->dp_w_lateral_change;
  a.w_lateral_velocity_change = cfg->dp_w_lateral_velocity_change;
  a.w_longitudinal_velocity_bias = cfg->dp_w_longitudinal_velocity_bias;
  a.w_longitudinal_velocity_change = cfg->dp_w_longitudinal_velocity_change;
  a.wheel_base = cfg->wheel_base;
  a.ref_s0 = ref_s0;
  a.ref_inv_ds = ref_s1 > ref_s0 ? (double)(in->R - 1) / (ref_s1 - ref_s0) : 0.0;
  a.ref = in->ref; a.barrier = in->barrier; a.start = in->start; a.static_poly = in->static_poly;
  a.static_nv = in->static_nv; a.dyn_time = in->dyn_time; a.dyn_samples = in->dyn_samples; a.dyn_poly = in->dyn_poly;
  a.dyn_nv = in->dyn_nv;
  a.trajectory = out->trajectory; a.coarse = out->coarse; a.xytheta = out->xytheta; a.ok = out->ok; a.cost = out->cost;
  a.waypoints = out->waypoints;
  // the barrier grid: built on the host, uploaded to the handle's scratch (slot 0)
  std::vector<int> gs, gi;
  std::vector<double> gxy;
  dp::build_grid(host_barrier, in->NB, a.lat.radius, &a, &gs, &gi, &gxy);
  {
    char* g = nullptr;
    const size_t b0 = (gs.size() * sizeof(int) + 255) & ~(size_t)255, b1 = (gi.size() * sizeof(int) + 255) & ~(size_t)255;
    const size_t b2 = gxy.size() * sizeof(double);
    int rcg = cilqr_internal_scratch(h, 0, b0 + b1 + b2, &g);
    if (rcg != CILQR_OK) return rcg;
    CKH(cudaMemcpyAsync(g, gs.data(), gs.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CKH(cudaMemcpyAsync(g + b0, gi.data(), gi.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CKH(cudaMemcpyAsync(g + b0 + b1, gxy.data(), b2, cudaMemcpyHostToDevice, st));
    CKH(cudaStreamSynchronize(st));  // gs / gi / gxy are locals
    a.grid_start = (const int*)g;
    a.grid_idx = (const int*)(g + b0);
    a.grid_xy = (const double*)(g + b0 + b1);
  }
  // the per-sample bounds of the dynamic obstacles go to shared memory while two CTAs per SM still fit (~110 KB each)
  a.use_sample_bounds = dp::smem_bytes(a.lat.K, in->n_static + in->n_dyn, in->n_dyn, in->T, true) <= 110 * 1024 ? 1 : 0;
  const size_t smem = dp::smem_bytes(a.lat.K, in->n_static + in->n_dyn, in->n_dyn, in->T, a.use_sample_bounds != 0);
  int per_sm = 1;
  CKH(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp::dp_plan_kernel, dp::kMaxThreads, smem));
  if (per_sm < 1) per_sm = 1;
  long long blocks = (long long)cilqr_internal_num_sms(h) * per_sm;  // persistent: CTAs stride over the scenarios
  if (blocks > in->B) blocks = in->B;
  cudaEvent_t e0, e1;
  int rc = cilqr_internal_events(h, 1, &e0, &e1);
  if (rc != CILQR_OK) return rc;
  CKH(cudaEventRecord(e0, st));
  dp::dp_plan_kernel<<<(unsigned)blocks, dp::kMaxThreads, smem, st>>>(a);
  CKH(cudaGetLastError());
  CKH(cudaEventRecord(e1, st));
  return CILQR_OK;
}

}  // namespace

extern "C" {

// called once from cilqr_create: function attributes are per-device state shared by every handle
int cilqr_internal_dp_set_smem(int optin_bytes) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, (const void*)dp::dp_plan_kernel) != cudaSuccess) return CILQR_E_CUDA;
  return cudaFuncSetAttribute((const void*)dp::dp_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              optin_bytes - (int)fa.sharedSizeBytes) == cudaSuccess
             ? CILQR_OK
             : CILQR_E_CUDA;
}

void cilqr_dp_default_config(CilqrDpConfig* c) {
  if (!c) return;
  c->tf = 8.0;  // planner_config.h:94-99
  c->delta_t = 0.1;
  c->dp_nominal_velocity = 10.0;
  c->dp_w_obstacle = 1000.0;
  c->dp_w_lateral = 0.1;
  c->dp_w_lateral_change = 0.5;
  c->dp_w_lateral_velocity_change = 1.0;
  c->dp_w_longitudinal_velocity_bias = 10.0;
  c->dp_w_longitudinal_velocity_change = 1.0;
  c->max_velocity = 20.0;  // vehicle_param.h:26-46
  c->width = 1.942;
  c->wheel_base = 1.0;
  c->front_hang_length = 0.96;
  c->rear_hang_length = 0.929;
}

int cilqr_dp_num_knots(const CilqrDpConfig* c) {
  if (!c || !(c->tf > 0) || !(c->delta_t > 0)) return CILQR_E_INVALID;
  dp::Lattice L;
  dp::make_lattice(c->tf, c->delta_t, c->max_velocity, c->width, c->wheel_base, c->front_hang_length, c->rear_hang_length,
                   &L);
  return L.K;
}

int cilqr_dp_plan_batch_device(cilqr_handle* h, const CilqrDpConfig* cfg, const CilqrDpIn* in, const CilqrDpOut* out,
                               void* cuda_stream) {
  int rc = validate(h, cfg, in, out);
  if (rc != CILQR_OK || in->B == 0) return rc;
  CKH(cudaSetDevice(cilqr_internal_device(h)));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : cilqr_internal_stream(h);
  // first / last station of the centre line (for the station-search guess) and the barrier (for its grid)
  double ends[2];
  std::vector<double> hb((size_t)(in->NB > 0 ? in->NB : 1) * 2);
  CKH(cudaMemcpyAsync(&ends[0], in->ref, sizeof(double), cudaMemcpyDeviceToHost, st));
  CKH(cudaMemcpyAsync(&ends[1], in->ref + (size_t)(in->R - 1) * 7, sizeof(double), cudaMemcpyDeviceToHost, st));
  if (in->NB > 0) CKH(cudaMemcpyAsync(hb.data(), in->barrier, (size_t)in->NB * 16, cudaMemcpyDeviceToHost, st));
  CKH(cudaStreamSynchronize(st));
  return launch(h, cfg, in, out, ends[0], ends[1], hb.data(), st);
}

int cilqr_dp_plan_batch(cilqr_handle* h, const CilqrDpConfig* cfg, const CilqrDpIn* in, const CilqrDpOut* out) {
  int rc = validate(h, cfg, in, out);
  if (rc != CILQR_OK || in->B == 0) return rc;
  CKH(cudaSetDevice(cilqr_internal_device(h)));
  const int K = cilqr_dp_num_knots(cfg);
  if (K < 2) return CILQR_E_INVALID;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t B = in->B;
  const size_t n[16] = {up((size_t)in->R * 7 * 8), up((size_t)in->NB * 2 * 8), up(B * 3 * 8),
                        up(B * in->n_static * in->V * 2 * 8), up(B * in->n_static * 4), up(B * in->n_dyn * in->T * 8),
                        up(B * in->n_dyn * 4), up(B * in->n_dyn * in->T * in->V * 2 * 8), up(B * in->n_dyn * 4),
                        out->trajectory ? up(B * K * 13 * 8) : 0, out->coarse ? up(B * K * 6 * 8) : 0,
                        out->xytheta ? up(B * K * 3 * 8) : 0, up(B * 4), out->cost ? up(B * 8) : 0,
                        out->waypoints ? up(B * dp::NT * 3 * 8) : 0, 0};
  size_t total = 0;
  for (size_t x : n) total += x;
  char* base = nullptr;
  rc = cilqr_internal_scratch(h, 1, total, &base);
  if (rc != CILQR_OK) return rc;
  char* p[16];
  {
    char* q = base;
    for (int i = 0; i < 16; ++i) {
      p[i] = q;
      q += n[i];
    }
  }
  cudaStream_t st = cilqr_internal_stream(h);
  const void* src[9] = {in->ref, in->barrier, in->start, in->static_poly, in->static_nv, in->dyn_time,
                        in->dyn_samples, in->dyn_poly, in->dyn_nv};
  const size_t raw[9] = {(size_t)in->R * 7 * 8, (size_t)in->NB * 2 * 8, B * 3 * 8, B * in->n_static * in->V * 2 * 8,
                         B * in->n_static * 4, B * in->n_dyn * in->T * 8, B * in->n_dyn * 4,
                         B * in->n_dyn * in->T * in->V * 2 * 8, B * in->n_dyn * 4};
  for (int i = 0; i < 9; ++i)
    if (raw[i] && src[i]) CKH(cudaMemcpyAsync(p[i], src[i], raw[i], cudaMemcpyHostToDevice, st));
  CilqrDpIn din = *in;
  din.ref = (const double*)p[0]; din.barrier = (const double*)p[1]; din.start = (const double*)p[2];
  din.static_poly = (const double*)p[3]; din.static_nv = (const int32_t*)p[4]; din.dyn_time = (const double*)p[5];
  din.dyn_samples = (const int32_t*)p[6]; din.dyn_poly = (const double*)p[7]; din.dyn_nv = (const int32_t*)p[8];
  CilqrDpOut dout;
  dout.trajectory = out->trajectory ? (double*)p[9] : nullptr;
  dout.coarse = out->coarse ? (double*)p[10] : nullptr;
  dout.xytheta = out->xytheta ? (double*)p[11] : nullptr;
  dout.ok = (int32_t*)p[12];
  dout.cost = out->cost ? (double*)p[13] : nullptr;
  dout.waypoints = out->waypoints ? (double*)p[14] : nullptr;
  rc = launch(h, cfg, &din, &dout, in->ref[0], in->ref[(size_t)(in->R - 1) * 7], in->barrier, st);
  if (rc != CILQR_OK) return rc;
  if (out->trajectory) CKH(cudaMemcpyAsync(out->trajectory, dout.trajectory, B * K * 13 * 8, cudaMemcpyDeviceToHost, st));
  if (out->coarse) CKH(cudaMemcpyAsync(out->coarse, dout.coarse, B * K * 6 * 8, cudaMemcpyDeviceToHost, st));
  if (out->xytheta) CKH(cudaMemcpyAsync(out->xytheta, dout.xytheta, B * K * 3 * 8, cudaMemcpyDeviceToHost, st));
  CKH(cudaMemcpyAsync(out->ok, dout.ok, B * 4, cudaMemcpyDeviceToHost, st));
  if (out->cost) CKH(cudaMemcpyAsync(out->cost, dout.cost, B * 8, cudaMemcpyDeviceToHost, st));
  if (out->waypoints) CKH(cudaMemcpyAsync(out->waypoints, dout.waypoints, B * dp::NT * 3 * 8, cudaMemcpyDeviceToHost, st));
  CKH(cudaStreamSynchronize(st));
  return CILQR_OK;
}

int cilqr_dp_last_kernel_ms(cilqr_handle* h, float* ms) {
  if (!h || !ms) return CILQR_E_INVALID;
  cudaEvent_t e0, e1;
  int rc = cilqr_internal_events(h, 1, &e0, &e1);
  if (rc != CILQR_OK) return rc;
  CKH(cudaSetDevice(cilqr_internal_device(h)));
  CKH(cudaEventSynchronize(e1));
  CKH(cudaEventElapsedTime(ms, e0, e1));
  return CILQR_OK;
}

}  // extern "C"
