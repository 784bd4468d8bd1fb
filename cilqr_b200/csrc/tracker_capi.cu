// tracker_capi.cu -- extern "C" entry points of the batched Tracker initial guess (include/cilqr_b200.h,
// cilqr_tracker_*).  Compiled with -fmad=false (see tracker_kernel.cuh).
#include "tracker_kernel.cuh"

#include <string.h>

#include <algorithm>

#include "../../include/cilqr_b200.h"
#include "cilqr_internal.h"

#define CKT(call)                                                        \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) return cilqr_internal_fail(h, e_, #call);     \
  } while (0)

extern "C" {

void cilqr_tracker_default_config(CilqrTrackerConfig* c) {
  if (!c) return;
  c->sumulation_dt = 0.01;  // planner_config.h:36-43
  c->dt = 0.1;
  c->tolerance = 0.01;
  c->lat_weight_l = 1e-1;  // :18-25
  c->lat_weight_theta = 1e-12;
  c->lat_weight_delta = 1e-12;
  c->lat_weight_delta_rate = 0.1;
  c->lat_preview_time = 0.2;
  c->lon_weight_s = 5.0 * 1e-1;  // :27-34
  c->lon_weight_v = 1e-12;
  c->lon_weight_a = 1e-12;
  c->lon_weight_j = 0.1;
  c->wheel_base = 1.0;  // vehicle_param.h:26-64
  c->delta_min = -40.0 / 180 * M_PI;
  c->delta_max = 40.0 / 180 * M_PI;
  c->min_acceleration = -5.0;
  c->max_acceleration = 5.0;
  c->delta_rate_min = c->delta_min / 3.0;
  c->delta_rate_max = c->delta_max / 3.0;
  c->jerk_min = -10.0;
  c->jerk_max = 10.0;
  c->max_num_iteration = 150;
}

// called once from cilqr_create (function attributes are per-device state shared by every handle)
int cilqr_internal_tracker_set_smem(int optin_bytes) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, (const void*)trk::tracker_kernel) != cudaSuccess) return CILQR_E_CUDA;
  return cudaFuncSetAttribute((const void*)trk::tracker_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              optin_bytes - (int)fa.sharedSizeBytes) == cudaSuccess
             ? CILQR_OK
             : CILQR_E_CUDA;
}

int cilqr_tracker_batch_device(cilqr_handle* h, const CilqrTrackerConfig* cfg, int B, int K, const double* start,
                               const double* coarse_traj, double* traj, double* guess_states, double* guess_controls,
                               int32_t* ok, void* cuda_stream) {
  if (!h || !cfg || B < 0 || K < 2) return CILQR_E_INVALID;
  if (!(cfg->sumulation_dt > 0.0) || cfg->max_num_iteration < 0) return CILQR_E_INVALID;
  if (B == 0) return CILQR_OK;
  if (!start || !coarse_traj || !ok) return CILQR_E_INVALID;
  CKT(cudaSetDevice(cilqr_internal_device(h)));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : cilqr_internal_stream(h);
  trk::Args a;
  memset(&a, 0, sizeof(a));
  static_assert(sizeof(trk::Config) == sizeof(CilqrTrackerConfig), "tracker config mirrors");
  memcpy(&a.c, cfg, sizeof(a.c));
  a.B = B;
  a.K = K;
  a.start = start;
  a.coarse = coarse_traj;
  a.traj = traj;
  a.guess_states = guess_states;
  a.guess_controls = guess_controls;
  a.ok = ok;
  // threads per CTA: as many as the per-thread x / y columns allow in the SM's shared memory, at most 128
  const size_t per_thread = trk::smem_bytes_per_thread(K);
  int threads = (int)((size_t)(cilqr_internal_smem_optin(h) - 2048) / per_thread) / 32 * 32;
  threads = std::min(threads, 128);
  if (threads < 32) return CILQR_E_SMEM;
  const size_t smem = per_thread * threads;
  long long blocks = ((long long)B + threads - 1) / threads;
  blocks = std::min<long long>(blocks, (long long)cilqr_internal_num_sms(h) * 4);
  cudaEvent_t e0, e1;
  int rc = cilqr_internal_events(h, 2, &e0, &e1);
  if (rc != CILQR_OK) return rc;
  CKT(cudaEventRecord(e0, st));
  trk::tracker_kernel<<<(unsigned)blocks, threads, smem, st>>>(a);
  CKT(cudaGetLastError());
  CKT(cudaEventRecord(e1, st));
  return CILQR_OK;
}

int cilqr_tracker_batch(cilqr_handle* h, const CilqrTrackerConfig* cfg, int B, int K, const double* start,
                        const double* coarse_traj, double* traj, double* guess_states, double* guess_controls,
                        int32_t* ok) {
  if (!h || !cfg || B < 0 || K < 2) return CILQR_E_INVALID;
  if (B == 0) return CILQR_OK;
  if (!start || !coarse_traj || !ok) return CILQR_E_INVALID;
  CKT(cudaSetDevice(cilqr_internal_device(h)));
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_start = up((size_t)B * 4 * 8), b_co = up((size_t)B * K * 13 * 8), b_tr = traj ? b_co : 0;
  const size_t b_gx = guess_states ? up((size_t)B * K * 6 * 8) : 0, b_gu = guess_controls ? up((size_t)B * (K - 1) * 2 * 8) : 0;
  const size_t b_ok = up((size_t)B * 4);
  char* p = nullptr;
  int rc = cilqr_internal_scratch(h, 2, b_start + b_co + b_tr + b_gx + b_gu + b_ok, &p);
  if (rc != CILQR_OK) return rc;
  double* d_start = (double*)p; p += b_start;
  double* d_co = (double*)p; p += b_co;
  double* d_tr = traj ? (double*)p : nullptr; p += b_tr;
  double* d_gx = guess_states ? (double*)p : nullptr; p += b_gx;
  double* d_gu = guess_controls ? (double*)p : nullptr; p += b_gu;
  int32_t* d_ok = (int32_t*)p;
  cudaStream_t st = cilqr_internal_stream(h);
  CKT(cudaMemcpyAsync(d_start, start, (size_t)B * 4 * 8, cudaMemcpyHostToDevice, st));
  CKT(cudaMemcpyAsync(d_co, coarse_traj, (size_t)B * K * 13 * 8, cudaMemcpyHostToDevice, st));
  rc = cilqr_tracker_batch_device(h, cfg, B, K, d_start, d_co, d_tr, d_gx, d_gu, d_ok, st);
  if (rc != CILQR_OK) return rc;
  if (traj) CKT(cudaMemcpyAsync(traj, d_tr, (size_t)B * K * 13 * 8, cudaMemcpyDeviceToHost, st));
  if (guess_states) CKT(cudaMemcpyAsync(guess_states, d_gx, (size_t)B * K * 6 * 8, cudaMemcpyDeviceToHost, st));
  if (guess_controls) CKT(cudaMemcpyAsync(guess_controls, d_gu, (size_t)B * (K - 1) * 2 * 8, cudaMemcpyDeviceToHost, st));
  CKT(cudaMemcpyAsync(ok, d_ok, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  CKT(cudaStreamSynchronize(st));
  return CILQR_OK;
}

int cilqr_tracker_last_kernel_ms(cilqr_handle* h, float* ms) {
  if (!h || !ms) return CILQR_E_INVALID;
  cudaEvent_t e0, e1;
  int rc = cilqr_internal_events(h, 2, &e0, &e1);
  if (rc != CILQR_OK) return rc;
  CKT(cudaEventSynchronize(e1));
  CKT(cudaEventElapsedTime(ms, e0, e1));
  return CILQR_OK;
}

}  // extern "C"
