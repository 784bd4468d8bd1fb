// multi_capi.cu -- the multi-GPU entry points of include/cilqr_b200.h (cilqr_multi_*, cilqr_plan_sharded).
//
// The reference plans one trajectory per call on one CPU thread (planning_node.cc:82-88); a batch of independent
// scenarios shards trivially: device r of G solves the contiguous ids [r * per, (r + 1) * per), per = ceil(B / G)
// (SURVEY 8(e)).  There is no exchange inside the solve; the ONE collective is an all-gather of the per-shard
// result blocks [states | controls | status] over NVLink, for consumers that need every result on every GPU.
// One host process drives all GPUs (the C++ host of the reference is a single process): one solver handle and one
// host thread per device, ncclCommInitAll for the communicators.  NCCL is resolved at run time (dlopen of
// libnccl.so.2), so the library itself does not depend on it; without NCCL the sharded solve still works and only a
// request for the gathered copy fails (CILQR_E_NCCL).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cilqr_b200.h"
#include "cilqr_internal.h"

namespace {

// the few NCCL entry points used, with the types of nccl.h (2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;      // ncclSuccess == 0
typedef int ncclDataType_t;    // ncclFloat64 == 8
constexpr ncclDataType_t kNcclFloat64 = 8;
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return CommInitAll && CommDestroy && GroupStart && GroupEnd && AllGather;
  }
};

struct Shard {
  int device = 0;
  cilqr_handle* h = nullptr;
  cudaStream_t stream = nullptr;
  char* in_buf = nullptr;
  size_t in_bytes = 0;
  double* block = nullptr;  // [states | controls | status] of `per` scenarios
  size_t block_bytes = 0;
  double* extra = nullptr;  // the optional `result` records
  size_t extra_bytes = 0;
  int rc = CILQR_OK;
};

}  // namespace

struct cilqr_multi {
  int G = 0;
  std::vector<Shard> sh;
  NcclApi nccl;
  std::vector<ncclComm_t> comms;
  std::string err;
};

namespace {

size_t up256(size_t x) { return (x + 255) / 256 * 256; }

int grow(void** p, size_t* have, size_t need) {
  if (*have >= need) return CILQR_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  if (cudaMalloc(p, need) != cudaSuccess) return CILQR_E_CUDA;
  *have = need;
  return CILQR_OK;
}

// one shard: inputs host -> device, solve, (later) results device -> host
int run_shard(Shard* s, const CilqrBatchIn* in, const CilqrBatchOut* out, int b0, int nb, int per) {
  if (cudaSetDevice(s->device) != cudaSuccess) return CILQR_E_CUDA;
  const size_t K = in->N + 1, N = in->N, M = in->M_max;
  const size_t b_start = 4 * 8, b_coarse = K * 6 * 8, b_corr = K * M * 3 * 8, b_cnt = K * 4;
  const size_t b_ll = (size_t)in->S_left * 7 * 8, b_lr = (size_t)in->S_right * 7 * 8;
  const size_t need = up256(b_start * per) + up256(b_coarse * per) + up256(b_corr * per) + up256(b_cnt * per) +
                      up256(b_ll * per) + up256(b_lr * per);
  int rc = grow((void**)&s->in_buf, &s->in_bytes, need);
  if (rc != CILQR_OK) return rc;
  const size_t blk = ((size_t)per * (K * 6 + N * 2 + CILQR_STATUS_DOUBLES)) * sizeof(double);
  rc = grow((void**)&s->block, &s->block_bytes, blk);
  if (rc != CILQR_OK) return rc;
  if (out->result) {
    rc = grow((void**)&s->extra, &s->extra_bytes, (size_t)per * K * CILQR_TRAJPOINT_DOUBLES * sizeof(double));
    if (rc != CILQR_OK) return rc;
  }
  if (cudaMemsetAsync(s->block, 0, blk, s->stream) != cudaSuccess) return CILQR_E_CUDA;
  if (nb <= 0) return CILQR_OK;
  char* p = s->in_buf;
  auto put = [&](const void* host, size_t per_b) -> char* {
    char* d = p;
    p += up256(per_b * per);
    if (cudaMemcpyAsync(d, (const char*)host + per_b * b0, per_b * nb, cudaMemcpyHostToDevice, s->stream) != cudaSuccess)
      rc = CILQR_E_CUDA;
    return d;
  };
  CilqrBatchIn din = *in;
  din.B = nb;
  din.start = (const double*)put(in->start, b_start);
  din.coarse = (const double*)put(in->coarse, b_coarse);
  din.corridor = (const double*)put(in->corridor, b_corr);
  din.corridor_cnt = (const int32_t*)put(in->corridor_cnt, b_cnt);
  din.lane_left = (const double*)put(in->lane_left, b_ll);
  din.lane_right = (const double*)put(in->lane_right, b_lr);
  if (rc != CILQR_OK) return rc;
  CilqrBatchOut dout;
  memset(&dout, 0, sizeof(dout));
  dout.states = s->block;
  dout.controls = s->block + (size_t)per * K * 6;
  dout.status = dout.controls + (size_t)per * N * 2;
  dout.result = out->result ? s->extra : nullptr;
  return cilqr_plan_batch_device(s->h, &din, &dout, s->stream);
}

}  // namespace

extern "C" {

int cilqr_multi_create(const CilqrParams* params, int n_devices, const int* devices, int N_max, int M_max, int S_max,
                       int B_max_per_device, cilqr_multi** out) {
  if (!out) return CILQR_E_INVALID;
  *out = nullptr;
  if (!params || n_devices < 1) return CILQR_E_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
    cudaGetLastError();
    return CILQR_E_NO_DEVICE;
  }
  if (n_devices > count) return CILQR_E_NO_DEVICE;
  cilqr_multi* m = new (std::nothrow) cilqr_multi();
  if (!m) return CILQR_E_INVALID;
  m->G = n_devices;
  m->sh.resize(n_devices);
  for (int r = 0; r < n_devices; ++r) {
    Shard& s = m->sh[r];
    s.device = devices ? devices[r] : r;
    const int rc = cilqr_create(params, s.device, N_max, M_max, S_max, B_max_per_device, &s.h);
    if (rc != CILQR_OK) {
      cilqr_multi_destroy(m);
      return rc;
    }
    s.stream = cilqr_internal_stream(s.h);
  }
  *out = m;
  return CILQR_OK;
}

void cilqr_multi_destroy(cilqr_multi* m) {
  if (!m) return;
  for (Shard& s : m->sh) {
    if (s.h) {
      cudaSetDevice(s.device);
      cudaStreamSynchronize(s.stream);
      if (s.in_buf) cudaFree(s.in_buf);
      if (s.block) cudaFree(s.block);
      if (s.extra) cudaFree(s.extra);
    }
  }
  for (ncclComm_t c : m->comms)
    if (c && m->nccl.CommDestroy) m->nccl.CommDestroy(c);
  for (Shard& s : m->sh)
    if (s.h) cilqr_destroy(s.h);
  delete m;
}

int cilqr_multi_devices(const cilqr_multi* m) { return m ? m->G : 0; }
const char* cilqr_multi_last_error(const cilqr_multi* m) { return m ? m->err.c_str() : ""; }
int cilqr_multi_shard_size(const cilqr_multi* m, int B) { return m && m->G > 0 && B >= 0 ? (B + m->G - 1) / m->G : 0; }

int cilqr_plan_sharded(cilqr_multi* m, const CilqrBatchIn* in, const CilqrBatchOut* out, double* const* gathered_dev) {
  if (!m || !in || !out) return CILQR_E_INVALID;
  if (!out->states || !out->controls || !out->status) return CILQR_E_INVALID;
  if (out->trajectory || out->init_states || out->init_controls || out->cost_hist || out->iter_states ||
      out->iter_controls || out->hist_len)
    return CILQR_E_INVALID;  // the sharded path returns states / controls / status (+ result)
  if (in->B < 0 || in->init_mode != CILQR_INIT_IQR) return CILQR_E_INVALID;  // (caller guesses: single-GPU entry points)
  if (in->B == 0) return CILQR_OK;
  const int G = m->G, B = in->B, per = (B + G - 1) / G;
  const size_t K = in->N + 1, N = in->N;
  if (gathered_dev && m->comms.empty()) {
    if (!m->nccl.load()) {
      m->err = "libnccl.so.2 could not be loaded";
      return CILQR_E_NCCL;
    }
    // (NCCL picks few channels on some boxes: ~140 GB/s per GPU for this all-gather against 444 GB/s with 32 channels,
    // profiles/r02_k_nccl_channels.txt; a setting the user made wins)
    setenv("NCCL_MIN_NCHANNELS", "32", 0);
    std::vector<int> devs(G);
    for (int r = 0; r < G; ++r) devs[r] = m->sh[r].device;
    m->comms.assign(G, nullptr);
    const ncclResult_t e = m->nccl.CommInitAll(m->comms.data(), G, devs.data());
    if (e != 0) {
      m->err = std::string("ncclCommInitAll: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(e) : "error");
      m->comms.clear();
      return CILQR_E_NCCL;
    }
  }
  // ---- one host thread per device: inputs in, solve enqueued
  std::vector<std::thread> th;
  for (int r = 0; r < G; ++r) {
    const int b0 = r * per, nb = std::max(0, std::min(per, B - b0));
    th.emplace_back([=]() { m->sh[r].rc = run_shard(&m->sh[r], in, out, b0, nb, per); });
  }
  for (std::thread& t : th) t.join();
  int rc = CILQR_OK;
  for (int r = 0; r < G; ++r)
    if (m->sh[r].rc != CILQR_OK) rc = m->sh[r].rc;
  // ---- the one collective: every GPU receives every shard's result block (ordered after the solve on each stream)
  if (rc == CILQR_OK && gathered_dev) {
    const size_t cnt = (size_t)per * (K * 6 + N * 2 + CILQR_STATUS_DOUBLES);
    ncclResult_t e = m->nccl.GroupStart();
    for (int r = 0; r < G && e == 0; ++r) {
      cudaSetDevice(m->sh[r].device);
      e = m->nccl.AllGather(m->sh[r].block, gathered_dev[r], cnt, kNcclFloat64, m->comms[r], m->sh[r].stream);
    }
    const ncclResult_t e2 = m->nccl.GroupEnd();
    if (e != 0 || e2 != 0) {
      m->err = std::string("ncclAllGather: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(e ? e : e2) : "error");
      rc = CILQR_E_NCCL;
    }
  }
  // ---- results back to the host, then wait and check every launch
  for (int r = 0; r < G; ++r) {
    Shard& s = m->sh[r];
    const int b0 = r * per, nb = std::max(0, std::min(per, B - b0));
    cudaSetDevice(s.device);
    if (rc == CILQR_OK && nb > 0) {
      const double* st = s.block;
      const double* ct = s.block + (size_t)per * K * 6;
      const double* ss = ct + (size_t)per * N * 2;
      cudaMemcpyAsync(out->states + (size_t)b0 * K * 6, st, (size_t)nb * K * 6 * 8, cudaMemcpyDeviceToHost, s.stream);
      cudaMemcpyAsync(out->controls + (size_t)b0 * N * 2, ct, (size_t)nb * N * 2 * 8, cudaMemcpyDeviceToHost, s.stream);
      cudaMemcpyAsync(out->status + (size_t)b0 * CILQR_STATUS_DOUBLES, ss, (size_t)nb * CILQR_STATUS_DOUBLES * 8,
                      cudaMemcpyDeviceToHost, s.stream);
      if (out->result)
        cudaMemcpyAsync(out->result + (size_t)b0 * K * CILQR_TRAJPOINT_DOUBLES, s.extra,
                        (size_t)nb * K * CILQR_TRAJPOINT_DOUBLES * 8, cudaMemcpyDeviceToHost, s.stream);
    }
  }
  for (int r = 0; r < G; ++r) {
    Shard& s = m->sh[r];
    cudaSetDevice(s.device);
    if (cudaStreamSynchronize(s.stream) != cudaSuccess && rc == CILQR_OK) rc = CILQR_E_CUDA;
    const int rs = cilqr_synchronize(s.h);
    if (rs != CILQR_OK && rc == CILQR_OK) {
      rc = rs;
      m->err = cilqr_last_cuda_error(s.h);
    }
  }
  return rc;
}

}  // extern "C"
