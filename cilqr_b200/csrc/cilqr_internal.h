// cilqr_internal.h -- what the other translation units of libcilqr_b200.so need from the handle
// (defined in cilqr_capi.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include "../../include/cilqr_b200.h"

extern "C" {
int cilqr_internal_device(const cilqr_handle* h);
cudaStream_t cilqr_internal_stream(cilqr_handle* h);
int cilqr_internal_num_sms(const cilqr_handle* h);
int cilqr_internal_smem_optin(const cilqr_handle* h);
// records the CUDA error text on the handle and returns CILQR_E_CUDA
int cilqr_internal_fail(cilqr_handle* h, cudaError_t e, const char* where);
// grow-only device scratch owned by the handle (freed by cilqr_destroy); slots 0, 1: dp planner, 2: tracker
int cilqr_internal_scratch(cilqr_handle* h, int slot, size_t bytes, char** out);
// a pair of timing events owned by the handle, per slot
int cilqr_internal_events(cilqr_handle* h, int slot, cudaEvent_t* e0, cudaEvent_t* e1);
// dp_capi.cu: raises dp_plan_kernel's dynamic shared-memory limit (once per device, from cilqr_create)
int cilqr_internal_dp_set_smem(int bytes);
// tracker_capi.cu: the same for tracker_kernel
int cilqr_internal_tracker_set_smem(int bytes);
}
