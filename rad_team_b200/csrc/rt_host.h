// rt_host.h — host-side helpers shared by the translation units of librt_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/rt_b200.h"

// Last CUDA failure seen by this library (for rt_status_string(RT_ERR_CUDA)).
const char* rt_last_cuda_error();
void rt_note_cuda_error(cudaError_t err, const char* what);
void rt_note_error_text(const char* text);

// A handle belongs to the device that was current when it was created; its entry points run there
// whatever device the calling thread has selected since, and restore the caller's choice on return.
struct RtDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit RtDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~RtDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  RtDeviceGuard(const RtDeviceGuard&) = delete;
  RtDeviceGuard& operator=(const RtDeviceGuard&) = delete;
};

// K4b handle (rt_stats.cu); rt_comm.cu merges into / out of its state.
struct rt_moments {
  int device = 0;
  double* state = nullptr;     // [3] count, mean, M2
  double* partials = nullptr;  // [max_blocks][3]
  int max_blocks = 0;
};

// NVTX ranges around the C-ABI entry points (rt_env_step / rasterize / reset, rollout close) so that an nsys timeline
// reads in the library's own terms.  Header-only NVTX3: a no-op costing a few nanoseconds unless a profiler is attached.
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define RT_HAVE_NVTX 1
#endif
#endif
struct RtRange {
#ifdef RT_HAVE_NVTX
  explicit RtRange(const char* name) { nvtxRangePushA(name); }
  ~RtRange() { nvtxRangePop(); }
#else
  explicit RtRange(const char*) {}
#endif
  RtRange(const RtRange&) = delete;
  RtRange& operator=(const RtRange&) = delete;
};

#define RT_CUDA(call)                         \
  do {                                        \
    cudaError_t rt_e_ = (call);               \
    if (rt_e_ != cudaSuccess) {               \
      rt_note_cuda_error(rt_e_, #call);       \
      return RT_ERR_CUDA;                     \
    }                                         \
  } while (0)

#define RT_CUDA_OR(call, cleanup)             \
  do {                                        \
    cudaError_t rt_e_ = (call);               \
    if (rt_e_ != cudaSuccess) {               \
      rt_note_cuda_error(rt_e_, #call);       \
      cleanup;                                \
      return RT_ERR_CUDA;                     \
    }                                         \
  } while (0)

#define RT_CUDA_LAUNCH_CHECK() RT_CUDA(cudaGetLastError())
