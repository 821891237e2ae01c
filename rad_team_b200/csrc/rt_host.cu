// rt_host.cu — error bookkeeping for librt_b200.so.
#include <cstdio>

#include "rt_host.h"

namespace {
thread_local char g_err[512] = "no CUDA error recorded";
}

const char* rt_last_cuda_error() { return g_err; }

void rt_note_cuda_error(cudaError_t err, const char* what) {
  std::snprintf(g_err, sizeof g_err, "CUDA error %d (%s) in %s", (int)err, cudaGetErrorString(err), what);
}

void rt_note_error_text(const char* text) { std::snprintf(g_err, sizeof g_err, "%s", text); }
