// rt_stats.cu — the src/math_utils interfaces as batched sm_100a kernels:
//   rt_welford_*    per-stream WelfordsOnlineStandardization   (standardize.py:7-107)
//   rt_moments_*    K4b: batch running moments, warp-shuffle Chan merges, NCCL-mergeable
//   rt_normalize    Normalizer.normalize                        (normalize.py:25-55)
//   rt_lognormalize BensLogNormalizer                           (normalize.py:79-115)
//   rt_estimator_*  IntensityResamplingEstimator                (estimator.py:6-109)
#include <math_constants.h>

#include <new>
#include <vector>

#include "rt_device.cuh"
#include "rt_host.h"
#include "rt_moments.cuh"

using namespace rt;

// ============================================================================================
// Per-stream Welford: state double[n][6] = mean, M2, var, std, zmax, zmin; int32[n] count.
struct rt_welford {
  int device = 0;
  long long n = 0;
  double* state = nullptr;
  int32_t* count = nullptr;
};

namespace {

__global__ void __launch_bounds__(256) k_welford_reset(double* __restrict__ st, int32_t* __restrict__ cnt,
                                                      long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2* p = reinterpret_cast<double2*>(st + i * 6);
  p[0] = make_double2(0.0, 0.0);
  p[1] = make_double2(0.0, 1.0);
  p[2] = make_double2(0.0, 0.0);
  cnt[i] = 0;
}

__device__ __forceinline__ WelfordState load_w(const double* st, const int32_t* cnt, long long i) {
  const double2* p = reinterpret_cast<const double2*>(st + i * 6);
  const double2 a = p[0], b = p[1], c = p[2];
  WelfordState w;
  w.mean = a.x; w.m2 = a.y; w.var = b.x; w.sd = b.y; w.zmax = c.x; w.zmin = c.y;
  w.count = cnt[i];
  return w;
}

__global__ void __launch_bounds__(256)
k_welford_update(double* __restrict__ st, int32_t* __restrict__ cnt, const double* __restrict__ reading,
                 int32_t* __restrict__ err, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = reading[i];
  if (!(r >= 0.0)) {  // standardize.py:63
    if (err) err[i] = RT_ERR_ASSERT;
    return;
  }
  WelfordState w = load_w(st, cnt, i);
  welford_update(w, r);
  double2* p = reinterpret_cast<double2*>(st + i * 6);
  p[0] = make_double2(w.mean, w.m2);
  p[1] = make_double2(w.var, w.sd);
  p[2] = make_double2(w.zmax, w.zmin);
  cnt[i] = w.count;
  if (err) err[i] = RT_OK;
}

__global__ void __launch_bounds__(256)
k_welford_standardize(const double* __restrict__ st, const double* __restrict__ reading,
                      double* __restrict__ z, int32_t* __restrict__ err, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = reading[i];
  if (!(r >= 0.0)) {  // standardize.py:92
    if (err) err[i] = RT_ERR_ASSERT;
    z[i] = CUDART_NAN;
    return;
  }
  const double2* p = reinterpret_cast<const double2*>(st + i * 6);
  const double mean = p[0].x, sd = p[1].y;
  z[i] = __ddiv_rn(__dsub_rn(r, mean), sd);
  if (err) err[i] = RT_OK;
}

}  // namespace

extern "C" {

int rt_welford_create(int64_t n, rt_welford** out) {
  if (n < 1 || !out) return RT_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    rt_note_cuda_error(cudaGetLastError(), "cudaGetDeviceCount");
    return RT_ERR_CUDA;
  }
  rt_welford* w = new (std::nothrow) rt_welford();
  if (!w) return RT_ERR_NOMEM;
  w->n = n;
  cudaGetDevice(&w->device);
  if (cudaMalloc((void**)&w->state, (size_t)n * 48) != cudaSuccess ||
      cudaMalloc((void**)&w->count, (size_t)n * 4) != cudaSuccess) {
    rt_note_cuda_error(cudaGetLastError(), "cudaMalloc(welford)");
    rt_welford_destroy(w);
    return RT_ERR_NOMEM;
  }
  const int rc = rt_welford_reset(w, nullptr);
  if (rc != RT_OK) {
    rt_welford_destroy(w);
    return rc;
  }
  RT_CUDA_OR(cudaDeviceSynchronize(), { rt_welford_destroy(w); });
  *out = w;
  return RT_OK;
}

int rt_welford_destroy(rt_welford* w) {
  if (!w) return RT_OK;
  RtDeviceGuard on_device(w->device);
  if (w->state) cudaFree(w->state);
  if (w->count) cudaFree(w->count);
  delete w;
  return RT_OK;
}

int rt_welford_reset(rt_welford* w, void* stream) {
  if (!w) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(w->device);
  k_welford_reset<<<(unsigned)((w->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w->state, w->count, w->n);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_welford_update(rt_welford* w, const double* reading_dev, int32_t* err_dev, void* stream) {
  if (!w || !reading_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(w->device);
  k_welford_update<<<(unsigned)((w->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w->state, w->count, reading_dev, err_dev, w->n);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_welford_standardize(rt_welford* w, const double* reading_dev, double* z_dev, int32_t* err_dev,
                           void* stream) {
  if (!w || !reading_dev || !z_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(w->device);
  k_welford_standardize<<<(unsigned)((w->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w->state, reading_dev, z_dev, err_dev, w->n);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_welford_buffers(rt_welford* w, void** state_dev, void** count_dev) {
  if (!w) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(w->device);
  if (state_dev) *state_dev = w->state;
  if (count_dev) *count_dev = w->count;
  return RT_OK;
}

}  // extern "C"

// ============================================================================================
// K4b: batch running moments.  state double[3] = {count, mean, M2}.
//
// Pass 1 (k_moments_partial): each thread folds 16-byte vectors of fp32 readings into a private
// (n, mean, M2) — four values at a time: exact two-pass moments of the quad in registers, then one
// Chan merge — so there is one fp64 divide per 4 readings, not per reading.  Warps combine with
// five shuffle-down Chan merges, warps of a CTA through shared memory, and each CTA writes one
// partial triple.  Pass 2 (k_moments_final, one warp) merges the CTA partials in index order and
// folds the result into the running state.  Every merge order is fixed by (n, grid), so the result
// is deterministic run to run; it differs from the sequential recurrence only by rounding.
// (struct rt_moments lives in rt_host.h: rt_comm.cu launches on its state too)

namespace {

__device__ __forceinline__ Mom warp_merge(Mom m) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Mom o;
    o.n = __shfl_down_sync(0xffffffffu, m.n, off);
    o.mean = __shfl_down_sync(0xffffffffu, m.mean, off);
    o.m2 = __shfl_down_sync(0xffffffffu, m.m2, off);
    m = chan(m, o);
  }
  return m;
}

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Per-thread accumulation uses SHIFTED sums: with K = the thread's first reading, S1 = sum(x - K) and
// S2 = sum((x - K)^2) in fp64 (five fp64 operations per reading, no divide).  K lies inside the data, so
// M2 = S2 - S1^2/n loses at most a digit or two of fp64 — far inside the 1e-10 the tests ask for — and
// the fp64 pipe stays well under the HBM rate.  Everything above thread level is a Chan merge.
struct Shifted {
  double k, s1, s2, n;
};
__device__ __forceinline__ void fold_one(Shifted& a, float x) {
  const double dx = (double)x - a.k;
  a.s1 += dx;
  a.s2 += dx * dx;
}
__device__ __forceinline__ void fold_quad(Shifted& a, const float4 v) {
  fold_one(a, v.x);
  fold_one(a, v.y);
  fold_one(a, v.z);
  fold_one(a, v.w);
  a.n += 4.0;
}

constexpr int kMomThreads = 256;
constexpr int kMomUnroll = 4;  // 16-byte loads in flight per thread

__global__ void __launch_bounds__(kMomThreads)
k_moments_partial(const float* __restrict__ x0, long long n0, int head, double* __restrict__ partials) {
  // x0 may start anywhere on a 4-byte boundary (a row of a [T, E, A] buffer, a slice): its first `head` (< 4)
  // readings are folded one by one like the tail, the 16-byte vector body starts at the aligned x
  __shared__ Mom s_w[kMomThreads / 32];
  const float* __restrict__ x = x0 + head;
  const long long n = n0 - head;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * kMomThreads;
  long long i = (long long)blockIdx.x * kMomThreads + threadIdx.x;
  Shifted sh = {0.0, 0.0, 0.0, 0.0};
  if (i < n4) sh.k = (double)x[4 * i];
  // Software pipeline: the kMomUnroll loads of the NEXT round are issued before the current round is folded, so
  // every thread keeps kMomUnroll 16-byte loads in flight whatever order the scheduler gives the folds (left to
  // itself ptxas interleaved load / fold / load / fold: one load in flight per thread, 4.3 TB/s instead of 6).
  // Two accumulator pairs halve the dependent fp64 chain; they are added once at the end (fixed order).
  Shifted sb = {sh.k, 0.0, 0.0, 0.0};
  float4 cur[kMomUnroll], nxt[kMomUnroll];
  bool have = i + (kMomUnroll - 1) * stride < n4;
  if (have) {
#pragma unroll
    for (int u = 0; u < kMomUnroll; ++u) cur[u] = ld_stream_f4(x + 4 * (i + u * stride));
  }
  while (have) {
    const long long j = i + kMomUnroll * stride;
    const bool more = j + (kMomUnroll - 1) * stride < n4;
    if (more) {
#pragma unroll
      for (int u = 0; u < kMomUnroll; ++u) nxt[u] = ld_stream_f4(x + 4 * (j + u * stride));
    }
#pragma unroll
    for (int u = 0; u < kMomUnroll; ++u) fold_quad((u & 1) ? sb : sh, cur[u]);
#pragma unroll
    for (int u = 0; u < kMomUnroll; ++u) cur[u] = nxt[u];
    i = j;
    have = more;
  }
  for (; i < n4; i += stride) fold_quad(sh, ld_stream_f4(x + 4 * i));
  sh.s1 += sb.s1;
  sh.s2 += sb.s2;
  sh.n += sb.n;
  Mom acc = {0.0, 0.0, 0.0};
  if (sh.n > 0.0) {
    acc.n = sh.n;
    acc.mean = sh.k + sh.s1 / sh.n;
    acc.m2 = sh.s2 - sh.s1 * (sh.s1 / sh.n);
    if (acc.m2 < 0.0) acc.m2 = 0.0;
  }
  // tail (n % 4 readings): one single-reading merge each, by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    Mom q = {1.0, (double)x[(n4 << 2) + threadIdx.x], 0.0};
    acc = chan(acc, q);
  }
  if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + head) {  // head: threads of the second warp
    Mom q = {1.0, (double)x0[threadIdx.x - 32], 0.0};
    acc = chan(acc, q);
  }
  acc = warp_merge(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    Mom m = threadIdx.x < kMomThreads / 32 ? s_w[threadIdx.x] : Mom{0.0, 0.0, 0.0};
    m = warp_merge(m);
    if (threadIdx.x == 0) {
      partials[3 * blockIdx.x + 0] = m.n;
      partials[3 * blockIdx.x + 1] = m.mean;
      partials[3 * blockIdx.x + 2] = m.m2;
    }
  }
}

// One CTA of 1024 threads: thread i merges partials i, i+1024, ... (independent loads, issued together),
// then the fixed shuffle / shared-memory tree, then the running state.  (A single warp walking the
// partials took 26 us of dependent loads — ncu, profiles/r02g — against 175 us for the streaming pass.)
constexpr int kFinalThreads = 1024;
__global__ void __launch_bounds__(kFinalThreads)
k_moments_final(const double* __restrict__ partials, int n_partials, double* __restrict__ state, int replace) {
  __shared__ Mom s_w[kFinalThreads / 32];
  Mom m = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n_partials; i += kFinalThreads) {
    Mom p = {partials[3 * i], partials[3 * i + 1], partials[3 * i + 2]};
    m = chan(m, p);
  }
  m = warp_merge(m);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    Mom t = s_w[threadIdx.x];
    t = warp_merge(t);
    if (threadIdx.x == 0) {
      Mom s = {state[0], state[1], state[2]};
      if (replace) s = Mom{0.0, 0.0, 0.0};
      s = chan(s, t);
      state[0] = s.n;
      state[1] = s.mean;
      state[2] = s.m2;
    }
  }
}

// rank-ordered merge of gathered triples (strictly sequential: bit-identical on every rank)
__global__ void k_moments_merge(const double* __restrict__ gathered, int n_ranks, int stride,
                                double* __restrict__ state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Mom s = {0.0, 0.0, 0.0};
  for (int r = 0; r < n_ranks; ++r) {
    Mom p = {gathered[stride * r], gathered[stride * r + 1], gathered[stride * r + 2]};
    s = chan(s, p);
  }
  state[0] = s.n;
  state[1] = s.mean;
  state[2] = s.m2;
}

__global__ void __launch_bounds__(256)
k_moments_standardize(const double* __restrict__ state, const float* __restrict__ x0, float* __restrict__ out0,
                      long long n0, int head) {
  // head (< 4): readings in front of the first 16-byte boundary, when input and output share their misalignment
  const float* __restrict__ x = x0 + head;
  float* __restrict__ out = out0 + head;
  const long long n = n0 - head;
  const double cnt = state[0], mean = state[1], m2 = state[2];
  double sd = 1.0;
  if (cnt >= 2.0) {
    const double s = sqrt(m2 / (cnt - 1.0));
    sd = s > 1.0 ? s : 1.0;  // standardize.py:74
  }
  const double inv = 1.0 / sd;  // one divide per thread; (x - mean) * inv differs from the quotient by
                                // < 1 ulp of fp64, invisible after rounding to the fp32 output
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    const float4 v0 = ld_stream_f4(x + 4 * i), v1 = ld_stream_f4(x + 4 * (i + stride));
    float4 o0, o1;
    o0.x = (float)(((double)v0.x - mean) * inv); o0.y = (float)(((double)v0.y - mean) * inv);
    o0.z = (float)(((double)v0.z - mean) * inv); o0.w = (float)(((double)v0.w - mean) * inv);
    o1.x = (float)(((double)v1.x - mean) * inv); o1.y = (float)(((double)v1.y - mean) * inv);
    o1.z = (float)(((double)v1.z - mean) * inv); o1.w = (float)(((double)v1.w - mean) * inv);
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * i), "f"(o0.x), "f"(o0.y), "f"(o0.z),
                 "f"(o0.w) : "memory");
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * (i + stride)), "f"(o1.x), "f"(o1.y),
                 "f"(o1.z), "f"(o1.w) : "memory");
  }
  for (; i < n4; i += stride) {
    const float4 v = ld_stream_f4(x + 4 * i);
    float4 o;
    o.x = (float)(((double)v.x - mean) * inv); o.y = (float)(((double)v.y - mean) * inv);
    o.z = (float)(((double)v.z - mean) * inv); o.w = (float)(((double)v.w - mean) * inv);
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * i), "f"(o.x), "f"(o.y), "f"(o.z),
                 "f"(o.w) : "memory");
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const long long j = (n4 << 2) + threadIdx.x;
    out[j] = (float)(((double)x[j] - mean) * inv);
  }
  if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + head)
    out0[threadIdx.x - 32] = (float)(((double)x0[threadIdx.x - 32] - mean) * inv);
}

// input and output misaligned DIFFERENTLY with respect to 16 bytes: 4-byte accesses (rare; same arithmetic)
__global__ void __launch_bounds__(256)
k_moments_standardize_scalar(const double* __restrict__ state, const float* __restrict__ x, float* __restrict__ out,
                             long long n) {
  const double cnt = state[0], mean = state[1], m2 = state[2];
  double sd = 1.0;
  if (cnt >= 2.0) {
    const double s = sqrt(m2 / (cnt - 1.0));
    sd = s > 1.0 ? s : 1.0;
  }
  const double inv = 1.0 / sd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)(((double)x[i] - mean) * inv);
}

}  // namespace

extern "C" {

int rt_moments_create(rt_moments** out) {
  if (!out) return RT_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    rt_note_cuda_error(cudaGetLastError(), "cudaGetDeviceCount");
    return RT_ERR_CUDA;
  }
  rt_moments* m = new (std::nothrow) rt_moments();
  if (!m) return RT_ERR_NOMEM;
  int n_sm = 148;
  cudaGetDevice(&m->device);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, m->device);
  m->max_blocks = n_sm * 8;  // 8 resident 256-thread CTAs per SM: one full wave
  if (cudaMalloc((void**)&m->state, 3 * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&m->partials, (size_t)m->max_blocks * 3 * sizeof(double)) != cudaSuccess) {
    rt_note_cuda_error(cudaGetLastError(), "cudaMalloc(moments)");
    rt_moments_destroy(m);
    return RT_ERR_NOMEM;
  }
  RT_CUDA_OR(cudaMemset(m->state, 0, 3 * sizeof(double)), { rt_moments_destroy(m); });
  *out = m;
  return RT_OK;
}

int rt_moments_destroy(rt_moments* m) {
  if (!m) return RT_OK;
  RtDeviceGuard on_device(m->device);
  if (m->state) cudaFree(m->state);
  if (m->partials) cudaFree(m->partials);
  delete m;
  return RT_OK;
}

int rt_moments_reset(rt_moments* m, void* stream) {
  if (!m) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(m->device);
  RT_CUDA(cudaMemsetAsync(m->state, 0, 3 * sizeof(double), (cudaStream_t)stream));
  return RT_OK;
}

int rt_moments_update(rt_moments* m, const float* x_dev, int64_t n, void* stream) {
  if (!m || !x_dev || n < 0) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(m->device);
  RtRange nvtx_range("rt_moments_update");
  if (((uintptr_t)x_dev & 3) != 0) return RT_ERR_BAD_ARG;  // not even a float boundary
  if (n == 0) return RT_OK;
  // any contiguous float32 view works (a row readings[t] of a [T, E, A] buffer, x[1:]): up to three readings in
  // front of the first 16-byte boundary are folded singly, the vector loads start at the boundary
  int head = (int)(((16 - ((uintptr_t)x_dev & 15)) & 15) >> 2);
  if (head > n) head = (int)n;
  long long blocks = (((n - head) >> 2) + kMomThreads * kMomUnroll - 1) / (kMomThreads * kMomUnroll);
  if (blocks < 1) blocks = 1;
  if (blocks > m->max_blocks) blocks = m->max_blocks;
  k_moments_partial<<<(unsigned)blocks, kMomThreads, 0, (cudaStream_t)stream>>>(x_dev, n, head, m->partials);
  k_moments_final<<<1, kFinalThreads, 0, (cudaStream_t)stream>>>(m->partials, (int)blocks, m->state, 0);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_moments_standardize(rt_moments* m, const float* x_dev, float* out_dev, int64_t n, void* stream) {
  if (!m || !x_dev || !out_dev || n < 0) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(m->device);
  RtRange nvtx_range("rt_moments_standardize");
  if ((((uintptr_t)x_dev | (uintptr_t)out_dev) & 3) != 0) return RT_ERR_BAD_ARG;
  if (n == 0) return RT_OK;
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (long long)m->max_blocks * 4) blocks = (long long)m->max_blocks * 4;
  if ((((uintptr_t)x_dev ^ (uintptr_t)out_dev) & 15) != 0) {  // no common 16-byte phase: 4-byte accesses
    k_moments_standardize_scalar<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(m->state, x_dev, out_dev, n);
  } else {
    int head = (int)(((16 - ((uintptr_t)x_dev & 15)) & 15) >> 2);
    if (head > n) head = (int)n;
    k_moments_standardize<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(m->state, x_dev, out_dev, n, head);
  }
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_moments_state(rt_moments* m, void** state_dev) {
  if (!m || !state_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(m->device);
  *state_dev = m->state;
  return RT_OK;
}

int rt_moments_merge(rt_moments* m, const double* gathered_dev, int32_t n_ranks, int32_t stride_doubles,
                     void* stream) {
  if (!m || !gathered_dev || n_ranks < 1 || stride_doubles < 3) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(m->device);
  k_moments_merge<<<1, 32, 0, (cudaStream_t)stream>>>(gathered_dev, n_ranks, stride_doubles, m->state);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

}  // extern "C"

// ============================================================================================
// Elementwise normalisers.
namespace {

__global__ void __launch_bounds__(256)
k_normalize(const double* __restrict__ cur, const double* __restrict__ mx, const double* __restrict__ mn,
            double* __restrict__ out, int32_t* __restrict__ err, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r;
  const bool ok = normalize_minmax(cur[i], mx[i], mn != nullptr, mn ? mn[i] : 0.0, r);
  out[i] = r;
  if (err) err[i] = ok ? RT_OK : RT_ERR_ASSERT;
}

// normalize.py:110: log(step + v, max) * 1 / log(step * max, max), math.log(x, b) = ln x / ln b
__global__ void __launch_bounds__(256)
k_lognormalize(const double* __restrict__ cur, double mx, int step, double* __restrict__ out,
               int32_t* __restrict__ err, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = cur[i];
  double r = CUDART_NAN;
  int rc = RT_ERR_ASSERT;
  if (v >= 0.0 && mx > 0.0 && step > 0 && mx != 1.0) {  // normalize.py:90 (max == 1: ZeroDivisionError)
    const double lb = dlog(mx);
    const double num = __ddiv_rn(dlog(__dadd_rn((double)step, v)), lb);
    const double den = __ddiv_rn(dlog(__dmul_rn((double)step, mx)), lb);
    const double q = __ddiv_rn(__dmul_rn(num, 1.0), den);
    if (q >= 0.0 && q <= 1.0) {  // :111-113
      r = q;
      rc = RT_OK;
    }
  }
  out[i] = r;
  if (err) err[i] = rc;
}

}  // namespace

extern "C" {

int rt_normalize(const double* current_dev, const double* max_dev, const double* min_dev, double* out_dev,
                 int32_t* err_dev, int64_t n, void* stream) {
  if (!current_dev || !max_dev || !out_dev || n < 0) return RT_ERR_BAD_ARG;
  if (n == 0) return RT_OK;
  k_normalize<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(current_dev, max_dev, min_dev,
                                                                          out_dev, err_dev, n);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_lognormalize(const double* current_dev, double max, int32_t step_size, double* out_dev, int32_t* err_dev,
                    int64_t n, void* stream) {
  if (!current_dev || !out_dev || n < 0) return RT_ERR_BAD_ARG;
  if (n == 0) return RT_OK;
  k_lognormalize<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(current_dev, max, step_size,
                                                                             out_dev, err_dev, n);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

}  // extern "C"

// ============================================================================================
// Stand-alone IntensityResamplingEstimator: n estimators, `n_keys` key slots, `capacity` readings
// each.  Readings are nodes {value, key, next-larger-of-same-key}; per key the nodes form a list
// sorted by value (median = walk to the middle), and the node array itself is insertion order
// (get_buffer = filtered scan).
struct rt_estimator {
  int device = 0;
  long long n = 0;
  int n_keys = 0, capacity = 0;
  double* value = nullptr;   // [n][capacity]
  int32_t* key = nullptr;    // [n][capacity]
  int32_t* next = nullptr;   // [n][capacity]  +1 encoded
  int32_t* head = nullptr;   // [n][n_keys]    +1 encoded, smallest value of the key
  int32_t* klen = nullptr;   // [n][n_keys]    readings per key
  int32_t* used = nullptr;   // [n]
  double* maxmin = nullptr;  // [n][2]
};

namespace {

__device__ __forceinline__ double est_median(const rt_estimator& s, long long i, int key) {
  const int n = s.klen[i * s.n_keys + key];
  const int k1 = (n - 1) >> 1, k2 = n >> 1;
  int cur = s.head[i * s.n_keys + key];
  double m1 = 0.0, m2 = 0.0;
  for (int j = 0; j <= k2; ++j) {
    const double v = s.value[i * s.capacity + (cur - 1)];
    if (j == k1) m1 = v;
    if (j == k2) m2 = v;
    cur = s.next[i * s.capacity + (cur - 1)];
  }
  return (n & 1) ? m2 : __ddiv_rn(__dadd_rn(m1, m2), 2.0);  // statistics.median
}

__global__ void __launch_bounds__(128)
k_est_update(const rt_estimator s, const int32_t* __restrict__ key, const double* __restrict__ value,
             double* __restrict__ est_out, int32_t* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  const int k = key[i];
  const int idx = s.used[i];
  if (k < 0 || k >= s.n_keys || idx >= s.capacity) {  // key out of range / reading pool full
    if (err) err[i] = idx >= s.capacity ? RT_ERR_NOMEM : RT_ERR_BAD_ARG;
    if (est_out) est_out[i] = CUDART_NAN;  // never leave the output uninitialised
    return;
  }
  const double v = value[i];
  // sorted insert (after equal values)
  int pred = 0, cur = s.head[i * s.n_keys + k];
  for (int guard = 0; cur != 0 && !(s.value[i * s.capacity + (cur - 1)] > v) && guard <= idx; ++guard) {
    pred = cur;
    cur = s.next[i * s.capacity + (cur - 1)];
  }
  s.value[i * s.capacity + idx] = v;
  s.key[i * s.capacity + idx] = k;
  s.next[i * s.capacity + idx] = cur;
  if (pred == 0)
    s.head[i * s.n_keys + k] = idx + 1;
  else
    s.next[i * s.capacity + (pred - 1)] = idx + 1;
  s.klen[i * s.n_keys + k] += 1;
  s.used[i] = idx + 1;
  const double est = est_median(s, i, k);  // estimator.py:44
  double mx = s.maxmin[2 * i], mn = s.maxmin[2 * i + 1];
  est_maxmin(est, mx, mn);  // :45-48
  s.maxmin[2 * i] = mx;
  s.maxmin[2 * i + 1] = mn;
  if (est_out) est_out[i] = est;
  if (err) err[i] = RT_OK;
}

__global__ void __launch_bounds__(128)
k_est_estimate(const rt_estimator s, const int32_t* __restrict__ key, double* __restrict__ est_out,
               int32_t* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  const int k = key[i];
  if (k < 0 || k >= s.n_keys || s.klen[i * s.n_keys + k] == 0) {  // estimator.py:68-69
    est_out[i] = CUDART_NAN;
    if (err) err[i] = RT_ERR_KEY;
    return;
  }
  est_out[i] = est_median(s, i, k);
  if (err) err[i] = RT_OK;
}

__global__ void __launch_bounds__(128)
k_est_buffer(const rt_estimator s, const int32_t* __restrict__ key, double* __restrict__ out,
             int32_t* __restrict__ n_out, int max_out, int32_t* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  const int k = key[i];
  if (k < 0 || k >= s.n_keys || s.klen[i * s.n_keys + k] == 0) {  // estimator.py:57-58
    n_out[i] = 0;
    if (err) err[i] = RT_ERR_KEY;
    return;
  }
  int m = 0;
  const int used = s.used[i];
  for (int j = 0; j < used; ++j)
    if (s.key[i * s.capacity + j] == k) {
      if (m < max_out) out[i * max_out + m] = s.value[i * s.capacity + j];
      ++m;
    }
  n_out[i] = m;
  if (err) err[i] = RT_OK;
}

}  // namespace

extern "C" {

int rt_estimator_create(int64_t n, int32_t n_keys, int32_t capacity, rt_estimator** out) {
  if (n < 1 || n_keys < 1 || capacity < 1 || !out) return RT_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    rt_note_cuda_error(cudaGetLastError(), "cudaGetDeviceCount");
    return RT_ERR_CUDA;
  }
  rt_estimator* s = new (std::nothrow) rt_estimator();
  if (!s) return RT_ERR_NOMEM;
  s->n = n;
  s->n_keys = n_keys;
  s->capacity = capacity;
  cudaGetDevice(&s->device);
  const size_t nc = (size_t)n * capacity, nk = (size_t)n * n_keys;
  bool ok = cudaMalloc((void**)&s->value, nc * 8) == cudaSuccess &&
            cudaMalloc((void**)&s->key, nc * 4) == cudaSuccess &&
            cudaMalloc((void**)&s->next, nc * 4) == cudaSuccess &&
            cudaMalloc((void**)&s->head, nk * 4) == cudaSuccess &&
            cudaMalloc((void**)&s->klen, nk * 4) == cudaSuccess &&
            cudaMalloc((void**)&s->used, (size_t)n * 4) == cudaSuccess &&
            cudaMalloc((void**)&s->maxmin, (size_t)n * 16) == cudaSuccess;
  if (!ok) {
    rt_note_cuda_error(cudaGetLastError(), "cudaMalloc(estimator)");
    rt_estimator_destroy(s);
    return RT_ERR_NOMEM;
  }
  const int rc = rt_estimator_reset(s, nullptr);
  if (rc != RT_OK) {
    rt_estimator_destroy(s);
    return rc;
  }
  RT_CUDA_OR(cudaDeviceSynchronize(), { rt_estimator_destroy(s); });
  *out = s;
  return RT_OK;
}

int rt_estimator_destroy(rt_estimator* s) {
  if (!s) return RT_OK;
  RtDeviceGuard on_device(s->device);
  void* ptrs[] = {s->value, s->key, s->next, s->head, s->klen, s->used, s->maxmin};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete s;
  return RT_OK;
}

int rt_estimator_reset(rt_estimator* s, void* stream) {  // estimator.py:25-29
  if (!s) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nk = (size_t)s->n * s->n_keys;
  RT_CUDA(cudaMemsetAsync(s->head, 0, nk * 4, st));
  RT_CUDA(cudaMemsetAsync(s->klen, 0, nk * 4, st));
  RT_CUDA(cudaMemsetAsync(s->used, 0, (size_t)s->n * 4, st));
  RT_CUDA(cudaMemsetAsync(s->maxmin, 0, (size_t)s->n * 16, st));
  return RT_OK;
}

int rt_estimator_update(rt_estimator* s, const int32_t* key_dev, const double* value_dev, double* est_dev,
                        int32_t* err_dev, void* stream) {
  if (!s || !key_dev || !value_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(s->device);
  k_est_update<<<(unsigned)((s->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*s, key_dev, value_dev,
                                                                              est_dev, err_dev);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_estimator_estimate(rt_estimator* s, const int32_t* key_dev, double* est_dev, int32_t* err_dev,
                          void* stream) {
  if (!s || !key_dev || !est_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(s->device);
  k_est_estimate<<<(unsigned)((s->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*s, key_dev, est_dev,
                                                                                err_dev);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_estimator_buffer(rt_estimator* s, const int32_t* key_dev, double* out_dev, int32_t* n_out_dev,
                        int32_t max_out, int32_t* err_dev, void* stream) {
  if (!s || !key_dev || !out_dev || !n_out_dev || max_out < 1) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(s->device);
  k_est_buffer<<<(unsigned)((s->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*s, key_dev, out_dev,
                                                                              n_out_dev, max_out, err_dev);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_estimator_maxmin(rt_estimator* s, void** maxmin_dev) {
  if (!s || !maxmin_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(s->device);
  *maxmin_dev = s->maxmin;
  return RT_OK;
}

}  // extern "C"
