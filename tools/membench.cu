// membench.cu — write-stream micro-benchmarks on B200: which store pattern reaches the fill ceiling?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu && tools/membench
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void st_cs(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_wb(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// (a) torch-like fill: every thread 4 x float4 (64 B), block 128 threads -> 8 KB contiguous per block, one shot
template <bool CS>
__global__ void k_fill_oneshot(float* out, long long n4) {
  const long long base = (long long)blockIdx.x * 512 + threadIdx.x;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long i = base + u * 128;
    if (i < n4) { if (CS) st_cs(out + 4 * i, v); else st_wb(out + 4 * i, v); }
  }
}
// (b) grid-stride persistent fill
template <bool CS>
__global__ void k_fill_stride(float* out, long long n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    if (CS) st_cs(out + 4 * i, v); else st_wb(out + 4 * i, v);
  }
}
// (c) one warp per 12 KB chunk (768 float4), 24 stores per lane, chunks strided by total warps; `reads`: each warp
//     also loads 1 KB (4 x 8 B per lane) per chunk from a second array, two chunks ahead
template <bool CS, bool READS>
__global__ void k_chunks(float* out, const uint2* rec, long long n_chunks, int rec_stride) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  unsigned acc = 0;
  for (long long c = gw; c < n_chunks; c += nw) {
    if (READS) {
      const uint2* p = rec + c * rec_stride;
#pragma unroll
      for (int m = 0; m < 4; ++m) acc += p[lane + 32 * m].x;
    }
    float* o = out + c * 3072;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)(acc & 1));
#pragma unroll 4
    for (int i = lane; i < 768; i += 32) { if (CS) st_cs(o + 4 * i, v); else st_wb(o + 4 * i, v); }
  }
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
  const long long E = 65536, n = E * 3072, n4 = n / 4;  // 805 MB
  const int rec_stride = 364;
  float* out; uint2* rec;
  CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&rec, E * rec_stride * 8)); CK(cudaMemset(rec, 0, E * rec_stride * 8));
  int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  const double gb = n * 4 / 1e9, gbr = gb + E * 1024 / 1e9;
  auto rep = [&](const char* name, float ms, double g) { printf("%-58s %.4f ms  %.0f GB/s\n", name, ms, g / ms * 1e3); };
  rep("cudaMemset", time_ms([&] { cudaMemsetAsync(out, 0, n * 4); }, 20), gb);
  rep("fill one-shot 128thr x 4 float4, st.global", time_ms([&] { k_fill_oneshot<false><<<(unsigned)((n4 + 511) / 512), 128>>>(out, n4); }, 20), gb);
  rep("fill one-shot 128thr x 4 float4, st.global.cs", time_ms([&] { k_fill_oneshot<true><<<(unsigned)((n4 + 511) / 512), 128>>>(out, n4); }, 20), gb);
  for (int c : {8, 32}) {
    char nm[96];
    snprintf(nm, 96, "fill grid-stride %d CTAs/SM x 256, st.global", c);
    rep(nm, time_ms([&] { k_fill_stride<false><<<sm * c, 256>>>(out, n4); }, 20), gb);
    snprintf(nm, 96, "fill grid-stride %d CTAs/SM x 256, st.global.cs", c);
    rep(nm, time_ms([&] { k_fill_stride<true><<<sm * c, 256>>>(out, n4); }, 20), gb);
  }
  for (int c : {2, 8, 24, 55}) {
    char nm[96];
    unsigned nb = c == 55 ? (unsigned)(E / 8) : (unsigned)(sm * c);
    snprintf(nm, 96, "warp-per-12KB-chunk %u CTAs x 256, st.global", nb);
    rep(nm, time_ms([&] { k_chunks<false, false><<<nb, 256>>>(out, rec, E, rec_stride); }, 20), gb);
    snprintf(nm, 96, "warp-per-12KB-chunk %u CTAs x 256, st.global.cs", nb);
    rep(nm, time_ms([&] { k_chunks<true, false><<<nb, 256>>>(out, rec, E, rec_stride); }, 20), gb);
    snprintf(nm, 96, "  + 1 KB record read per chunk, st.global", nb);
    rep(nm, time_ms([&] { k_chunks<false, true><<<nb, 256>>>(out, rec, E, rec_stride); }, 20), gbr);
    snprintf(nm, 96, "  + 1 KB record read per chunk, st.global.cs", nb);
    rep(nm, time_ms([&] { k_chunks<true, true><<<nb, 256>>>(out, rec, E, rec_stride); }, 20), gbr);
  }
  return 0;
}
