// rt_moments.cuh — (count, mean, M2) triples and Chan's pairwise merge, shared by the batch-moments kernels
// (rt_stats.cu) and the cross-GPU exchange (rt_comm.cu).  Plain IEEE operations in a fixed order (the library
// is compiled with -fmad=false), so a merge gives the same bits on every rank and in the CPU oracle.
#pragma once

namespace rt {

struct Mom {
  double n, mean, m2;
};

__device__ __forceinline__ Mom chan(const Mom a, const Mom b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Mom r;
  r.n = a.n + b.n;
  const double delta = b.mean - a.mean;
  const double f = b.n / r.n;
  r.mean = a.mean + delta * f;
  r.m2 = a.m2 + b.m2 + delta * delta * (a.n * f);
  return r;
}

}  // namespace rt
