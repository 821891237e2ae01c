// rt_device.cuh — device-side arithmetic shared by the kernels of librt_b200.so.
//
// Everything here follows the written spec in DESIGN.md §3 (RNG streams, the
// deterministic log/exp, the Poisson sampler, the line-of-sight predicate, the
// reference's Welford / Normalizer recurrences).  All floating point is
// expressed with explicit round-to-nearest intrinsics (__dadd_rn, __fmul_rn, …)
// so that ptxas can never contract a multiply-add: the CPU oracle is compiled
// with -ffp-contract=off and the two must agree bit for bit.  (The translation
// units are also built with -fmad=false; the intrinsics make the intent local.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/rt_b200.h"

namespace rt {

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011).  Pure integer work.
struct Philox {
  uint32_t c1, c2, c3;  // counter words 1..3 (word 0 is the block index)
  uint32_t k0, k1;      // key

  __device__ __forceinline__ void block(uint32_t c0, uint32_t out[4]) const {
    uint32_t a = c0, b = c1, c = c2, d = c3, x = k0, y = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, a), lo0 = 0xD2511F53u * a;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c), lo1 = 0xCD9E8D57u * c;
      a = hi1 ^ b ^ x;
      b = lo1;
      c = hi0 ^ d ^ y;
      d = lo0;
      x += 0x9E3779B9u;
      y += 0xBB67AE85u;
    }
    out[0] = a; out[1] = b; out[2] = c; out[3] = d;
  }
};

// RNG streams
enum : int { STREAM_MEASURE = 0, STREAM_SCENARIO = 1 };

// A sequence of uniform doubles in [0,1).  Draw k is words (2(k&1), 2(k&1)+1)
// of Philox block k>>1 under counter (block, agent|stream<<8|tick<<16, episode,
// env_lo) and key (seed_lo, seed_hi ^ env_hi); or, in the injected test mode,
// element k of the caller's table.
struct Draws {
  Philox ph;
  const double* inj;
  int inj_k;
  int k;
  bool overflow;
  uint32_t cache[4];
  int cached_block;

  __device__ __forceinline__ void init(uint64_t seed, int64_t gid, int stream, int agent, int tick,
                                       int episode) {
    ph.c1 = (uint32_t)agent | ((uint32_t)stream << 8) | ((uint32_t)tick << 16);
    ph.c2 = (uint32_t)episode;
    ph.c3 = (uint32_t)((uint64_t)gid & 0xffffffffull);
    ph.k0 = (uint32_t)(seed & 0xffffffffull);
    ph.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)((uint64_t)gid >> 32);
    inj = nullptr;
    inj_k = 0;
    k = 0;
    overflow = false;
    cached_block = -1;
  }

  __device__ __forceinline__ double next() {
    const int i = k++;
    if (inj != nullptr) {
      if (i >= inj_k) {
        overflow = true;
        return 0.5;
      }
      return inj[i];
    }
    const int blk = i >> 1;
    if (blk != cached_block) {
      ph.block((uint32_t)blk, cache);
      cached_block = blk;
    }
    const uint32_t lo = (i & 1) ? cache[2] : cache[0];
    const uint32_t hi = (i & 1) ? cache[3] : cache[1];
    const uint64_t bits = ((uint64_t)hi << 32) | lo;
    return __dmul_rn((double)(bits >> 11), 1.0 / 9007199254740992.0);
  }

  // Two consecutive draws.  From an even position of the Philox sequence they are the two halves of ONE block:
  // a straight-line path without the block cache (PTRS consumes its uniforms in exactly such pairs); any other
  // case — injected table, odd position — goes through next().  Same values either way.
  __device__ __forceinline__ void next_pair(double& u0, double& u1) {
    if (inj == nullptr && (k & 1) == 0) {
      uint32_t w[4];
      ph.block((uint32_t)(k >> 1), w);
      k += 2;
      u0 = __dmul_rn((double)((((uint64_t)w[1] << 32) | w[0]) >> 11), 1.0 / 9007199254740992.0);
      u1 = __dmul_rn((double)((((uint64_t)w[3] << 32) | w[2]) >> 11), 1.0 / 9007199254740992.0);
    } else {
      u0 = next();
      u1 = next();
    }
  }
};

// ---------------------------------------------------------------------------
// Deterministic natural log / exp / log-gamma.
__device__ __forceinline__ double dlog(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  if (x == 0.0) return -CUDART_INF;
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  int k = 0;
  if ((b >> 52) == 0ull) {
    x = __dmul_rn(x, 18014398509481984.0);
    b = (unsigned long long)__double_as_longlong(x);
    k = -54;
  }
  k += (int)((b >> 52) & 0x7ffull) - 1023;
  const unsigned long long mant = b & 0x000fffffffffffffull;
  if (mant >= 0x6a09e667f3bcdull) {
    k += 1;
    b = mant | 0x3fe0000000000000ull;
  } else {
    b = mant | 0x3ff0000000000000ull;
  }
  const double f = __dadd_rn(__longlong_as_double((long long)b), -1.0);
  const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
  const double z = __dmul_rn(s, s);
  const double w = __dmul_rn(z, z);
  const double t1 = __dmul_rn(w, __dadd_rn(Lg2, __dmul_rn(w, __dadd_rn(Lg4, __dmul_rn(w, Lg6)))));
  const double t2 = __dmul_rn(
      z, __dadd_rn(Lg1, __dmul_rn(w, __dadd_rn(Lg3, __dmul_rn(w, __dadd_rn(Lg5, __dmul_rn(w, Lg7)))))));
  const double R = __dadd_rn(t2, t1);
  const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
  const double dk = (double)k;
  // dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f)
  const double inner = __dadd_rn(__dmul_rn(s, __dadd_rn(hfsq, R)), __dmul_rn(dk, ln2_lo));
  return __dsub_rn(__dmul_rn(dk, ln2_hi), __dsub_rn(__dsub_rn(hfsq, inner), f));
}

// domain: x <= 0 (the sampler only needs exp(-lambda), lambda in (0, 10))
__device__ __forceinline__ double dexp(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
               P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x == 0.0) return 1.0;
  int k = (int)__dsub_rn(__dmul_rn(invln2, x), 0.5);  // truncation toward zero
  const double dk = (double)k;
  const double hi = __dsub_rn(x, __dmul_rn(dk, ln2_hi));
  const double lo = __dmul_rn(dk, ln2_lo);
  const double r = __dsub_rn(hi, lo);
  const double t = __dmul_rn(r, r);
  double p = __dadd_rn(P4, __dmul_rn(t, P5));
  p = __dadd_rn(P3, __dmul_rn(t, p));
  p = __dadd_rn(P2, __dmul_rn(t, p));
  p = __dadd_rn(P1, __dmul_rn(t, p));
  const double c = __dsub_rn(r, __dmul_rn(t, p));
  double y = __dsub_rn(1.0, __dsub_rn(__dsub_rn(lo, __ddiv_rn(__dmul_rn(r, c), __dsub_rn(2.0, c))), hi));
  if (k < -1000) {
    y = __dmul_rn(y, __longlong_as_double((long long)(1023 - 1000) << 52));
    k += 1000;
  }
  return __dmul_rn(y, __longlong_as_double((long long)(1023 + k) << 52));
}

__device__ __forceinline__ double dloggam(double x) {
  const double a[10] = {8.333333333333333e-02,  -2.777777777777778e-03, 7.936507936507937e-04,
                        -5.952380952380952e-04, 8.417508417508418e-04,  -1.917526917526918e-03,
                        6.410256410256410e-03,  -2.955065359477124e-02, 1.796443723688307e-01,
                        -1.39243221690590e+00};
  if (x == 1.0 || x == 2.0) return 0.0;
  long long n = 0;
  if (x < 7.0) n = (long long)__dsub_rn(7.0, x);
  double x0 = __dadd_rn(x, (double)n);
  const double inv = __ddiv_rn(1.0, x0);
  const double x2 = __dmul_rn(inv, inv);
  double gl0 = a[9];
#pragma unroll
  for (int k = 8; k >= 0; --k) {
    gl0 = __dmul_rn(gl0, x2);
    gl0 = __dadd_rn(gl0, a[k]);
  }
  // gl0/x0 + 0.5*lg2pi + (x0-0.5)*log(x0) - x0, left to right
  double gl = __dadd_rn(__ddiv_rn(gl0, x0), __dmul_rn(0.5, 1.8378770664093453e+00));
  gl = __dadd_rn(gl, __dmul_rn(__dsub_rn(x0, 0.5), dlog(x0)));
  gl = __dsub_rn(gl, x0);
  if (x < 7.0) {
    for (long long k = 1; k <= n; ++k) {
      gl = __dsub_rn(gl, dlog(__dsub_rn(x0, 1.0)));
      x0 = __dsub_rn(x0, 1.0);
    }
  }
  return gl;
}

// ---------------------------------------------------------------------------
// Squeeze for the PTRS acceptance test  log V + log(1/alpha) - log(a/us^2 + b) <= -lam + k log lam - lgamma(k+1).
// The exact test (four dlog, the log-gamma series, three divisions: a ~450-instruction dependent fp64 chain run
// by a handful of lanes while the rest of the warp waits) was a quarter of k_step's time.  Here both sides are
// evaluated APPROXIMATELY — single-precision lg2.approx for the logarithms, three Stirling terms for lgamma —
// together with a rigorous bound `margin` on |approximate difference - the difference the exact code computes|.
// Only when the approximate difference clears the margin is the decision taken from it (it is then the same
// decision, on any hardware); otherwise the caller evaluates the exact test.  Returns +1 accept, -1 reject,
// 0 undecided.
//
// Error budget (natural-log units).  flog(x) = ln2 * lg2.approx((float)x): rounding x to float moves ln x by
// <= 2^-24; lg2.approx.f32 is within 2^-22 absolute on [0.5, 2] and 2 ulp (<= 2^-22 |result|) elsewhere, so
// |flog(x) - ln x| <= 2^-21 (1 + |ln x|).  The code charges c = 2^-20 per unit and then multiplies the whole
// budget by 4.  The float argument a/us^2 + b carries seven roundings (<= 7 * 2^-24 in its logarithm, which is
// >= 2.2: still inside 2^-21 (1 + |ln|)).  ln(k+1) is taken as loglam (the exact test's own value) + flog((k+1)/lam)
// with the quotient formed in float (three roundings); it enters lgamma multiplied by (k + 1/2), so its error is
// charged x-fold.  Stirling's series cut after 1/(360 x^3) is within 1/(1260 x^5) < 2.5e-8 for x >= 8; below
// that ln k! comes from a table.  The fp64 evaluation error of either side (exact code or this one) is below
// 2^-45 of the sum of the magnitudes of its terms; 2^-40 is charged.
// (tests/test_squeeze_margin.py replays this budget on the CPU against an adversarial logarithm.)
__device__ __forceinline__ float flog2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ int ptrs_squeeze(double V, double us, double a, double b, double invalpha, double lam,
                                            double loglam, long long k) {
  const double ln2 = 6.93147180559945286227e-01, c = 9.5367431640625e-07 /* 2^-20 */;
  const double x = (double)(k + 1);
  const float usf = (float)us;
  const float wf = __fadd_rn(__fdiv_rn((float)a, __fmul_rn(usf, usf)), (float)b);
  const double Lv = __dmul_rn(ln2, (double)flog2_approx((float)V));
  const double Li = __dmul_rn(ln2, (double)flog2_approx((float)invalpha));
  const double Lw = __dmul_rn(ln2, (double)flog2_approx(wf));
  double G, gerr, gmag;  // lgamma(k + 1), the bound on its error (before the factor 4), its magnitude terms
  if (k < 7) {
    // ln k! for k = 0..6, correctly rounded
    const double lf = k < 2 ? 0.0 : k == 2 ? 6.93147180559945286227e-01 : k == 3 ? 1.79175946922805495731e+00
                    : k == 4 ? 3.17805383034794575181e+00 : k == 5 ? 4.78749174278204581157e+00
                             : 6.57925121201010121297e+00;
    G = lf;
    gerr = 0.0;
    gmag = lf;
  } else {
    // ln x = ln lam (the exact test's own value) + ln(x / lam), the quotient formed in float
    const double Lq = __dmul_rn(ln2, (double)flog2_approx(__fdiv_rn((float)x, (float)lam)));
    const double Lx = __dadd_rn(loglam, Lq);
    // lgamma(x) ~ (x - 1/2) ln x - x + ln(2 pi)/2 + 1/(12 x) - 1/(360 x^3)
    const double ix = (double)__frcp_rn((float)x);
    const double ix3 = __dmul_rn(__dmul_rn(ix, ix), ix);
    G = __dsub_rn(__dmul_rn(__dsub_rn(x, 0.5), Lx), x);
    G = __dadd_rn(G, 9.18938533204672669541e-01);
    G = __dadd_rn(G, __dsub_rn(__dmul_rn(ix, 8.33333333333333287074e-02), __dmul_rn(ix3, 2.77777777777777788381e-03)));
    gerr = __dmul_rn(__dmul_rn(c, x), __dadd_rn(1.0, fabs(Lq)));
    gmag = __dadd_rn(__dmul_rn(x, fabs(Lx)), x);
  }
  const double kl = __dmul_rn((double)k, loglam);
  const double lhs = __dsub_rn(__dadd_rn(Lv, Li), Lw);
  const double rhs = __dsub_rn(__dsub_rn(kl, lam), G);
  const double diff = __dsub_rn(lhs, rhs);
  const double logs = __dadd_rn(__dadd_rn(fabs(Lv), fabs(Li)), __dadd_rn(fabs(Lw), 3.0));
  const double mags = __dadd_rn(__dadd_rn(lam, fabs(kl)), __dadd_rn(gmag, logs));
  const double margin =
      __dadd_rn(__dmul_rn(4.0, __dadd_rn(__dadd_rn(__dmul_rn(c, logs), gerr), 1.1920928955078125e-07 /* 2^-23 */)),
                __dmul_rn(9.094947017729282e-13 /* 2^-40 */, mags));
  if (!(margin < 1.0)) return 0;  // huge rates, or anything not finite: the exact test decides
  if (diff < -margin) return 1;
  if (diff > margin) return -1;
  return 0;
}

// ---------------------------------------------------------------------------
// Poisson sampler: Knuth multiplication below 10, Hoermann's PTRS from 10 up —
// the algorithm numpy.random.Generator.poisson uses, on our uniforms.
constexpr int kPoissonMaxIters = 1024;

__device__ __forceinline__ long long poisson_draw(double lam, Draws& d, int& flags) {
  if (!(lam >= 0.0) || lam > 1073741824.0) {
    flags |= RT_FLAG_LAMBDA;
    return 0;
  }
  if (lam == 0.0) return 0;
  if (lam < 10.0) {
    const double enlam = dexp(-lam);
    long long X = 0;
    double prod = 1.0;
    for (int it = 0; it < kPoissonMaxIters; ++it) {
      const double U = d.next();
      if (d.overflow) break;
      prod = __dmul_rn(prod, U);
      if (prod > enlam)
        X += 1;
      else
        return X;
    }
    flags |= RT_FLAG_DRAWS;
    return X;
  }
  const double slam = __dsqrt_rn(lam);
  const double loglam = dlog(lam);
  const double b = __dadd_rn(0.931, __dmul_rn(2.53, slam));
  const double a = __dadd_rn(-0.059, __dmul_rn(0.02483, b));
  const double invalpha = __dadd_rn(1.1239, __ddiv_rn(1.1328, __dsub_rn(b, 3.4)));
  const double vr = __dsub_rn(0.9277, __ddiv_rn(3.6224, __dsub_rn(b, 2.0)));
  long long k = 0;
  for (int it = 0; it < kPoissonMaxIters; ++it) {
    double u0, V;
    d.next_pair(u0, V);
    if (d.overflow) break;
    const double U = __dsub_rn(u0, 0.5);
    const double us = __dsub_rn(0.5, fabs(U));
    // floor((2*a/us + b)*U + lam + 0.43)
    const double t = __dadd_rn(__ddiv_rn(__dmul_rn(2.0, a), us), b);
    k = (long long)floor(__dadd_rn(__dadd_rn(__dmul_rn(t, U), lam), 0.43));
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0 || (us < 0.013 && V > us)) continue;
#ifndef RT_PTRS_NO_SQUEEZE
    {  // cheap outer test first: decides all but ~1e-4 of the cases that get here, never differently
      const int sq = ptrs_squeeze(V, us, a, b, invalpha, lam, loglam, k);
      if (sq > 0) return k;
      if (sq < 0) continue;
    }
#endif
    const double lhs = __dsub_rn(__dadd_rn(dlog(V), dlog(invalpha)),
                                 dlog(__dadd_rn(__ddiv_rn(a, __dmul_rn(us, us)), b)));
    const double rhs = __dsub_rn(__dadd_rn(-lam, __dmul_rn((double)k, loglam)),
                                 dloggam(__dadd_rn((double)k, 1.0)));
    if (lhs <= rhs) return k;
  }
  flags |= RT_FLAG_DRAWS;
  return k < 0 ? 0 : k;
}

// ---------------------------------------------------------------------------
// Line of sight: proper-crossing test with the half-open rule.
__device__ __forceinline__ float orient(float ax, float ay, float bx, float by, float cx, float cy) {
  return __fsub_rn(__fmul_rn(__fsub_rn(bx, ax), __fsub_rn(cy, ay)),
                   __fmul_rn(__fsub_rn(by, ay), __fsub_rn(cx, ax)));
}
__device__ __forceinline__ bool seg_hits_edge(float px, float py, float qx, float qy, float4 e) {
  const float d1 = orient(e.x, e.y, e.z, e.w, px, py);
  const float d2 = orient(e.x, e.y, e.z, e.w, qx, qy);
  const float d3 = orient(px, py, qx, qy, e.x, e.y);
  const float d4 = orient(px, py, qx, qy, e.z, e.w);
  return ((d1 > 0.0f) != (d2 > 0.0f)) && ((d3 > 0.0f) != (d4 > 0.0f));
}
// number of polygons (groups of 4 edge slots, here in shared memory) crossed
__device__ __forceinline__ int polys_crossed(const float4* __restrict__ edges, int n_poly, float px,
                                             float py, float qx, float qy) {
  int n = 0;
  for (int p = 0; p < n_poly; ++p) {
    bool hit = false;
#pragma unroll
    for (int j = 0; j < RT_EDGES_PER_POLY; ++j)
      hit |= seg_hits_edge(px, py, qx, qy, edges[RT_EDGES_PER_POLY * p + j]);
    n += hit ? 1 : 0;
  }
  return n;
}

// ---------------------------------------------------------------------------
// WelfordsOnlineStandardization state (standardize.py:17-31) and recurrences.
struct WelfordState {
  double mean, m2, var, sd, zmax, zmin;
  int count;
};

__device__ __forceinline__ void welford_reset(WelfordState& w) {
  w.mean = 0.0; w.m2 = 0.0; w.var = 0.0; w.sd = 1.0; w.zmax = 0.0; w.zmin = 0.0; w.count = 0;
}
__device__ __forceinline__ double welford_standardize(const WelfordState& w, double r) {
  return __ddiv_rn(__dsub_rn(r, w.mean), w.sd);  // standardize.py:93
}
// standardize.py:46-80; caller has checked reading >= 0.  Returns z of the reading.
__device__ __forceinline__ double welford_update(WelfordState& w, double r) {
  w.count += 1;
  if (w.count == 1) {
    w.mean = r;
  } else {
    const double delta = __dsub_rn(r, w.mean);
    const double mean_new = __dadd_rn(w.mean, __ddiv_rn(delta, (double)w.count));
    const double m2_new = __dadd_rn(w.m2, __dmul_rn(delta, __dsub_rn(r, mean_new)));
    w.mean = mean_new;
    w.m2 = m2_new;
    w.var = __ddiv_rn(m2_new, (double)(w.count - 1));
    const double sd = __dsqrt_rn(w.var);
    w.sd = sd > 1.0 ? sd : 1.0;
  }
  const double z = welford_standardize(w, r);
  if (z > w.zmax) w.zmax = z;
  if (z < w.zmin) w.zmin = z;
  return z;
}

// Normalizer.normalize (normalize.py:25-55).  Returns false where the reference asserts.
__device__ __forceinline__ bool normalize_minmax(double cur, double mx, bool has_min, double mn,
                                                 double& out) {
  out = CUDART_NAN;
  if (cur == 0.0) { out = 0.0; return true; }
  if (!(mx >= cur)) return false;
  double offset;
  if (has_min && mn != 0.0) {
    if (!(cur >= mn)) return false;
    offset = mn < 0.0 ? fabs(mn) : 0.0;
  } else {
    mn = 0.0;
    if (cur < 0.0) { out = 0.0; return true; }
    offset = 0.0;
  }
  if (!(__dadd_rn(mx, offset) > 0.0)) return false;
  const double lo = __dadd_rn(mn, offset);
  const double r = __ddiv_rn(__dsub_rn(__dadd_rn(cur, offset), lo), __dsub_rn(__dadd_rn(mx, offset), lo));
  if (!(r >= 0.0 && r <= 1.0)) return false;
  out = r;
  return true;
}

// estimator.py:45-48 — running max/min of the estimate, 0 meaning "unset"
__device__ __forceinline__ void est_maxmin(double est, double& mx, double& mn) {
  if (est > mx || mx == 0.0) mx = est;
  if (est < mn || mn == 0.0) mn = est;
}

}  // namespace rt
