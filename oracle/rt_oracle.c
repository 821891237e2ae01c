/*
 * rt_oracle.c — CPU ORACLE for the RAD-TEAM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call this file.  The product
 * (rad_team_b200/, librt_b200.so) never touches it.
 *
 * It is a scalar, one-env-at-a-time restatement in plain C of
 *   (a) the reference's own code for the pieces the reference has
 *       [PARITY PINNED: tests/test_oracle_*.py check every function below
 *       against the reference's golden vectors and against fixtures generated
 *       by importing the unmodified reference (tests/golden/make_golden.py)]:
 *         a1/a4  motion, clip, L1 reward/done  src/envs/simple_gridworld.py:15-31,131-167
 *         a3     reset                          src/envs/simple_gridworld.py:114-129
 *         a6     Welford standardiser           src/math_utils/standardize.py:36-93
 *         a7     min-max Normalizer             src/math_utils/normalize.py:25-55
 *         a8     BensLogNormalizer              src/math_utils/normalize.py:79-115
 *         a9     IntensityResamplingEstimator   src/math_utils/estimator.py:31-70
 *   (b) the physics the mounted reference snapshot does NOT contain
 *       [PARITY UNPINNED by the reference — SURVEY.md §0/§8c: no code, no test;
 *       the authority is the written spec in DESIGN.md §3, restated here]:
 *         N1 inverse-square + background, N2 line-of-sight ray-vs-edge tests,
 *         N3 Poisson sampling (the published algorithm numpy.random.Generator
 *            uses: Knuth multiplication below 10, Hoermann's PTRS from 10 up,
 *            with the 10-term log-gamma series; cross-checked draw-for-draw
 *            against NumPy in tests/test_oracle_physics.py),
 *         N4 collision with obstructions, N5 heat-map layout,
 *         scenario generation, Philox4x32-10 (Random123 known answers).
 *
 * Data layout here is deliberately NOT the GPU's: one struct per env, a flat
 * time-ordered reading log that is scanned and qsort-ed for every median.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 * All floating point must stay un-contracted: the GPU side is compiled with
 * -fmad=false and the two must agree bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_AGENTS 8
#define ORC_MAX_POLYS 5
#define ORC_MAX_EDGES 20

#define FLAG_BAD_ACTION 0x01
#define FLAG_LOG_FULL 0x02
#define FLAG_LAMBDA 0x04
#define FLAG_DRAWS 0x08
#define FLAG_NORMALIZE 0x10

/* Same field order as rt_env_cfg in include/rt_b200.h so one ctypes struct
 * serves both sides of a parity test. */
typedef struct orc_cfg {
  int32_t n_envs, n_agents, grid, max_steps;
  int32_t poly_min, poly_max, rect_min, rect_max;
  int32_t lognorm_step, fixed_scenario;
  int64_t env_id_base;
  uint64_t seed;
  double cell_pitch_cm, min_dist_cm, transmission;
  double src_lo, src_hi, bkg_lo, bkg_hi;
} orc_cfg;

typedef struct orc_welford {
  int32_t count;
  double mean, m2, var, std, zmax, zmin;
} orc_welford;

typedef struct orc_one {
  int32_t src[2];
  double src_int, bkg;
  int32_t n_poly;
  float edges[ORC_MAX_EDGES][4];
  int32_t agent[ORC_MAX_AGENTS][2], start[ORC_MAX_AGENTS][2], prev_dist[ORC_MAX_AGENTS];
  int32_t tick, episode, flags;
  orc_welford w;
  double e_max, e_min;
  uint16_t* visit;
  float* inten;
  int32_t n_log;
  uint16_t* log_cell;
  int32_t* log_val;
  /* last-step diagnostics */
  int32_t last_count[ORC_MAX_AGENTS];
  uint8_t last_los[ORC_MAX_AGENTS];
  double last_lambda[ORC_MAX_AGENTS], last_est[ORC_MAX_AGENTS], last_z[ORC_MAX_AGENTS],
      last_value[ORC_MAX_AGENTS];
} orc_one;

typedef struct orc_env {
  orc_cfg cfg;
  orc_one* e;
  float* lut; /* [T*A+1] */
  const double* inj;
  int32_t inj_k;
  int threads;
} orc_env;

/* ------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11).                                      */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* a draw sequence is (stream, global env, episode, tick, agent);
 * draw k lives in Philox block k>>1, words (2(k&1), 2(k&1)+1); the double is
 * the top 53 bits of hi:lo times 2^-53. */
typedef struct draws {
  const double* inj; /* injected uniforms, or NULL */
  int32_t inj_k;
  uint32_t ctr1, ctr2, ctr3, key0, key1;
  int32_t k;
  int overflow;
} draws;

static void draws_init(draws* d, const orc_env* env, int64_t gid, int stream, int agent, int tick,
                       int episode) {
  d->inj = NULL;
  d->inj_k = 0;
  d->ctr1 = (uint32_t)agent | ((uint32_t)stream << 8) | ((uint32_t)tick << 16);
  d->ctr2 = (uint32_t)episode;
  d->ctr3 = (uint32_t)((uint64_t)gid & 0xffffffffu);
  d->key0 = (uint32_t)(env->cfg.seed & 0xffffffffu);
  d->key1 = (uint32_t)(env->cfg.seed >> 32) ^ (uint32_t)((uint64_t)gid >> 32);
  d->k = 0;
  d->overflow = 0;
}

static double draws_next(draws* d) {
  int k = d->k++;
  if (d->inj) {
    if (k >= d->inj_k) {
      d->overflow = 1;
      return 0.5;
    }
    return d->inj[k];
  }
  uint32_t ctr[4] = {(uint32_t)(k >> 1), d->ctr1, d->ctr2, d->ctr3};
  uint32_t key[2] = {d->key0, d->key1};
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  uint32_t lo = o[2 * (k & 1)], hi = o[2 * (k & 1) + 1];
  uint64_t bits = ((uint64_t)hi << 32) | lo;
  return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------------ */
/* Deterministic log / exp.  The classic fdlibm argument
 * reductions and minimax polynomials, evaluated with plain IEEE + - * / in a
 * fixed order so the CUDA kernel (compiled -fmad=false) returns the same bits.
 * Domain: finite x > 0 (and x == 0 -> -inf) for log; x in [-745, 0] for exp.  */
static double bits_to_double(uint64_t b) {
  double d;
  memcpy(&d, &b, 8);
  return d;
}
static uint64_t double_to_bits(double d) {
  uint64_t b;
  memcpy(&b, &d, 8);
  return b;
}

double orc_log(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                      Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                      Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                      Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                      Lg7 = 1.479819860511658591e-01;
  if (x == 0.0) return -INFINITY;
  uint64_t b = double_to_bits(x);
  int k = 0;
  if ((b >> 52) == 0) { /* subnormal: scale by 2^54 */
    x = x * 18014398509481984.0;
    b = double_to_bits(x);
    k = -54;
  }
  k += (int)((b >> 52) & 0x7ff) - 1023;
  uint64_t mant = b & 0x000fffffffffffffull;
  /* m in [1,2); fold to [sqrt(2)/2, sqrt(2)) */
  if (mant >= 0x6a09e667f3bcdull) { /* m >= sqrt(2) (truncated) */
    k += 1;
    b = mant | 0x3fe0000000000000ull; /* m/2 */
  } else {
    b = mant | 0x3ff0000000000000ull;
  }
  double f = bits_to_double(b) - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double hfsq = 0.5 * f * f;
  double dk = (double)k;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

double orc_exp(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                      invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
                      P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                      P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x == 0.0) return 1.0;
  /* k = round-half-away(x / ln2) for x <= 0 */
  int k = (int)(invln2 * x - 0.5);
  double dk = (double)k;
  double hi = x - dk * ln2_hi;
  double lo = dk * ln2_lo;
  double r = hi - lo;
  double t = r * r;
  double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  /* scale by 2^k, k in [-1075, 0]: two exact power-of-two multiplies */
  if (k < -1000) {
    y = y * bits_to_double((uint64_t)(1023 - 1000) << 52);
    k += 1000;
  }
  return y * bits_to_double((uint64_t)(1023 + k) << 52);
}

/* log-gamma by the 10-term asymptotic series with upward shift to x >= 7. */
double orc_loggam(double x) {
  static const double a[10] = {8.333333333333333e-02,  -2.777777777777778e-03, 7.936507936507937e-04,
                               -5.952380952380952e-04, 8.417508417508418e-04,  -1.917526917526918e-03,
                               6.410256410256410e-03,  -2.955065359477124e-02, 1.796443723688307e-01,
                               -1.39243221690590e+00};
  if (x == 1.0 || x == 2.0) return 0.0;
  int64_t n = 0;
  if (x < 7.0) n = (int64_t)(7.0 - x);
  double x0 = x + (double)n;
  double x2 = (1.0 / x0) * (1.0 / x0);
  double gl0 = a[9];
  for (int k = 8; k >= 0; --k) {
    gl0 = gl0 * x2;
    gl0 = gl0 + a[k];
  }
  double gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * orc_log(x0) - x0;
  if (x < 7.0) {
    for (int64_t k = 1; k <= n; ++k) {
      gl = gl - orc_log(x0 - 1.0);
      x0 = x0 - 1.0;
    }
  }
  return gl;
}

#define POISSON_MAX_ITERS 1024

static int64_t poisson_draw(double lam, draws* d, int* flags) {
  if (!(lam >= 0.0) || lam > 1073741824.0) {
    *flags |= FLAG_LAMBDA;
    return 0;
  }
  if (lam == 0.0) return 0;
  if (lam < 10.0) { /* multiplication method */
    double enlam = orc_exp(-lam);
    int64_t X = 0;
    double prod = 1.0;
    for (int it = 0; it < POISSON_MAX_ITERS; ++it) {
      double U = draws_next(d);
      if (d->overflow) break;
      prod = prod * U;
      if (prod > enlam)
        X += 1;
      else
        return X;
    }
    *flags |= FLAG_DRAWS;
    return X;
  }
  /* transformed rejection with squeeze (PTRS) */
  double slam = sqrt(lam);
  double loglam = orc_log(lam);
  double b = 0.931 + 2.53 * slam;
  double a = -0.059 + 0.02483 * b;
  double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  double vr = 0.9277 - 3.6224 / (b - 2.0);
  int64_t k = 0;
  for (int it = 0; it < POISSON_MAX_ITERS; ++it) {
    double U = draws_next(d) - 0.5;
    double V = draws_next(d);
    if (d->overflow) break;
    double us = 0.5 - fabs(U);
    k = (int64_t)floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0 || (us < 0.013 && V > us)) continue;
    double lhs = orc_log(V) + orc_log(invalpha) - orc_log(a / (us * us) + b);
    double rhs = -lam + (double)k * loglam - orc_loggam((double)k + 1.0);
    if (lhs <= rhs) return k;
  }
  *flags |= FLAG_DRAWS;
  return k < 0 ? 0 : k;
}

/* Standalone sampler for unit tests: consume u[0..n_u) in order. */
int64_t orc_poisson(double lam, const double* u, int32_t n_u, int32_t* used, int32_t* flags) {
  draws d;
  memset(&d, 0, sizeof d);
  d.inj = u;
  d.inj_k = n_u;
  int f = 0;
  int64_t r = poisson_draw(lam, &d, &f);
  if (used) *used = d.k;
  if (flags) *flags = f;
  return r;
}

/* Philox-backed uniforms for unit tests (stream, gid, episode, tick, agent). */
void orc_uniforms(uint64_t seed, int64_t gid, int32_t stream, int32_t agent, int32_t tick,
                  int32_t episode, int32_t n, double* out) {
  orc_env tmp;
  memset(&tmp, 0, sizeof tmp);
  tmp.cfg.seed = seed;
  draws d;
  draws_init(&d, &tmp, gid, stream, agent, tick, episode);
  for (int i = 0; i < n; ++i) out[i] = draws_next(&d);
}

/* ------------------------------------------------------------------------ */
/* a6  WelfordsOnlineStandardization (standardize.py:36-93)                   */
void orc_welford_reset(orc_welford* w) {
  w->count = 0;
  w->mean = 0.0;
  w->m2 = 0.0;
  w->var = 0.0;
  w->std = 1.0;
  w->zmax = 0.0;
  w->zmin = 0.0;
}

double orc_welford_standardize(const orc_welford* w, double reading) {
  return (reading - w->mean) / w->std; /* standardize.py:93 */
}

/* returns 0, or 1 where the reference's `assert reading >= 0` fires (:63) */
int orc_welford_update(orc_welford* w, double reading) {
  if (!(reading >= 0.0)) return 1;
  w->count += 1;
  if (w->count == 1) {
    w->mean = reading; /* :66-67 */
  } else {
    double delta = reading - w->mean;
    double mean_new = w->mean + delta / (double)w->count;             /* :69 */
    double m2_new = w->m2 + (reading - w->mean) * (reading - mean_new); /* :70 */
    w->mean = mean_new;
    w->m2 = m2_new;
    w->var = m2_new / (double)(w->count - 1); /* :73 */
    double sd = sqrt(w->var);
    w->std = sd > 1.0 ? sd : 1.0; /* max(sqrt(var), 1)  :74 */
  }
  double z = orc_welford_standardize(w, reading); /* :76 */
  if (z > w->zmax) w->zmax = z;
  if (z < w->zmin) w->zmin = z;
  return 0;
}

/* a7  Normalizer.normalize (normalize.py:25-55).  has_min == 0 <=> min=None.
 * Returns 0 ok / 1 assertion; *out is the reference's return value. */
int orc_normalize(double cur, double max, int has_min, double min, double* out) {
  *out = NAN;
  if (cur == 0.0) { /* :33-34 */
    *out = 0.0;
    return 0;
  }
  if (!(max >= cur)) return 1; /* :35 */
  double offset;
  if (has_min && min != 0.0) { /* `if min:` :39 — 0 and None are both falsy */
    if (!(cur >= min)) return 1; /* :40 */
    offset = min < 0.0 ? fabs(min) : 0.0;
  } else {
    min = 0.0;
    if (cur < 0.0) { /* :47-48 */
      *out = 0.0;
      return 0;
    }
    offset = 0.0;
  }
  if (!(max + offset > 0.0)) return 1; /* :51 */
  double r = ((cur + offset) - (min + offset)) / ((max + offset) - (min + offset)); /* :53 */
  if (!(r >= 0.0 && r <= 1.0)) return 1; /* :54 */
  *out = r;
  return 0;
}

/* a8  BensLogNormalizer.normalize_incremental_logscale (normalize.py:79-115).
 * math.log(x, base) is log(x)/log(base) in CPython; libm log on both. */
int orc_lognormalize(double cur, double max, int32_t step, double* out) {
  *out = NAN;
  if (!(cur >= 0.0 && max > 0.0 && step > 0)) return 1; /* :90 */
  if (max == 1.0) return 2; /* ZeroDivisionError inside math.log(x, 1) */
  double lb = log(max);
  double num = log((double)step + cur) / lb;
  double den = log((double)step * max) / lb;
  double r = num * 1.0 / den; /* :110 */
  if (!(r >= 0.0 && r <= 1.0)) return 1; /* :111-113 */
  *out = r;
  return 0;
}

/* a9  statistics.median of a buffer (estimator.py:70): sort, middle or mean of
 * the middle two. */
static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
double orc_median(const double* v, int32_t n) {
  double* tmp = (double*)malloc(sizeof(double) * (size_t)n);
  memcpy(tmp, v, sizeof(double) * (size_t)n);
  qsort(tmp, (size_t)n, sizeof(double), cmp_double);
  double m = (n & 1) ? tmp[n / 2] : (tmp[n / 2 - 1] + tmp[n / 2]) / 2.0;
  free(tmp);
  return m;
}
/* running max/min of the estimate with the reference's 0-as-unset sentinel
 * (estimator.py:45-48) */
void orc_est_maxmin(double est, double* mx, double* mn) {
  if (est > *mx || *mx == 0.0) *mx = est;
  if (est < *mn || *mn == 0.0) *mn = est;
}

/* Sequential batch moments: the reference recurrence over a float buffer,
 * state = {count, mean, M2}. */
void orc_moments_update(double state[3], const float* x, int64_t n) {
  double cnt = state[0], mean = state[1], m2 = state[2];
  for (int64_t i = 0; i < n; ++i) {
    double r = (double)x[i];
    cnt += 1.0;
    if (cnt == 1.0) {
      mean = r;
    } else {
      double mean_new = mean + (r - mean) / cnt;
      m2 = m2 + (r - mean) * (r - mean_new);
      mean = mean_new;
    }
  }
  state[0] = cnt; state[1] = mean; state[2] = m2;
}
/* Chan et al. pairwise merge, a <- a (+) b. */
void orc_moments_merge(double a[3], const double b[3]) {
  if (b[0] == 0.0) return;
  if (a[0] == 0.0) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; return; }
  double n = a[0] + b[0];
  double delta = b[1] - a[1];
  double mean = a[1] + delta * (b[0] / n);
  double m2 = a[2] + b[2] + delta * delta * (a[0] * (b[0] / n));
  a[0] = n; a[1] = mean; a[2] = m2;
}
void orc_moments_standardize(const double state[3], const float* x, float* out, int64_t n) {
  double sd = 1.0;
  if (state[0] >= 2.0) {
    double s = sqrt(state[2] / (state[0] - 1.0));
    sd = s > 1.0 ? s : 1.0;
  }
  for (int64_t i = 0; i < n; ++i) out[i] = (float)(((double)x[i] - state[1]) / sd);
}

/* ------------------------------------------------------------------------ */
/* N2: segment-vs-edge test, half-open crossing rule, fp32.   */
static float orient(float ax, float ay, float bx, float by, float cx, float cy) {
  float l = (bx - ax) * (cy - ay);
  float r = (by - ay) * (cx - ax);
  return l - r;
}
static int seg_hits_edge(float px, float py, float qx, float qy, const float e[4]) {
  float d1 = orient(e[0], e[1], e[2], e[3], px, py);
  float d2 = orient(e[0], e[1], e[2], e[3], qx, qy);
  float d3 = orient(px, py, qx, qy, e[0], e[1]);
  float d4 = orient(px, py, qx, qy, e[2], e[3]);
  return ((d1 > 0.0f) != (d2 > 0.0f)) && ((d3 > 0.0f) != (d4 > 0.0f));
}
/* number of polygons (groups of 4 edge slots) the segment crosses */
static int polys_crossed(const orc_one* s, float px, float py, float qx, float qy) {
  int n = 0;
  for (int p = 0; p < s->n_poly; ++p) {
    int hit = 0;
    for (int j = 0; j < 4; ++j) hit |= seg_hits_edge(px, py, qx, qy, s->edges[4 * p + j]);
    n += hit;
  }
  return n;
}
int orc_segment_hits(const float seg[4], const float edge[4]) {
  return seg_hits_edge(seg[0], seg[1], seg[2], seg[3], edge);
}
/* Batch form for the exhaustive predicate tests: out[i] = 1 when segment i = (px,py,qx,qy) hits at least one of
 * the four edges of the axis-aligned rectangle (x0,y0,x1,y1), the edges built in the order scenario generation
 * uses — i.e. whether that obstruction counts towards n_blocked. */
void orc_rect_blocks_batch(const float* seg, int64_t n, const float rect[4], uint8_t* out) {
  const float X0 = rect[0], Y0 = rect[1], X1 = rect[2], Y1 = rect[3];
  const float e[4][4] = {{X0, Y0, X1, Y0}, {X1, Y0, X1, Y1}, {X1, Y1, X0, Y1}, {X0, Y1, X0, Y0}};
  for (int64_t i = 0; i < n; ++i) {
    int hit = 0;
    for (int j = 0; j < 4; ++j) hit |= seg_hits_edge(seg[4 * i], seg[4 * i + 1], seg[4 * i + 2], seg[4 * i + 3], e[j]);
    out[i] = (uint8_t)hit;
  }
}

static int l1(const int32_t a[2], const int32_t b[2]) { return abs(a[0] - b[0]) + abs(a[1] - b[1]); }

/* ------------------------------------------------------------------------ */
/* Scenario generation, RNG stream 1.                         */
static int draw_int(draws* d, int n) { /* U{0..n-1} */
  int v = (int)floor(draws_next(d) * (double)n);
  return v >= n ? n - 1 : v;
}

static void gen_scenario(const orc_env* env, orc_one* s, int64_t gid) {
  const orc_cfg* c = &env->cfg;
  const int G = c->grid, A = c->n_agents;
  draws d;
  draws_init(&d, env, gid, 1, 0, 0, s->episode);
  s->src[0] = draw_int(&d, G);
  s->src[1] = draw_int(&d, G);
  s->src_int = c->src_lo + draws_next(&d) * (c->src_hi - c->src_lo);
  s->bkg = c->bkg_lo + draws_next(&d) * (c->bkg_hi - c->bkg_lo);
  for (int a = 0; a < A; ++a) {
    int ok = 0;
    for (int t = 0; t < 8 && !ok; ++t) {
      s->start[a][0] = draw_int(&d, G);
      s->start[a][1] = draw_int(&d, G);
      ok = l1(s->start[a], s->src) > 1;
    }
    if (!ok) {
      s->start[a][0] = (s->src[0] + G / 2) % G;
      s->start[a][1] = (s->src[1] + G / 2) % G;
    }
  }
  memset(s->edges, 0, sizeof s->edges);
  int np = c->poly_min + (c->poly_max > c->poly_min ? draw_int(&d, c->poly_max - c->poly_min + 1) : 0);
  s->n_poly = np;
  for (int p = 0; p < np; ++p) {
    for (int t = 0; t < 8; ++t) {
      int w = c->rect_min + draw_int(&d, c->rect_max - c->rect_min + 1);
      int h = c->rect_min + draw_int(&d, c->rect_max - c->rect_min + 1);
      if (w > G) w = G;
      if (h > G) h = G;
      int x0 = draw_int(&d, G - w + 1);
      int y0 = draw_int(&d, G - h + 1);
      int bad = s->src[0] >= x0 && s->src[0] < x0 + w && s->src[1] >= y0 && s->src[1] < y0 + h;
      for (int a = 0; a < A; ++a)
        bad |= s->start[a][0] >= x0 && s->start[a][0] < x0 + w && s->start[a][1] >= y0 &&
               s->start[a][1] < y0 + h;
      if (bad) continue; /* a polygon that never fits stays an all-zero slot */
      float X0 = (float)x0, Y0 = (float)y0, X1 = (float)(x0 + w), Y1 = (float)(y0 + h);
      float q[4][4] = {{X0, Y0, X1, Y0}, {X1, Y0, X1, Y1}, {X1, Y1, X0, Y1}, {X0, Y1, X0, Y0}};
      memcpy(s->edges[4 * p], q, sizeof q);
      break;
    }
  }
}

/* ------------------------------------------------------------------------ */
static void rasterize_one(const orc_env* env, const orc_one* s, float* maps) {
  const int G = env->cfg.grid, cells = G * G, A = env->cfg.n_agents;
  float* loc = maps;
  float* inten = maps + cells;
  float* vis = maps + 2 * cells;
  memset(loc, 0, sizeof(float) * (size_t)cells);
  for (int a = 0; a < A; ++a) loc[s->agent[a][1] * G + s->agent[a][0]] = 1.0f;
  memcpy(inten, s->inten, sizeof(float) * (size_t)cells);
  for (int i = 0; i < cells; ++i) vis[i] = env->lut[s->visit[i]];
}

static void reset_one(const orc_env* env, orc_one* s, int64_t gid) {
  const int G = env->cfg.grid, A = env->cfg.n_agents;
  s->episode += 1;
  if (!env->cfg.fixed_scenario) gen_scenario(env, s, gid);
  for (int a = 0; a < A; ++a) {
    s->agent[a][0] = s->start[a][0]; /* simple_gridworld.py:120 */
    s->agent[a][1] = s->start[a][1];
    s->prev_dist[a] = l1(s->agent[a], s->src); /* :121 */
    s->last_count[a] = 0;
    s->last_los[a] = 0;
    s->last_lambda[a] = s->last_est[a] = s->last_z[a] = s->last_value[a] = 0.0;
  }
  s->tick = 0;
  s->n_log = 0;
  orc_welford_reset(&s->w);
  s->e_max = s->e_min = 0.0;
  memset(s->visit, 0, sizeof(uint16_t) * (size_t)(G * G));
  memset(s->inten, 0, sizeof(float) * (size_t)(G * G));
}

static void step_one(const orc_env* env, orc_one* s, int64_t gid, int64_t local, const int32_t* act,
                     int32_t* reward, uint8_t* done) {
  static const int DX[5] = {0, 0, 1, -1, 0}, DY[5] = {0, 1, 0, 0, -1}; /* :26-31 */
  const orc_cfg* c = &env->cfg;
  const int G = c->grid, A = c->n_agents;
  for (int a = 0; a < A; ++a)
    if (act[a] < 1 || act[a] > 4) { /* Action(action) raises before any state change (:138) */
      s->flags |= FLAG_BAD_ACTION;
      return;
    }
  if (s->tick >= c->max_steps) {
    s->flags |= FLAG_LOG_FULL;
    return;
  }
  float sx = (float)s->src[0] + 0.5f, sy = (float)s->src[1] + 0.5f;
  for (int a = 0; a < A; ++a) {
    /* a4: motion + clip (:141) */
    int ox = s->agent[a][0], oy = s->agent[a][1];
    int nx = ox + DX[act[a]], ny = oy + DY[act[a]];
    nx = nx < 0 ? 0 : (nx > G - 1 ? G - 1 : nx);
    ny = ny < 0 ? 0 : (ny > G - 1 ? G - 1 : ny);
    /* N4: a move whose centre-to-centre path crosses an obstruction edge is refused */
    if ((nx != ox || ny != oy) &&
        polys_crossed(s, (float)ox + 0.5f, (float)oy + 0.5f, (float)nx + 0.5f, (float)ny + 0.5f) > 0) {
      nx = ox;
      ny = oy;
    }
    s->agent[a][0] = nx;
    s->agent[a][1] = ny;
    /* a4: reward / done on the L1 distance (:144-159) */
    int dist = l1(s->agent[a], s->src);
    if (dist <= 1) {
      done[a] = 1;
      reward[a] = 1;
    } else if (dist < s->prev_dist[a]) {
      done[a] = 0;
      reward[a] = 0;
    } else {
      done[a] = 0;
      reward[a] = -1;
    }
    s->prev_dist[a] = dist;

    /* N2 + N1: line of sight and expected count */
    int nb = polys_crossed(s, (float)nx + 0.5f, (float)ny + 0.5f, sx, sy);
    int ddx = nx - s->src[0], ddy = ny - s->src[1];
    double d2 = (double)(ddx * ddx + ddy * ddy) * (c->cell_pitch_cm * c->cell_pitch_cm);
    double d2min = c->min_dist_cm * c->min_dist_cm;
    if (d2 < d2min) d2 = d2min;
    double t = s->src_int;
    for (int i = 0; i < nb; ++i) t = t * c->transmission;
    double lam = t / d2 + s->bkg;

    /* N3: Poisson count */
    draws d;
    draws_init(&d, env, gid, 0, a, s->tick, s->episode);
    if (env->inj) {
      d.inj = env->inj + ((size_t)local * (size_t)A + (size_t)a) * (size_t)env->inj_k;
      d.inj_k = env->inj_k;
    }
    int64_t cnt = poisson_draw(lam, &d, &s->flags);

    /* a9: estimator.update((x, y), count) and its median (estimator.py:31-48) */
    int cell = ny * G + nx;
    s->log_cell[s->n_log] = (uint16_t)cell;
    s->log_val[s->n_log] = (int32_t)cnt;
    s->n_log += 1;
    double buf[65536 / 64]; /* small stack fast path; falls back to heap below */
    int m = 0;
    for (int i = 0; i < s->n_log; ++i) m += s->log_cell[i] == cell;
    double* vals = m <= (int)(sizeof buf / sizeof buf[0]) ? buf : (double*)malloc(sizeof(double) * (size_t)m);
    m = 0;
    for (int i = 0; i < s->n_log; ++i)
      if (s->log_cell[i] == cell) vals[m++] = (double)s->log_val[i];
    double est = orc_median(vals, m);
    if (vals != buf) free(vals);
    orc_est_maxmin(est, &s->e_max, &s->e_min);
    s->visit[cell] = (uint16_t)(s->visit[cell] + 1);

    /* a6 + a7: Welford update, z-score, min-max against the running z range */
    orc_welford_update(&s->w, est);
    double z = orc_welford_standardize(&s->w, est);
    double val;
    if (orc_normalize(z, s->w.zmax, 1, s->w.zmin, &val)) {
      s->flags |= FLAG_NORMALIZE;
      val = 0.0;
    }
    s->inten[cell] = (float)val;

    s->last_count[a] = (int32_t)cnt;
    s->last_los[a] = (uint8_t)nb;
    s->last_lambda[a] = lam;
    s->last_est[a] = est;
    s->last_z[a] = z;
    s->last_value[a] = val;
  }
  s->tick += 1;
}

/* ------------------------------------------------------------------------ */
orc_env* orc_create(const orc_cfg* cfg) {
  orc_env* env = (orc_env*)calloc(1, sizeof(orc_env));
  env->cfg = *cfg;
  const int G = cfg->grid, cells = G * G, cap = cfg->max_steps * cfg->n_agents;
  env->e = (orc_one*)calloc((size_t)cfg->n_envs, sizeof(orc_one));
  for (int i = 0; i < cfg->n_envs; ++i) {
    orc_one* s = &env->e[i];
    s->visit = (uint16_t*)calloc((size_t)cells, sizeof(uint16_t));
    s->inten = (float*)calloc((size_t)cells, sizeof(float));
    s->log_cell = (uint16_t*)calloc((size_t)cap, sizeof(uint16_t));
    s->log_val = (int32_t*)calloc((size_t)cap, sizeof(int32_t));
    s->episode = -1;
    orc_welford_reset(&s->w);
  }
  env->lut = (float*)calloc((size_t)cap + 1, sizeof(float));
  for (int v = 1; v <= cap; ++v) { /* LUT[0] = 0: unvisited cells stay dark */
    double r;
    orc_lognormalize((double)v, (double)cap, cfg->lognorm_step, &r);
    env->lut[v] = (float)r;
  }
#ifdef _OPENMP
  env->threads = omp_get_max_threads();
#else
  env->threads = 1;
#endif
  return env;
}

void orc_destroy(orc_env* env) {
  if (!env) return;
  for (int i = 0; i < env->cfg.n_envs; ++i) {
    free(env->e[i].visit);
    free(env->e[i].inten);
    free(env->e[i].log_cell);
    free(env->e[i].log_val);
  }
  free(env->e);
  free(env->lut);
  free(env);
}

int orc_set_threads(orc_env* env, int n) {
#ifdef _OPENMP
  if (n < 1) n = omp_get_num_procs();
  env->threads = n;
#else
  env->threads = 1;
#endif
  return env->threads;
}

void orc_set_scenario(orc_env* env, const int32_t* src_xy, const double* src_int, const double* bkg,
                      const int32_t* start_xy, const int32_t* n_poly, const float* edges) {
  const int A = env->cfg.n_agents;
  for (int i = 0; i < env->cfg.n_envs; ++i) {
    orc_one* s = &env->e[i];
    if (src_xy) { s->src[0] = src_xy[2 * i]; s->src[1] = src_xy[2 * i + 1]; }
    if (src_int) s->src_int = src_int[i];
    if (bkg) s->bkg = bkg[i];
    if (start_xy)
      for (int a = 0; a < A; ++a) {
        s->start[a][0] = start_xy[(i * A + a) * 2];
        s->start[a][1] = start_xy[(i * A + a) * 2 + 1];
      }
    if (n_poly) s->n_poly = n_poly[i];
    if (edges) memcpy(s->edges, edges + (size_t)i * ORC_MAX_EDGES * 4, sizeof s->edges);
  }
}

void orc_inject_uniforms(orc_env* env, const double* u, int32_t k) {
  env->inj = u;
  env->inj_k = k;
}

void orc_rasterize(orc_env* env, float* maps) {
  const int cells = env->cfg.grid * env->cfg.grid;
#pragma omp parallel for schedule(static) num_threads(env->threads)
  for (int i = 0; i < env->cfg.n_envs; ++i) rasterize_one(env, &env->e[i], maps + (size_t)i * 3 * cells);
}

void orc_reset(orc_env* env, const uint8_t* mask, int32_t* obs_xy, float* maps) {
  const int A = env->cfg.n_agents, cells = env->cfg.grid * env->cfg.grid;
#pragma omp parallel for schedule(static) num_threads(env->threads)
  for (int i = 0; i < env->cfg.n_envs; ++i) {
    orc_one* s = &env->e[i];
    if (!mask || mask[i]) reset_one(env, s, env->cfg.env_id_base + i);
    if (obs_xy)
      for (int a = 0; a < A; ++a) {
        obs_xy[(i * A + a) * 2] = s->agent[a][0];
        obs_xy[(i * A + a) * 2 + 1] = s->agent[a][1];
      }
    if (maps) rasterize_one(env, s, maps + (size_t)i * 3 * cells);
  }
}

void orc_step(orc_env* env, const int32_t* actions, int32_t* obs_xy, float* maps, int32_t* reward,
              uint8_t* done) {
  const int A = env->cfg.n_agents, cells = env->cfg.grid * env->cfg.grid;
#pragma omp parallel for schedule(static) num_threads(env->threads)
  for (int i = 0; i < env->cfg.n_envs; ++i) {
    orc_one* s = &env->e[i];
    int32_t r[ORC_MAX_AGENTS] = {0};
    uint8_t dn[ORC_MAX_AGENTS] = {0};
    step_one(env, s, env->cfg.env_id_base + i, i, actions + (size_t)i * A, r, dn);
    for (int a = 0; a < A; ++a) {
      if (obs_xy) {
        obs_xy[(i * A + a) * 2] = s->agent[a][0];
        obs_xy[(i * A + a) * 2 + 1] = s->agent[a][1];
      }
      if (reward) reward[i * A + a] = r[a];
      if (done) done[i * A + a] = dn[a];
    }
    if (maps) rasterize_one(env, s, maps + (size_t)i * 3 * cells);
  }
}

int32_t orc_errors(const orc_env* env) {
  int32_t f = 0;
  for (int i = 0; i < env->cfg.n_envs; ++i) f |= env->e[i].flags;
  return f;
}

/* Export one logical state array in the dense layout of rt_buf
 * (include/rt_b200.h); returns bytes written, 0 for an unknown / unsupported id. */
size_t orc_export(const orc_env* env, int which, void* out) {
  const int E = env->cfg.n_envs, A = env->cfg.n_agents, cells = env->cfg.grid * env->cfg.grid;
  const int cap = env->cfg.max_steps * A;
  char* o = (char*)out;
  size_t n = 0;
#define PUT(ptr, bytes) do { memcpy(o + n, (ptr), (bytes)); n += (bytes); } while (0)
  if (which == 25) { PUT(env->lut, sizeof(float) * (size_t)(cap + 1)); return n; }
  for (int i = 0; i < E; ++i) {
    const orc_one* s = &env->e[i];
    switch (which) {
      case 0: PUT(s->agent, sizeof(int32_t) * 2 * (size_t)A); break;
      case 1: PUT(s->prev_dist, sizeof(int32_t) * (size_t)A); break;
      case 2: PUT(s->start, sizeof(int32_t) * 2 * (size_t)A); break;
      case 3: PUT(s->src, sizeof(int32_t) * 2); break;
      case 4: PUT(&s->src_int, 8); break;
      case 5: PUT(&s->bkg, 8); break;
      case 6: PUT(&s->n_poly, 4); break;
      case 7: PUT(s->edges, sizeof s->edges); break;
      case 8: PUT(&s->tick, 4); break;
      case 9: PUT(&s->episode, 4); break;
      case 10: PUT(s->visit, sizeof(uint16_t) * (size_t)cells); break;
      case 11: PUT(s->inten, sizeof(float) * (size_t)cells); break;
      case 15: {
        double w[6] = {s->w.mean, s->w.m2, s->w.var, s->w.std, s->w.zmax, s->w.zmin};
        PUT(w, sizeof w);
      } break;
      case 16: PUT(&s->w.count, 4); break;
      case 17: {
        double m[2] = {s->e_max, s->e_min};
        PUT(m, sizeof m);
      } break;
      case 18: PUT(&s->flags, 4); break;
      case 19: PUT(s->last_count, sizeof(int32_t) * (size_t)A); break;
      case 20: PUT(s->last_los, (size_t)A); break;
      case 21: PUT(s->last_lambda, 8 * (size_t)A); break;
      case 22: PUT(s->last_est, 8 * (size_t)A); break;
      case 23: PUT(s->last_z, 8 * (size_t)A); break;
      case 24: PUT(s->last_value, 8 * (size_t)A); break;
      default: return 0;
    }
  }
#undef PUT
  return n;
}
