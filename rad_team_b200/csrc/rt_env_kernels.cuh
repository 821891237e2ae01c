// rt_env_kernels.cuh — the environment kernels of librt_b200.so (sm_100a).
//
//   k_step               K1+K2+K4a  motion / collision / reward (simple_gridworld.py:131-167), measurement
//                        (inverse square, line of sight against obstruction edges staged in shared memory
//                        by bulk async copies, Philox Poisson), each agent's estimator update
//                        (estimator.py:31-70), then per env the sequential Welford -> normalise chain
//                        (standardize.py:46-93 -> normalize.py:25-55) and the map / record updates.
//   k_raster_sparse_lsu  K3 (default)  visited-cell records -> shared-memory tile -> coalesced 16 B stores
//   k_raster_sparse      K3  the same through the bulk-copy (TMA) engine
//   k_raster_vec / _any  K3  from the dense maps (vectorised / any grid side)
//   k_reset_state, k_zero_maps   K5  masked episode reset + scenario generation.
//
// HBM layout is structure-of-arrays over envs so that neighbouring threads touch neighbouring addresses;
// the pool of reading buckets is time-major [T*A][E] for the same reason.
#pragma once
#include <math_constants.h>

#include "rt_device.cuh"

namespace rt {

struct EnvDev {
  int E, A, G, cells, T, cap;  // cap = T*A log entries per env
  long long gid_base;
  unsigned long long seed;
  double pitch2, dmin2, transmission, src_lo, src_span, bkg_lo, bkg_span;
  int poly_min, poly_max, rect_min, rect_max, fixed;
  int32_t* agent_xy;
  int32_t* prev_dist;
  int32_t* start_xy;
  int32_t* src_xy;
  double* src_int;
  double* bkg;
  int32_t* n_poly;
  float* edges;  // [E][20][4]
  int32_t* tick;
  int32_t* episode;
  uint16_t* visit;  // [E][cells]  dense visit counts: live only when live_dense, else materialised on request
  float* inten;     // [E][cells]  dense intensity map: likewise
  uint16_t* cellrec;  // [E][cells][4]  k_step's random-access state of a cell, ONE 8-byte record per cell so that a
                      //   visit reads one DRAM sector (and a revisit writes none): [0] EPISODE TAG (low 16 bits of the
                      //   env's episode number when the record was written: a record of an older episode counts as
                      //   empty, so a reset clears nothing here except once every 65,536 episodes), [1] first bucket
                      //   of the cell + 1 (0 = none; the visit count travels with that bucket), [2] record index in
                      //   vrec + 1 (0 = not visited), [3] unused
  uint2* vrec;      // [E][rec_stride]  compact list of the VISITED cells, what the sparse rasteriser reads:
                    //   [0] = {n_visited, 0}; [1..2] = the agents' cells, 8 x uint16 (0xFFFF = none);
                    //   [2 + k] = {cell | visits << 16, fp32 bits of the intensity value}, k = 1..n_visited
  int rec_stride;   // uint2 per env (even)
  int rec_prefix;   // uint2 per env moved by the rasteriser's bulk copy (<= rec_stride, even)
  uint4* log;       // [cap][E][2] bucket pool, time-major: a bucket = 7 readings of ONE cell in ascending order +
                    //   link word (low half: id + 1 of the cell's next bucket; high half, first bucket only: readings
                    //   of the cell); bucket id = tick * A + agent of the visit that opened it
  double* welford;  // [E][6]
  int32_t* welford_n;
  double* est_mm;  // [E][2]
  int32_t* flags;
  int32_t* last_count;
  uint8_t* last_los;
  double* last_lambda;
  double* last_est;
  double* last_z;
  double* last_value;
  float* lut;  // [cap+1]
  int32_t* ep_return;  // [E][A]
  uint8_t* ep_done;    // [E][A]
  float* rollout;      // [T][E][A] readings sink, or nullptr
  int32_t* step_flags; // host path, 16 B: word [host_seq & 1] = OR of the RT_FLAG_* that voided an env's step in THIS
                       //   launch, [2..3] = host_seq of the last launch; a launch clears the word of the NEXT call
                       //   (the other one), so no memset sits in front of a step
  unsigned long long host_seq;  // sequence number of the rt_env_step_host call this launch belongs to
  // small-batch host path: k_step writes obs_xy / reward / done straight into mapped host memory and its LAST CTA
  // publishes {flag word, sequence number} there (nullptr: results are copied back by the copy engine instead)
  volatile int32_t* host_flag_out;
  volatile unsigned long long* host_seq_out;
  unsigned* cta_counter;
  int live_dense;      // 1: k_step keeps the dense visit / intensity maps current (dense / generic rasterisers); 0: cell
                       //    records + visited-cell records are the only copy, dense maps are materialised on request
  uint4* prev_cells;   // [E] the agents' cells BEFORE the last step (8 x uint16), for the incremental observation update
  const double* inj;
  int inj_k;
};

// --------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk async copy (the TMA engine without a tensor map; SASS UBLKCP).
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialisation attribute may have
// its CTAs scheduled while the previous kernel in the stream drains; pdl_wait() blocks until that kernel has
// completed and its writes are visible, pdl_launch_dependents() lets the NEXT kernel's CTAs start early in turn.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "RT_WAIT:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra RT_DONE;\n"
      " bra RT_WAIT;\n"
      "RT_DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// --------------------------------------------------------------------------------------------
// K1 + K2 + K4a.  One thread per (env, agent); an env's agents are neighbouring threads of one
// CTA.  blockDim.x = envs_per_cta * A.
constexpr int kEdgeRowF4 = RT_MAX_EDGES + 1;  // 21 float4 = 336 B rows: conflict-free float4 reads

// Per-env state of the sequential chain, staged global -> shared by cp.async at the top of the CTA so that it
// occupies no registers while the per-agent phases run (it used to: 19 registers held across the sampler, the
// point of highest pressure).  Layout of one 96-byte row: [0..47] Welford mean, M2, var, std, zmax, zmin;
// [48..63] estimator max, min; [64..71] record header slot 0 {n_visited, 0}; [72] Welford count; [76] tick.
constexpr int kOwnF4 = 6;
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct StepSmem {  // carved from dynamic shared memory, see step_smem_bytes()
  uint64_t* bar;
  float4* edges;   // [epb][kEdgeRowF4]
  uint4* own;      // [epb][kOwnF4]  the env's sequential statistics, staged by cp.async (see OwnerState)
  double* est;     // [epb*A]  median estimate computed by the agent's own thread
  int* n_poly;     // [epb]
  int* skip;       // [epb]  RT_FLAG_* that void this env's step
  int* eflags;     // [epb]  flags raised by the agents' samplers
  int* cell;       // [epb*A]  the agent's cell after the move
  int* count;      // [epb*A]  Poisson count
  int* slot;       // [epb*A]  prefetched record slot of the cell (0 = none yet)
  int* nread;      // [epb*A]  readings at the cell incl. this one; 0 = deferred to the owner (cell clash)
};
__host__ __device__ inline size_t step_smem_bytes(int epb, int A) {
  return 16 + (size_t)epb * kEdgeRowF4 * 16 + (size_t)epb * kOwnF4 * 16 + (size_t)epb * A * 8 +
         (size_t)epb * 3 * 4 + (size_t)epb * A * 4 * 4;
}

// a9: IntensityResamplingEstimator.update + get_estimate (estimator.py:31-48,61-70) for one cell.
// The cell's readings live in a chain of 32-byte BUCKETS — seven values in ascending order and a link word:
// id (+1) of the next bucket in the low half and, in the chain's FIRST bucket, the number of readings of the
// cell in the high half.  Every bucket of a chain is full except the last, and all values of a bucket are <=
// those of the next.  One DRAM sector therefore holds the whole history of a cell visited up to seven times and
// a visit costs ceil(n / 7) dependent loads (the previous structure, a value-sorted linked list with one
// reading per node, needed ~n / 2); because the visit count travels with the bucket, a revisit leaves the cell
// record untouched — one scattered dirty sector less per agent-step.  Inserting reading `cnt` bubbles it into
// its bucket, cascades the evicted maximum down the chain, and picks the median — elements (n-1)/2 and n/2 of
// the NEW sequence of n readings — on the way.  A chain that needs one more bucket takes slot `idx` = tick * A +
// agent of the time-major pool, so no allocator state exists and a reset has nothing to clear.
// `cur` = first bucket (+1, 0 = none); `cell_rec` = the cell's 8-byte record, `tag` = the episode tag it gets with its
// first bucket; returns the median and, in n_out, the new number of readings.
__device__ __forceinline__ double bucket_insert_median(uint4* __restrict__ pool, size_t E, uint16_t* cell_rec,
                                                       uint32_t tag, int idx, uint32_t cnt, int cur, int& n_out) {
  if (cur == 0) {  // first reading of this cell in this episode
    uint4* np = pool + (size_t)idx * E * 2;
    np[0] = make_uint4(cnt, 0u, 0u, 0u);
    np[1] = make_uint4(0u, 0u, 0u, 1u << 16);
    *reinterpret_cast<uint32_t*>(cell_rec) = tag | ((uint32_t)(idx + 1) << 16);  // episode tag + first bucket
    n_out = 1;
    return (double)cnt;
  }
  uint4* bp = pool + (size_t)(cur - 1) * E * 2;
  uint4 p_lo = bp[0], p_hi = bp[1];
  const int n = (int)(p_hi.w >> 16) + 1;
  n_out = n;
  const int k1 = (n - 1) >> 1, k2 = n >> 1;
  uint32_t m1 = 0u, m2 = 0u, carry = cnt;
  bool pending = true;
  int base = 0;          // index, in the new sequence, of the first element of the current bucket
  int left = n - 1;      // old readings not yet passed
  // a well-formed chain has ceil((n - 1) / 7) buckets; the bound only protects against a corrupted (cyclic)
  // chain restored through rt_env_buffer — a kernel must never spin forever
  const int max_hops = (n + 5) / 7 + 1;
  for (int hop = 0; hop < max_hops; ++hop) {
    uint32_t v[7] = {p_lo.x, p_lo.y, p_lo.z, p_lo.w, p_hi.x, p_hi.y, p_hi.z};
    uint32_t next = p_hi.w & 0xffffu;
    const uint32_t count_word = hop == 0 ? (uint32_t)n << 16 : 0u;
    const int m = left < 7 ? left : 7;  // valid entries of this bucket before the insertion
    left -= m;
    int m_new = m;
    bool dirty = hop == 0;  // the first bucket carries the count
    if (pending && (m < 7 || carry < v[6])) {
      uint32_t x = carry;  // bubble pass: v[0..m-1] keep the m smallest of {v, carry}, x leaves with the largest
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const bool sw = i < m && v[i] > x;
        const uint32_t t = v[i];
        v[i] = sw ? x : t;
        x = sw ? t : x;
      }
      if (m < 7) {  // room left (this is the last bucket): the largest takes the free place
#pragma unroll
        for (int i = 0; i < 7; ++i)
          if (i == m) v[i] = x;
        m_new = m + 1;
        pending = false;
      } else {
        carry = x;  // evicted maximum moves on
      }
      dirty = true;
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      if (i < m_new && base + i == k1) m1 = v[i];
      if (i < m_new && base + i == k2) m2 = v[i];
    }
    base += m_new;
    if (next == 0u && pending) {  // the chain grows by one bucket holding the carried maximum
      uint4* np = pool + (size_t)idx * E * 2;
      np[0] = make_uint4(carry, 0u, 0u, 0u);
      np[1] = make_uint4(0u, 0u, 0u, 0u);
      if (base == k1) m1 = carry;
      if (base == k2) m2 = carry;
      next = (uint32_t)(idx + 1);
      pending = false;
      dirty = true;
      left = -1;  // finished
    }
    if (dirty) {
      bp[0] = make_uint4(v[0], v[1], v[2], v[3]);
      bp[1] = make_uint4(v[4], v[5], v[6], next | count_word);
    }
    if (left < 0 || next == 0u || (!pending && base > k2)) break;
    bp = pool + (size_t)(next - 1) * E * 2;
    p_lo = bp[0];
    p_hi = bp[1];
  }
  // statistics.median: middle element, or the mean of the middle two
  return (n & 1) ? (double)m2 : __ddiv_rn((double)((long long)m1 + (long long)m2), 2.0);
}

// K4a: the env's agents feed ONE Welford standardiser / normaliser, in agent order (one thread per env).
// The per-cell estimator work was already done in parallel by the agents' own threads, except for agents
// whose cell was also entered by an earlier agent of the env in this very step (nread == 0): those are
// replayed here, in order, against the updated bucket chain.
__device__ __forceinline__ void stats_chain(const EnvDev& d, const StepSmem& s, int e, int le, int tick,
                                            WelfordState& w, double emx, double emn, int nv) {
  const int A = d.A;
  const uint32_t etag = (uint32_t)d.episode[e] & 0xffffu;
  uint16_t* crec = d.cellrec + (size_t)e * d.cells * 4;
  uint16_t* visit = d.visit + (size_t)e * d.cells;
  uint2* rec = d.vrec + (size_t)e * d.rec_stride;
  float* inten = d.inten + (size_t)e * d.cells;
  int fl = s.eflags[le];
  for (int aa = 0; aa < A; ++aa) {
    const int i = le * A + aa;
    const int cell = s.cell[i];
    int n = s.nread[i], sl;
    double est;
    if (n != 0) {
      est = s.est[i];
      sl = s.slot[i];
    } else {
      const uint2 cr = *reinterpret_cast<const uint2*>(crec + 4 * cell);
      const bool fresh = (cr.x & 0xffffu) == etag;  // written in THIS episode (an earlier agent of this very step)
      est = bucket_insert_median(d.log + 2 * (size_t)e, (size_t)d.E, crec + 4 * cell, etag, tick * A + aa,
                                 (uint32_t)s.count[i], fresh ? (int)(cr.x >> 16) : 0, n);
      if (d.live_dense) visit[cell] = (uint16_t)n;
      sl = fresh ? (int)(cr.y & 0xffffu) : 0;
    }
    if (sl == 0) {  // first visit of this cell in the episode: open a record
      sl = ++nv;
      crec[4 * cell + 2] = (uint16_t)sl;
    }
    est_maxmin(est, emx, emn);

    // a6: Welford update + z-score; a7: min-max against the running z range
    const double z = welford_update(w, est);
    double val;
    if (!normalize_minmax(z, w.zmax, true, w.zmin, val)) {
      fl |= RT_FLAG_NORMALIZE;
      val = 0.0;
    }
    // one scattered 4-byte store per agent-step is worth ~10 % of this kernel (sensitivity probe):
    // the sparse rasterisers read the record below, so the dense map is only kept current when something reads it
    if (d.live_dense) inten[cell] = (float)val;
    // (n <= cap in any state this library produced; the clamp keeps the rasteriser's table lookup in bounds for a
    // garbage state restored through rt_env_buffer)
    rec[2 + sl] = make_uint2((uint32_t)cell | ((uint32_t)min(n, d.cap) << 16), __float_as_uint((float)val));
    const size_t ia = (size_t)e * A + aa;
    d.last_est[ia] = est;
    d.last_z[ia] = z;
    d.last_value[ia] = val;
  }
  double2* wp = reinterpret_cast<double2*>(d.welford + (size_t)e * 6);
  wp[0] = make_double2(w.mean, w.m2);
  wp[1] = make_double2(w.var, w.sd);
  wp[2] = make_double2(w.zmax, w.zmin);
  d.welford_n[e] = w.count;
  reinterpret_cast<double2*>(d.est_mm)[e] = make_double2(emx, emn);
  d.tick[e] = tick + 1;
  if (fl) atomicOr(&d.flags[e], fl);
  // record header: number of visited cells and where the agents stand now
  unsigned long long p0 = ~0ull, p1 = ~0ull;  // 8 x uint16 agent cells, 0xFFFF = no agent
  for (int aa = 0; aa < A; ++aa) {
    const unsigned long long c = (unsigned long long)(s.cell[le * A + aa] & 0xffff);
    const int sh = 16 * (aa & 3);
    if (aa < 4)
      p0 = (p0 & ~(0xffffull << sh)) | (c << sh);
    else
      p1 = (p1 & ~(0xffffull << sh)) | (c << sh);
  }
  const uint2 o1 = rec[1], o2 = rec[2];  // same 32-byte sector as rec[0], already in flight for nv
  d.prev_cells[e] = make_uint4(o1.x, o1.y, o2.x, o2.y);
  rec[0] = make_uint2((uint32_t)nv, 0u);
  rec[1] = make_uint2((uint32_t)p0, (uint32_t)(p0 >> 32));
  rec[2] = make_uint2((uint32_t)p1, (uint32_t)(p1 >> 32));
}

__device__ __forceinline__ void step_tile(const EnvDev& d, const int32_t* __restrict__ actions,
                                          int32_t* __restrict__ obs_xy, int32_t* __restrict__ reward,
                                          uint8_t* __restrict__ done, unsigned char* smem_raw) {
  const int A = d.A;
  const int epb = blockDim.x / A;
  StepSmem s;
  s.bar = reinterpret_cast<uint64_t*>(smem_raw);
  s.edges = reinterpret_cast<float4*>(smem_raw + 16);
  s.own = reinterpret_cast<uint4*>(smem_raw + 16 + (size_t)epb * kEdgeRowF4 * 16);
  s.est = reinterpret_cast<double*>(s.own + (size_t)epb * kOwnF4);
  s.n_poly = reinterpret_cast<int*>(s.est + epb * A);
  s.skip = s.n_poly + epb;
  s.eflags = s.skip + epb;
  s.cell = s.eflags + epb;
  s.count = s.cell + epb * A;
  s.slot = s.count + epb * A;
  s.nread = s.slot + epb * A;

  const int t = threadIdx.x;
  const int le = t / A;
  const int a = t - le * A;
  const int e0 = blockIdx.x * epb;
  const int e = e0 + le;
  const bool valid = e < d.E;
  const int n_valid = min(epb, d.E - e0);

  if (t == 0) mbar_init(s.bar, (uint32_t)n_valid);

  // Independent global loads first: they overlap the barrier set-up and the geometry copy.
  int act = 1, tick = 0, np = 0;
  int2 pos = make_int2(0, 0), src = make_int2(0, 0);
  int pdist = 0;
  if (valid) {
    act = actions[(size_t)e * A + a];
    tick = d.tick[e];
    pos = reinterpret_cast<const int2*>(d.agent_xy)[(size_t)e * A + a];
    pdist = d.prev_dist[(size_t)e * A + a];
    src = reinterpret_cast<const int2*>(d.src_xy)[e];
    if (a == 0) {
      np = d.n_poly[e];
      s.n_poly[le] = np;
      s.skip[le] = tick >= d.T ? RT_FLAG_LOG_FULL : 0;
      s.eflags[le] = 0;
    }
  }
  __syncthreads();  // barrier initialised, per-env slots initialised

  if (valid) {
    if (a == 0) {
      if (np > 0) {
        const uint32_t bytes = (uint32_t)np * RT_EDGES_PER_POLY * 16u;
        mbar_arrive_expect_tx(s.bar, bytes);
        bulk_g2s(s.edges + (size_t)le * kEdgeRowF4, d.edges + (size_t)e * RT_MAX_EDGES * 4, bytes, s.bar);
      } else {
        mbar_arrive(s.bar);
      }
    }
    if (act < 1 || act > 4) atomicOr(&s.skip[le], RT_FLAG_BAD_ACTION);  // simple_gridworld.py:138
  }

  // Per-env statistics are owned by the FIRST n_valid threads of the CTA (thread i <-> env e0 + i): the
  // sequential chain at the end is pure arithmetic + stores (the bucket walks happen in the agents' own
  // threads), so it runs best in fully populated warps.  Start the owners' loads now.
  double src_int = 0.0, bkg = 0.0;
  int episode = 0;
  if (valid) {
    src_int = d.src_int[e];
    bkg = d.bkg[e];
    episode = d.episode[e];
  }
  const bool owner = t < n_valid;
  const int e3 = e0 + t;
  if (owner) {
    unsigned char* row = reinterpret_cast<unsigned char*>(s.own + (size_t)t * kOwnF4);
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(d.welford + (size_t)e3 * 6);
    cp_async16(row, wsrc);
    cp_async16(row + 16, wsrc + 16);
    cp_async16(row + 32, wsrc + 32);
    cp_async16(row + 48, d.est_mm + (size_t)e3 * 2);
    cp_async8(row + 64, d.vrec + (size_t)e3 * d.rec_stride);
    cp_async4(row + 72, d.welford_n + e3);
    cp_async4(row + 76, d.tick + e3);
    cp_async_commit();
  }

  mbar_wait(s.bar, 0);  // geometry of every env of this CTA has landed in shared memory
  __syncthreads();      // skip flags complete

  // ---- K1: motion, collision, reward (one thread per agent) ------------------------------------
  const int skip = valid ? s.skip[le] : 0;
  const bool live = valid && !skip;
  const size_t ia = (size_t)e * A + a;
  const float4* my_edges = s.edges + (size_t)le * kEdgeRowF4;
  int nx = pos.x, ny = pos.y, cell = 0;
  unsigned ph = 0u, psl = 0u;
  if (valid) {
    if (skip) {  // reference raises before touching state: report the unchanged position
      reinterpret_cast<int2*>(obs_xy)[ia] = pos;
      reward[ia] = 0;
      done[ia] = 0;
    } else {
      np = s.n_poly[le];
      // a4: agent = clip(agent + direction, 0, size-1)   (simple_gridworld.py:26-31,141)
      const int dx = (act == 2) - (act == 3);
      const int dy = (act == 1) - (act == 4);
      nx = min(max(pos.x + dx, 0), d.G - 1);
      ny = min(max(pos.y + dy, 0), d.G - 1);
      // N4: refuse a move whose centre-to-centre path crosses an obstruction edge
      if ((nx != pos.x || ny != pos.y) && np > 0 &&
          polys_crossed(my_edges, np, (float)pos.x + 0.5f, (float)pos.y + 0.5f, (float)nx + 0.5f,
                        (float)ny + 0.5f) > 0) {
        nx = pos.x;
        ny = pos.y;
      }
      // a4: reward / done on the L1 distance to the source cell (:144-159)
      const int dist = abs(nx - src.x) + abs(ny - src.y);
      const bool dn = dist <= 1;
      const int rw = dn ? 1 : (dist < pdist ? 0 : -1);
      reinterpret_cast<int2*>(d.agent_xy)[ia] = make_int2(nx, ny);
      d.prev_dist[ia] = dist;
      reinterpret_cast<int2*>(obs_xy)[ia] = make_int2(nx, ny);
      reward[ia] = rw;
      done[ia] = dn ? 1 : 0;
      d.ep_return[ia] += rw;
      if (dn) d.ep_done[ia] = 1;
    }
    cell = ny * d.G + nx;
    s.cell[t] = cell;
    if (live) {  // cell state for the estimator: the loads fly while line of sight and sampler compute
      const size_t mcell = (size_t)e * d.cells + cell;
      const uint2 cr = reinterpret_cast<const uint2*>(d.cellrec)[mcell];  // episode tag | first bucket << 16, slot
      if ((cr.x & 0xffffu) == ((uint32_t)episode & 0xffffu)) {  // else: a record of an older episode = empty
        ph = cr.x >> 16;
        psl = cr.y & 0xffffu;
      }
      // The owner will write this cell's 8-byte visited-cell record at the end.  A partial write into a sector
      // that is NOT in the L2 is dear for whoever reads it next (the rasteriser): fetch the sector now, the write
      // then lands in a complete line (-1.6 % on the step, profiles/r03m).
      if (psl != 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.vrec + (size_t)e * d.rec_stride + 2 + psl));
    }
  }
  __syncthreads();  // every agent's new cell is known to its env-mates

  // ---- K2: measurement, then the agent's own estimator update ------------------------------------
  if (live) {
    bool clash = false;  // an earlier agent of this env entered the same cell in this very step
    for (int bb = 0; bb < a; ++bb) clash |= s.cell[le * A + bb] == cell;
    // N2 + N1: obstructions on the line of sight, expected count
    const int nb = np > 0 ? polys_crossed(my_edges, np, (float)nx + 0.5f, (float)ny + 0.5f,
                                          (float)src.x + 0.5f, (float)src.y + 0.5f)
                          : 0;
    const int ddx = nx - src.x, ddy = ny - src.y;
    double d2 = __dmul_rn((double)(ddx * ddx + ddy * ddy), d.pitch2);
    if (d2 < d.dmin2) d2 = d.dmin2;
    double strength = src_int;
    for (int i = 0; i < nb; ++i) strength = __dmul_rn(strength, d.transmission);
    const double lam = __dadd_rn(__ddiv_rn(strength, d2), bkg);

    // pull the cell's first bucket towards the L2 while the sampler computes.  (Loading it into registers here
    // costs spills across the sampler: 0.0507 vs 0.0487 ms, profiles/r03d.)
    if (!clash && ph != 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.log + ((size_t)(ph - 1) * d.E + e) * 2));

    // N3: Poisson count
    Draws dr;
    dr.init(d.seed, d.gid_base + e, STREAM_MEASURE, a, tick, episode);
    if (d.inj != nullptr) {
      dr.inj = d.inj + ((size_t)e * A + a) * (size_t)d.inj_k;
      dr.inj_k = d.inj_k;
    }
    int fl = 0;
    const long long cnt = poisson_draw(lam, dr, fl);
    if (fl) atomicOr(&s.eflags[le], fl);

    // a9 in parallel across the env's agents (different cells => independent lists)
    int n = 0;
    double est = 0.0;
    if (!clash) {
      uint16_t* cr16 = d.cellrec + ((size_t)e * d.cells + cell) * 4;
      est = bucket_insert_median(d.log + 2 * (size_t)e, (size_t)d.E, cr16, (uint32_t)episode & 0xffffu, tick * A + a,
                                 (uint32_t)cnt, (int)ph, n);
      if (d.live_dense) d.visit[(size_t)e * d.cells + cell] = (uint16_t)n;
    }
    s.count[t] = (int)cnt;
    s.slot[t] = (int)psl;
    s.nread[t] = n;
    s.est[t] = est;
    d.last_count[ia] = (int)cnt;
    if (d.rollout != nullptr) d.rollout[((size_t)tick * d.E + e) * A + a] = (float)cnt;
    d.last_los[ia] = (uint8_t)nb;
    d.last_lambda[ia] = lam;
  }
  __syncthreads();  // readings and estimates of all agents are in shared memory

  // ---- K4a: the env's sequential statistics -------------------------------------------------------
  if (owner) {
    const int skip3 = s.skip[t];
    if (skip3) {
      atomicOr(&d.flags[e3], skip3);
      atomicOr(d.step_flags + (d.host_seq & 1ull), skip3);
    } else {
      cp_async_wait_all();  // this thread's own copies: its row of the staged state is complete
      const unsigned char* row = reinterpret_cast<const unsigned char*>(s.own + (size_t)t * kOwnF4);
      const double2 w0 = *reinterpret_cast<const double2*>(row), w1 = *reinterpret_cast<const double2*>(row + 16),
                    w2 = *reinterpret_cast<const double2*>(row + 32), mm = *reinterpret_cast<const double2*>(row + 48);
      WelfordState w;
      w.mean = w0.x; w.m2 = w0.y; w.var = w1.x; w.sd = w1.y; w.zmax = w2.x; w.zmin = w2.y;
      w.count = *reinterpret_cast<const int*>(row + 72);
      stats_chain(d, s, e3, t, *reinterpret_cast<const int*>(row + 76), w, mm.x, mm.y,
                  (int)*reinterpret_cast<const uint32_t*>(row + 64));
    }
  }
}

#ifndef RT_STEP_MIN_CTAS
#define RT_STEP_MIN_CTAS 4  // resident CTAs per SM ptxas must allow for (register cap = 65536 / (192 * this))
#endif
__global__ void __launch_bounds__(192, RT_STEP_MIN_CTAS)
k_step(const EnvDev d, const int32_t* __restrict__ actions, int32_t* __restrict__ obs_xy,
       int32_t* __restrict__ reward, uint8_t* __restrict__ done) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // Wait FIRST, then let the dependent kernel (this step's rasteriser) start: its CTAs may then assume that everything
  // in front of k_step in the stream — the previous step's rasteriser, the consumer that read the observation tensor —
  // has completed, which is what allows k_raster_small to write its zeros before its own wait.  (Triggering before the
  // wait would let rasteriser t+1 start while rasteriser t is still patching the same tensor.)
  pdl_wait();  // the rasteriser of the previous step still reads the records this kernel updates
  pdl_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // host path bookkeeping (see EnvDev::step_flags)
    d.step_flags[(d.host_seq + 1ull) & 1ull] = 0;
    *reinterpret_cast<unsigned long long*>(d.step_flags + 2) = d.host_seq;
  }
  step_tile(d, actions, obs_xy, reward, done, smem_raw);
  if (d.host_seq_out != nullptr) {  // uniform over the grid
    __threadfence_system();         // this thread's result stores (host memory) are visible system-wide ...
    __syncthreads();                // ... for every thread of the CTA, before the CTA checks in
    if (threadIdx.x == 0 && atomicAdd(d.cta_counter, 1u) == gridDim.x - 1) {  // the last CTA: everything has landed
      *d.cta_counter = 0u;
      __threadfence();
      d.host_flag_out[d.host_seq & 1ull] = *reinterpret_cast<volatile int32_t*>(d.step_flags + (d.host_seq & 1ull));
      __threadfence_system();
      *d.host_seq_out = d.host_seq;
    }
  }
}

// --------------------------------------------------------------------------------------------
// K5: masked reset.  One thread per env regenerates the scenario (stream 1) and clears the
// scalar state; k_zero_maps clears the three persistent maps with 16-byte stores.
__device__ __forceinline__ int draw_int(Draws& dr, int n) {
  int v = (int)floor(__dmul_rn(dr.next(), (double)n));
  return v >= n ? n - 1 : v;
}

__global__ void __launch_bounds__(128)
k_reset_state(const EnvDev d, const uint8_t* __restrict__ mask, int32_t* __restrict__ obs_xy) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.E) return;
  const int A = d.A, G = d.G;
  const bool doit = mask == nullptr || mask[e] != 0;
  int2* agent = reinterpret_cast<int2*>(d.agent_xy) + (size_t)e * A;
  if (!doit) {
    if (obs_xy)
      for (int a = 0; a < A; ++a) reinterpret_cast<int2*>(obs_xy)[(size_t)e * A + a] = agent[a];
    return;
  }
  const int episode = d.episode[e] + 1;
  d.episode[e] = episode;
  int2* start = reinterpret_cast<int2*>(d.start_xy) + (size_t)e * A;
  int2 src;
  if (!d.fixed) {
    Draws dr;
    dr.init(d.seed, d.gid_base + e, STREAM_SCENARIO, 0, 0, episode);
    src.x = draw_int(dr, G);
    src.y = draw_int(dr, G);
    d.src_int[e] = __dadd_rn(d.src_lo, __dmul_rn(dr.next(), d.src_span));
    d.bkg[e] = __dadd_rn(d.bkg_lo, __dmul_rn(dr.next(), d.bkg_span));
    reinterpret_cast<int2*>(d.src_xy)[e] = src;
    int2 st[RT_MAX_AGENTS];
    for (int a = 0; a < A; ++a) {
      bool ok = false;
      int2 p = make_int2(0, 0);
      for (int t = 0; t < 8 && !ok; ++t) {
        p.x = draw_int(dr, G);
        p.y = draw_int(dr, G);
        ok = abs(p.x - src.x) + abs(p.y - src.y) > 1;
      }
      if (!ok) p = make_int2((src.x + G / 2) % G, (src.y + G / 2) % G);
      st[a] = p;
      start[a] = p;
    }
    float4* edges = reinterpret_cast<float4*>(d.edges) + (size_t)e * RT_MAX_EDGES;
    for (int j = 0; j < RT_MAX_EDGES; ++j) edges[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    int np = d.poly_min;
    if (d.poly_max > d.poly_min) np += draw_int(dr, d.poly_max - d.poly_min + 1);
    d.n_poly[e] = np;
    for (int p = 0; p < np; ++p) {
      for (int t = 0; t < 8; ++t) {
        int w = d.rect_min + draw_int(dr, d.rect_max - d.rect_min + 1);
        int h = d.rect_min + draw_int(dr, d.rect_max - d.rect_min + 1);
        w = min(w, G);
        h = min(h, G);
        const int x0 = draw_int(dr, G - w + 1);
        const int y0 = draw_int(dr, G - h + 1);
        bool bad = src.x >= x0 && src.x < x0 + w && src.y >= y0 && src.y < y0 + h;
        for (int a = 0; a < A; ++a)
          bad |= st[a].x >= x0 && st[a].x < x0 + w && st[a].y >= y0 && st[a].y < y0 + h;
        if (bad) continue;
        const float X0 = (float)x0, Y0 = (float)y0, X1 = (float)(x0 + w), Y1 = (float)(y0 + h);
        edges[4 * p + 0] = make_float4(X0, Y0, X1, Y0);
        edges[4 * p + 1] = make_float4(X1, Y0, X1, Y1);
        edges[4 * p + 2] = make_float4(X1, Y1, X0, Y1);
        edges[4 * p + 3] = make_float4(X0, Y1, X0, Y0);
        break;
      }
    }
  } else {
    src = reinterpret_cast<const int2*>(d.src_xy)[e];
  }
  for (int a = 0; a < A; ++a) {
    const size_t ia = (size_t)e * A + a;
    const int2 p = start[a];
    agent[a] = p;  // simple_gridworld.py:120
    d.prev_dist[ia] = abs(p.x - src.x) + abs(p.y - src.y);  // :121
    if (obs_xy) reinterpret_cast<int2*>(obs_xy)[ia] = p;
    d.last_count[ia] = 0;
    d.ep_return[ia] = 0;
    d.ep_done[ia] = 0;
    d.last_los[ia] = 0;
    d.last_lambda[ia] = 0.0;
    d.last_est[ia] = 0.0;
    d.last_z[ia] = 0.0;
    d.last_value[ia] = 0.0;
  }
  d.tick[e] = 0;
  {
    unsigned long long p0 = ~0ull, p1 = ~0ull;
    for (int a = 0; a < A; ++a) {
      const unsigned long long c = (unsigned long long)((start[a].y * G + start[a].x) & 0xffff);
      const int sh = 16 * (a & 3);
      if (a < 4)
        p0 = (p0 & ~(0xffffull << sh)) | (c << sh);
      else
        p1 = (p1 & ~(0xffffull << sh)) | (c << sh);
    }
    uint2* rec = d.vrec + (size_t)e * d.rec_stride;
    rec[0] = make_uint2(0u, 0u);
    rec[1] = make_uint2((uint32_t)p0, (uint32_t)(p0 >> 32));
    rec[2] = make_uint2((uint32_t)p1, (uint32_t)(p1 >> 32));
    d.prev_cells[e] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  }
  double2* wp = reinterpret_cast<double2*>(d.welford + (size_t)e * 6);
  wp[0] = make_double2(0.0, 0.0);  // standardize.py:36-44
  wp[1] = make_double2(0.0, 1.0);
  wp[2] = make_double2(0.0, 0.0);
  d.welford_n[e] = 0;
  reinterpret_cast<double2*>(d.est_mm)[e] = make_double2(0.0, 0.0);  // estimator.py:25-29
}

// dense visit (2 B) + dense intensity (4 B) per cell when a dense rasteriser keeps them, and the cell records (8 B per
// cell) only when their episode tag wraps; 16-byte stores, one CTA per env (which exits at once when it has nothing to do).
__global__ void __launch_bounds__(256)
k_zero_maps(const EnvDev d, const uint8_t* __restrict__ mask) {
  // A grid of a few CTAs per SM strides over the envs: in the default (sparse) mode almost every env has nothing to
  // clear, and one CTA per env — 65,536 CTAs that exit at once — cost 42 us of pure CTA scheduling (ncu, r2v).
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const size_t b16 = (size_t)d.cells * 2, b32 = (size_t)d.cells * 4, b64 = (size_t)d.cells * 8;
  for (int e = blockIdx.x; e < d.E; e += gridDim.x) {
    if (mask != nullptr && mask[e] == 0) continue;
    // cell records carry an episode tag: a reset leaves them alone, except when the 16-bit tag wraps (the episode
    // number k_reset_state has just written is a multiple of 65,536 — also the very first episode, over a zeroed slab)
    const bool wrap = (d.episode[e] & 0xffff) == 0;
    if (!d.live_dense && !wrap) continue;
    unsigned char* v = reinterpret_cast<unsigned char*>(d.visit) + (size_t)e * b16;
    unsigned char* c = reinterpret_cast<unsigned char*>(d.cellrec) + (size_t)e * b64;
    unsigned char* f = reinterpret_cast<unsigned char*>(d.inten) + (size_t)e * b32;
    if (wrap) {
      if ((d.cells & 1) == 0) {  // an env's cell records start 16-byte aligned only when cells is even
        for (size_t i = threadIdx.x * 16; i < b64; i += blockDim.x * 16) *reinterpret_cast<uint4*>(c + i) = z;
      } else {
        for (size_t i = threadIdx.x * 8; i < b64; i += blockDim.x * 8)
          *reinterpret_cast<uint2*>(c + i) = make_uint2(0u, 0u);
      }
    }
    // the dense visit / intensity maps are only kept when a dense rasteriser reads them; otherwise they are
    // materialised on request (rt_env_buffer clears them first) and a reset clears nothing here
    if (!d.live_dense) continue;
    if ((d.cells & 7) == 0) {
      for (size_t i = threadIdx.x * 16; i < b16; i += blockDim.x * 16) *reinterpret_cast<uint4*>(v + i) = z;
      for (size_t i = threadIdx.x * 16; i < b32; i += blockDim.x * 16) *reinterpret_cast<uint4*>(f + i) = z;
    } else {
      for (int i = threadIdx.x; i < d.cells; i += blockDim.x) {
        d.visit[(size_t)e * d.cells + i] = 0;
        d.inten[(size_t)e * d.cells + i] = 0.f;
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// K3: rasterise.  Vector path (cells % 8 == 0): a thread owns 8 consecutive cells of one env —
// one 16 B load of visit counts, two 16 B loads of intensities, six 16 B streaming stores
// (3 channels x 32 B).  A warp therefore writes 3 x 1 KB fully coalesced per pass.  Almost every
// cell of a map is unvisited, so the visit-table lookups are skipped when the 8 counts are all 0.
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {  // read-only data of an earlier kernel
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_cg_u4(const void* p) {  // data written earlier in THIS kernel: L2 only
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_plain_f4(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// one 8-cell chunk; bit j of `occ` = some agent stands on cell j of the chunk
__device__ __forceinline__ float bit_f(unsigned m, int j) { return (m >> j) & 1u ? 1.0f : 0.0f; }
__device__ __forceinline__ void raster_chunk(float* __restrict__ o, int cells, const uint4 vv, const uint4 i0,
                                             const uint4 i1, unsigned occ, const float* lut) {
  st_stream_f4(o, make_float4(bit_f(occ, 0), bit_f(occ, 1), bit_f(occ, 2), bit_f(occ, 3)));
  st_stream_f4(o + 4, make_float4(bit_f(occ, 4), bit_f(occ, 5), bit_f(occ, 6), bit_f(occ, 7)));
  o += cells;
  st_stream_f4(o, make_float4(__uint_as_float(i0.x), __uint_as_float(i0.y), __uint_as_float(i0.z),
                              __uint_as_float(i0.w)));
  st_stream_f4(o + 4, make_float4(__uint_as_float(i1.x), __uint_as_float(i1.y), __uint_as_float(i1.z),
                                  __uint_as_float(i1.w)));
  o += cells;
  if ((vv.x | vv.y | vv.z | vv.w) == 0u) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    st_stream_f4(o, z);
    st_stream_f4(o + 4, z);
  } else {
    st_stream_f4(o, make_float4(lut[vv.x & 0xffffu], lut[vv.x >> 16], lut[vv.y & 0xffffu], lut[vv.y >> 16]));
    st_stream_f4(o + 4, make_float4(lut[vv.z & 0xffffu], lut[vv.z >> 16], lut[vv.w & 0xffffu], lut[vv.w >> 16]));
  }
}

// kUnroll chunks per thread are loaded before any is stored (bytes in flight).
template <bool kLutInSmem, int kUnroll>
__global__ void __launch_bounds__(256)
k_raster_vec(const EnvDev d, float* __restrict__ out, long long n_chunks) {
  extern __shared__ float s_lut[];
  if (kLutInSmem) {
    for (int i = threadIdx.x; i <= d.cap; i += blockDim.x) s_lut[i] = d.lut[i];
    __syncthreads();
  }
  const float* lut = kLutInSmem ? s_lut : d.lut;
  const int cpe = d.cells >> 3;  // chunks per env
  const int A = d.A, G = d.G;
  // (env, chunk-in-env) advance incrementally: no integer division in the streaming loop
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long c0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int e = (int)(c0 / cpe);
  int q = (int)(c0 - (long long)e * cpe);
  const int de = (int)(stride / cpe);
  const int dq = (int)(stride - (long long)de * cpe);
  for (long long c = c0; c < n_chunks; c += kUnroll * stride) {
    uint4 vv[kUnroll], i0[kUnroll], i1[kUnroll];
    int ee[kUnroll], qq[kUnroll];
    bool ok[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      ee[u] = e;
      qq[u] = q;
      ok[u] = c + u * stride < n_chunks;
      if (ok[u]) {
        const size_t cell0 = (size_t)e * d.cells + (size_t)q * 8;
        vv[u] = ld_stream_u4(d.visit + cell0);
        i0[u] = ld_stream_u4(d.inten + cell0);
        i1[u] = ld_stream_u4(d.inten + cell0 + 4);
      }
      e += de;
      q += dq;
      if (q >= cpe) {
        q -= cpe;
        e += 1;
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (!ok[u]) continue;
      unsigned occ = 0u;
      const int2* agent = reinterpret_cast<const int2*>(d.agent_xy) + (size_t)ee[u] * A;
      for (int a = 0; a < A; ++a) {
        const int2 p = agent[a];
        const unsigned r = (unsigned)(p.y * G + p.x - qq[u] * 8);
        if (r < 8u) occ |= 1u << r;
      }
      raster_chunk(out + (size_t)ee[u] * 3 * d.cells + (size_t)qq[u] * 8, d.cells, vv[u], i0[u], i1[u], occ, lut);
    }
  }
}

// --------------------------------------------------------------------------------------------
// K3, sparse form (default): a map of a 120-step episode has at most a few dozen non-zero cells, so the
// rasteriser does not re-read the dense maps (6 KB/env).  It reads the env's compact record list kept
// by k_step — header, agent cells and the first records in ONE bulk async copy (TMA engine, UBLKCP) —
// composes the env's whole [3][G][G] tile in shared memory (zero, scatter records, mark agents) and
// ships it with ONE 12 KB bulk store.  HBM traffic per env-step: 12,288 B written + rec_prefix*8 B read.
// One CTA of 128 threads walks envs blockIdx.x, + gridDim.x, ... through a ring of kSpStages
// (tile, record buffer) stages; thread 0 issues every async operation.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
// L2 cache policies: the observation tensor is written once and not re-read by us (evict first), the
// record lists are re-read every step by k_step and the rasteriser (evict last).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kSpWarps = 8;      // warps per CTA, each an independent pipeline
constexpr int kSpThreads = kSpWarps * 32;
constexpr int kSpStages = 2;     // (tile, record buffer) stages per warp
__host__ __device__ inline size_t raster_sparse_smem_bytes(int cells, int prefix, int lut_entries) {
  // [mbarriers, 128 B] [kSpWarps*kSpStages tiles of 12*cells B] [as many record buffers of 8*prefix B] [table]
  return 128 + (size_t)kSpWarps * kSpStages * ((size_t)cells * 12 + (size_t)prefix * 8) + (size_t)lut_entries * 4;
}

// Every WARP is an autonomous pipeline over its own envs (global warp id, + total warps, ...): no CTA
// barrier in the loop.  Lane 0 issues the bulk copies and owns their completion (bulk groups and
// mbarrier arrivals are per thread); the lanes zero the tile, scatter the records and mark the agents.
// While the bulk store of one tile drains, the warp composes the other.
template <bool kLutInSmem>
__global__ void __launch_bounds__(kSpThreads, 1)
k_raster_sparse(const EnvDev d, float* __restrict__ out, int hints) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int cells = d.cells, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t tile_bytes = (uint32_t)cells * 12u, pre_bytes = (uint32_t)d.rec_prefix * 8u;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm) + w * kSpStages;
  unsigned char* tiles = sm + 128 + (size_t)w * kSpStages * tile_bytes;
  unsigned char* recs = sm + 128 + (size_t)kSpWarps * kSpStages * tile_bytes + (size_t)w * kSpStages * pre_bytes;
  float* s_lut = reinterpret_cast<float*>(sm + 128 + (size_t)kSpWarps * kSpStages * (tile_bytes + pre_bytes));

  if (kLutInSmem)
    for (int i = tid; i <= d.cap; i += kSpThreads) s_lut[i] = d.lut[i];
  if (lane == 0)
    for (int s = 0; s < kSpStages; ++s) mbar_init(full + s, 1u);
  __syncthreads();  // the only CTA-wide barrier: table staged, barriers initialised
  const float* lut = kLutInSmem ? s_lut : d.lut;

  const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
  const long long n_warps = (long long)gridDim.x * kSpWarps;
  const long long gw = (long long)blockIdx.x * kSpWarps + w;
  const int n_mine = gw < d.E ? (int)((d.E - gw + n_warps - 1) / n_warps) : 0;  // envs of this warp
  auto issue_load = [&](int k) {  // lane 0: record prefix of this warp's env k into stage k % kSpStages
    const int s = k % kSpStages;
    const size_t e = (size_t)(gw + (long long)k * n_warps);
    mbar_arrive_expect_tx(full + s, pre_bytes);
    if (hints & 2)
      bulk_g2s_hint(recs + (size_t)s * pre_bytes, d.vrec + e * d.rec_stride, pre_bytes, full + s, pol_last);
    else
      bulk_g2s(recs + (size_t)s * pre_bytes, d.vrec + e * d.rec_stride, pre_bytes, full + s);
  };
  if (lane == 0)
    for (int k = 0; k < kSpStages && k < n_mine; ++k) issue_load(k);

  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < n_mine; ++k) {
    const int s = k % kSpStages;
    const size_t e = (size_t)(gw + (long long)k * n_warps);
    float* tile = reinterpret_cast<float*>(tiles + (size_t)s * tile_bytes);
    const uint2* rec = reinterpret_cast<const uint2*>(recs + (size_t)s * pre_bytes);
    // the bulk store that last read this tile (kSpStages envs ago) must have finished reading it
    if (lane == 0 && k >= kSpStages) bulk_wait_read<kSpStages - 1>();
    __syncwarp();
    for (int i = lane; i < (cells * 3) >> 2; i += 32) reinterpret_cast<float4*>(tile)[i] = z4;
    mbar_wait(full + s, (uint32_t)((k / kSpStages) & 1));  // record prefix has landed
    __syncwarp();  // zeroing by all lanes is ordered before any lane's scatter
    const int nv = (int)rec[0].x;
    for (int i = lane; i < nv; i += 32) {
      const uint2 r = (3 + i < d.rec_prefix) ? rec[3 + i] : d.vrec[e * d.rec_stride + 3 + i];
      const int cell = (int)(r.x & 0xffffu);
      tile[cells + cell] = __uint_as_float(r.y);
      tile[2 * cells + cell] = lut[r.x >> 16];
    }
    if (lane < RT_MAX_AGENTS) {
      const uint32_t c = reinterpret_cast<const uint16_t*>(rec + 1)[lane];
      if (c != 0xffffu) tile[c] = 1.0f;
    }
    fence_proxy_async();  // generic-proxy writes of the tile -> visible to the bulk store
    __syncwarp();         // also: every lane is done reading this stage's record buffer
    if (lane == 0) {
      if (hints & 1)
        bulk_s2g_hint(out + e * 3 * (size_t)cells, tile, tile_bytes, pol_first);
      else
        bulk_s2g(out + e * 3 * (size_t)cells, tile, tile_bytes);
      bulk_commit();
      if (k + kSpStages < n_mine) issue_load(k + kSpStages);
    }
  }
  if (lane == 0) bulk_wait_all();
}

// K3, sparse form with LSU stores.  Measured on B200 (profiles/r01p): a pure stream of 12 KB bulk (TMA)
// stores tops out at 6.2 TB/s and interleaved record reads cost it disproportionately, while plain
// coalesced stores reach the 7.4 TB/s of a fill.  So here every warp owns ONE tile in shared memory
// (zeroed once), and per env: the first 128 record slots arrive in registers (4 x 8 B per lane, loaded
// one env ahead), the lanes scatter them into the tile, the warp streams the tile out with 24 coalesced
// 16-byte stores per lane, and the scattered cells are cleared again.  No barriers beyond __syncwarp.
constexpr int kSlWarps = 8;
constexpr int kSlThreads = kSlWarps * 32;
__host__ __device__ inline size_t raster_sl_smem_bytes(int cells, int lut_entries) {
  return (size_t)kSlWarps * cells * 12 + (size_t)lut_entries * 4;
}

template <bool kLutInSmem>
__global__ void __launch_bounds__(kSlThreads)
k_raster_sparse_lsu(const EnvDev d, float* __restrict__ out, int plain_stores) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int cells = d.cells, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float* tile = reinterpret_cast<float*>(sm + (size_t)w * cells * 12);
  float* s_lut = reinterpret_cast<float*>(sm + (size_t)kSlWarps * cells * 12);
  if (kLutInSmem)
    for (int i = tid; i <= d.cap; i += kSlThreads) s_lut[i] = d.lut[i];
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n4 = (cells * 3) >> 2;  // float4 per tile
  pdl_launch_dependents();
  for (int i = lane; i < n4; i += 32) reinterpret_cast<float4*>(tile)[i] = z4;
  __syncthreads();  // table staged (the only CTA-wide barrier)
  const float* lut = kLutInSmem ? s_lut : d.lut;
  pdl_wait();  // tiles are zeroed while the step kernel drains; its records are read from here on

  // Static strided assignment: warp gw takes envs gw, gw + n_warps, ...  The grid is deliberately MANY
  // short-lived CTAs (24 per SM, 2 resident): measured best — persistent CTAs leave a tail of stragglers,
  // and handing envs out through one atomic counter serialises on that address (profiles/r01u, r01v).
  const long long n_warps = (long long)gridDim.x * kSlWarps;
  const long long gw = (long long)blockIdx.x * kSlWarps + w;
  auto load4 = [&](long long e, uint2 (&r)[4]) {  // slots lane, lane+32, lane+64, lane+96 of env e
    const uint2* p = d.vrec + (size_t)e * d.rec_stride;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = lane + 32 * m;
      r[m] = j < d.rec_stride ? p[j] : make_uint2(0u, 0u);
    }
  };
  // records are prefetched two envs ahead (one env's compose + stream-out is much shorter than a DRAM
  // round trip)
  uint2 nx1[4] = {make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u)};
  uint2 nx2[4] = {make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u)};
  if (gw < d.E) load4(gw, nx1);
  if (gw + n_warps < d.E) load4(gw + n_warps, nx2);
  for (long long e = gw; e < d.E; e += n_warps) {
    uint2 cur[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      cur[m] = nx1[m];
      nx1[m] = nx2[m];
    }
    if (e + 2 * n_warps < d.E) load4(e + 2 * n_warps, nx2);
    const int nv = (int)__shfl_sync(0xffffffffu, cur[0].x, 0);
    // agent cells: slots 1 and 2 (lanes 1, 2) hold 8 x uint16
    const uint32_t w1x = __shfl_sync(0xffffffffu, cur[0].x, 1), w1y = __shfl_sync(0xffffffffu, cur[0].y, 1);
    const uint32_t w2x = __shfl_sync(0xffffffffu, cur[0].x, 2), w2y = __shfl_sync(0xffffffffu, cur[0].y, 2);
    int acell = 0xffff;
    if (lane < RT_MAX_AGENTS) {
      const uint32_t word = lane < 2 ? w1x : (lane < 4 ? w1y : (lane < 6 ? w2x : w2y));
      acell = (int)((lane & 1) ? (word >> 16) : (word & 0xffffu));
      if (acell != 0xffff) tile[acell] = 1.0f;
    }
    // records: slot j = lane + 32 m holds record j - 3 (valid when 3 <= j < 3 + nv)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = lane + 32 * m;
      if (j >= 3 && j - 3 < nv) {
        const int cell = (int)(cur[m].x & 0xffffu);
        tile[cells + cell] = __uint_as_float(cur[m].y);
        tile[2 * cells + cell] = lut[cur[m].x >> 16];
      }
    }
    for (int j = 128 + lane; j - 3 < nv; j += 32) {  // rare: more than 125 visited cells
      const uint2 r = d.vrec[(size_t)e * d.rec_stride + j];
      const int cell = (int)(r.x & 0xffffu);
      tile[cells + cell] = __uint_as_float(r.y);
      tile[2 * cells + cell] = lut[r.x >> 16];
    }
    __syncwarp();
    float* o = out + (size_t)e * 3 * cells;
    if (plain_stores)
      for (int i = lane; i < n4; i += 32) st_plain_f4(o + 4 * i, reinterpret_cast<const float4*>(tile)[i]);
    else
      for (int i = lane; i < n4; i += 32) st_stream_f4(o + 4 * i, reinterpret_cast<const float4*>(tile)[i]);
    __syncwarp();
    // clear exactly what was set
    if (acell != 0xffff) tile[acell] = 0.0f;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = lane + 32 * m;
      if (j >= 3 && j - 3 < nv) {
        const int cell = (int)(cur[m].x & 0xffffu);
        tile[cells + cell] = 0.0f;
        tile[2 * cells + cell] = 0.0f;
      }
    }
    for (int j = 128 + lane; j - 3 < nv; j += 32) {
      const int cell = (int)(d.vrec[(size_t)e * d.rec_stride + j].x & 0xffffu);
      tile[cells + cell] = 0.0f;
      tile[2 * cells + cell] = 0.0f;
    }
    __syncwarp();
  }
}

// K3 for SMALL batches (observation tensor L2-resident, e.g. BASELINE configs[1]: 4,096 envs = 50 MB).  The tile
// rasterisers above need 12 KB of shared memory per warp (2 CTAs per SM): 4,096 envs are 1.7 waves of a kernel whose
// every wave costs a full record-load latency, behind a k_step that is itself one latency chain.  Here every env
// gets its own warp and NO shared memory, so all warps are resident at once, and the work is split around the
// programmatic-dependent-launch wait:
//   before the wait (while k_step still runs — nothing here depends on it): the warp streams the env's 12 KB of
//     zeros with coalesced 16-byte stores.  The tensor's previous readers were launched before k_step and have exited;
//   after the wait: it loads the env's record list and patches the few non-zero cells (agents, visited cells) with
//     4-byte stores into lines it has just written — they are still in the L2.  __syncwarp orders the two phases.
__global__ void __launch_bounds__(256)
k_raster_small(const EnvDev d, float* __restrict__ out) {
  const int cells = d.cells, lane = threadIdx.x & 31;
  const long long e = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_launch_dependents();
  if (e >= d.E) {
    pdl_wait();
    return;
  }
  float* o = out + (size_t)e * 3 * cells;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n4 = (cells * 3) >> 2;
  for (int i = lane; i < n4; i += 32) st_plain_f4(o + 4 * i, z4);
  pdl_wait();  // k_step's records are complete and visible from here on
  const uint2* rec = d.vrec + (size_t)e * d.rec_stride;
  const uint2 h0 = rec[0];
  const int nv = (int)h0.x;
  uint2 r[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {  // slots lane, lane + 32, ...: header, agent cells, the first 125 records
    const int j = lane + 32 * m;
    r[m] = (j < 3 + nv && j < d.rec_stride) ? rec[j] : make_uint2(0u, 0u);
  }
  __syncwarp();  // the zeros of every lane are ordered before any lane's patch
  const uint32_t w1x = __shfl_sync(0xffffffffu, r[0].x, 1), w1y = __shfl_sync(0xffffffffu, r[0].y, 1);
  const uint32_t w2x = __shfl_sync(0xffffffffu, r[0].x, 2), w2y = __shfl_sync(0xffffffffu, r[0].y, 2);
  if (lane < RT_MAX_AGENTS) {
    const uint32_t word = lane < 2 ? w1x : (lane < 4 ? w1y : (lane < 6 ? w2x : w2y));
    const int acell = (int)((lane & 1) ? (word >> 16) : (word & 0xffffu));
    if (acell != 0xffff) o[acell] = 1.0f;
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int j = lane + 32 * m;
    if (j >= 3 && j - 3 < nv) {
      const int cell = (int)(r[m].x & 0xffffu);
      o[cells + cell] = __uint_as_float(r[m].y);
      o[2 * cells + cell] = d.lut[r[m].x >> 16];
    }
  }
  for (int j = 128 + lane; j - 3 < nv; j += 32) {  // rare: more than 125 visited cells
    const uint2 q = rec[j];
    const int cell = (int)(q.x & 0xffffu);
    o[cells + cell] = __uint_as_float(q.y);
    o[2 * cells + cell] = d.lut[q.x >> 16];
  }
}

// K3, incremental form (opt-in, rt_env_set_incremental): when the caller hands back the SAME tensor that
// holds this env's previous observation, a step changes at most 2A cells per env — the agents' old cells
// lose their location mark, the new cells get mark, intensity and visit value.  One thread per env.
__global__ void __launch_bounds__(256)
k_patch_incremental(const EnvDev d, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.E) return;
  const uint4 pv = d.prev_cells[e];
  const uint2* rec = d.vrec + (size_t)e * d.rec_stride;
  const uint2 c1 = rec[1], c2 = rec[2];
  float* o = out + (size_t)e * 3 * d.cells;
  const uint32_t prevw[4] = {pv.x, pv.y, pv.z, pv.w}, curw[4] = {c1.x, c1.y, c2.x, c2.y};
#pragma unroll
  for (int a = 0; a < RT_MAX_AGENTS; ++a) {  // clear first, mark afterwards: an agent may enter a cell another one left
    const uint32_t c = (prevw[a >> 1] >> (16 * (a & 1))) & 0xffffu;
    if (c != 0xffffu) o[c] = 0.0f;
  }
#pragma unroll
  for (int a = 0; a < RT_MAX_AGENTS; ++a) {
    const uint32_t c = (curw[a >> 1] >> (16 * (a & 1))) & 0xffffu;
    if (c != 0xffffu) {
      o[c] = 1.0f;
      const uint16_t* crc = d.cellrec + ((size_t)e * d.cells + c) * 4;
      const int sl = crc[0] == (uint16_t)(d.episode[e] & 0xffff) ? (int)crc[2] : 0;  // older episode = not visited
      if (sl != 0) {  // visited this episode (always, after a step; not for a voided env still on its start cell)
        const uint2 r = rec[2 + sl];  // {cell | visits << 16, fp32 intensity}
        o[d.cells + c] = __uint_as_float(r.y);
        o[2 * (size_t)d.cells + c] = d.lut[r.x >> 16];
      }
    }
  }
}

// Materialise the dense visit / intensity maps from the records (inspection / state export in sparse mode).
__global__ void __launch_bounds__(128)
k_densify(const EnvDev d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.E) return;
  const uint2* rec = d.vrec + (size_t)e * d.rec_stride;
  float* inten = d.inten + (size_t)e * d.cells;
  uint16_t* visit = d.visit + (size_t)e * d.cells;
  const int nv = (int)rec[0].x;
  for (int k = 1; k <= nv; ++k) {
    const uint2 r = rec[2 + k];
    inten[r.x & 0xffffu] = __uint_as_float(r.y);
    visit[r.x & 0xffffu] = (uint16_t)(r.x >> 16);
  }
}

// Generic path for any grid side (the reference default size=10 has 100 cells): one cell per thread.
__global__ void __launch_bounds__(256)
k_raster_any(const EnvDev d, float* __restrict__ out, long long n_cells_total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells_total; c += stride) {
    const int e = (int)(c / d.cells);
    const int cell = (int)(c - (long long)e * d.cells);
    float loc = 0.f;
    const int2* agent = reinterpret_cast<const int2*>(d.agent_xy) + (size_t)e * d.A;
    for (int a = 0; a < d.A; ++a) {
      const int2 p = agent[a];
      if (p.y * d.G + p.x == cell) loc = 1.0f;
    }
    float* o = out + (size_t)e * 3 * d.cells + cell;
    o[0] = loc;
    o[d.cells] = d.inten[c];
    o[2 * (size_t)d.cells] = d.lut[d.visit[c]];
  }
}

}  // namespace rt
