// rt_env.cu — host side of the batched environment: handle, device slab, launches, C ABI.
// See include/rt_b200.h for the contract of every entry point.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "rt_env_kernels.cuh"
#include "rt_host.h"

using namespace rt;

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
static inline void rt_cpu_relax() { _mm_pause(); }
#else
static inline void rt_cpu_relax() { std::atomic_signal_fence(std::memory_order_seq_cst); }
#endif

struct rt_env {
  rt_env_cfg cfg;
  EnvDev dev;
  int device = 0;
  int n_sm = 148;
  void* slab = nullptr;  // one allocation, carved at 256 B granularity
  bool slab_vmm = false; // RT_B200_SLAB_COMPRESS=1: the slab came from rt_obs_alloc (compressible memory)
  size_t slab_bytes = 0;
  void* buf_ptr[RT_BUF__COUNT] = {};
  size_t buf_bytes[RT_BUF__COUNT] = {};
  // host-facing variants: device mirrors + pinned staging of the small per-agent arrays
  int32_t* d_actions = nullptr;
  int32_t* d_obs_xy = nullptr;
  int32_t* d_reward = nullptr;
  uint8_t* d_done = nullptr;
  uint8_t* d_mask = nullptr;
  unsigned char* pinned = nullptr;
  size_t pinned_bytes = 0;
  int32_t* d_flag_or = nullptr;
  int64_t launches = 0;
  int epb = 64;
  bool lut_in_smem = true;
  cudaStream_t aux = nullptr;  // host path: results travel back here while the rasteriser streams
  cudaEvent_t ev_fork = nullptr;
  unsigned long long host_seq = 0;   // host path: results of call k have landed when the pinned sequence word reads k
  size_t zc_out_bytes = 256 << 10;   // host path: results up to this size are stored by k_step itself (RT_B200_ZC_OUT_KB)
  int raster_ctas = 32, raster_unroll = 2;  // dense rasteriser: CTAs per SM (grid cap), chunks in flight per thread
  bool incremental = false;                 // rt_env_set_incremental: patch the caller's tensor when it is the same as last time
  const float* last_maps = nullptr;         // tensor that holds this env's latest observation
  bool zerocopy = true;                     // host path: k_step reads pinned caller actions in place
  bool sparse = true;                       // RT_B200_RASTER=dense selects the dense (LSU) rasteriser
  int sparse_ctas = 24;                     // k_raster_sparse_lsu: CTAs (8 warp pipelines each) per SM in the grid (2 resident)
  bool sparse_lsu = true;                   // RT_B200_RASTER=sparse_tma selects the bulk-store variant
  bool small_raster = false;                // L2-resident observation tensor: k_raster_small (RT_B200_RASTER=small forces it)
  bool pdl = true;                          // RT_B200_PDL=0: plain stream-ordered launches of k_step / the rasteriser
  int sparse_hints = 19;                    // bit 0/1: TMA variant L2 hints; bit 4: plain st.global (else .cs) in the LSU variant
};

namespace {

struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  }
};

// Launch with (optionally) the programmatic-stream-serialisation attribute: the kernel's CTAs may be scheduled
// while the previous kernel of the stream drains; the kernel itself calls pdl_wait() before touching its inputs.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl,
                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

bool is_pinned_or_device(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged ||
         at.type == cudaMemoryTypeDevice;
}

// device-side alias of a pinned / managed / device allocation (nullptr for pageable memory)
const void* device_alias(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (at.type == cudaMemoryTypeUnregistered) return nullptr;
  return at.devicePointer;
}

// {sum of returns, sum of episode lengths, agents that reached the source}: integer partial sums per
// thread, warp-reduced, one fp64 atomicAdd per warp (integers well below 2^53: order-independent).
__global__ void __launch_bounds__(256)
k_episode_stats(const int32_t* __restrict__ ep_return, const uint8_t* __restrict__ ep_done,
                const int32_t* __restrict__ tick, int E, int A, double* __restrict__ out) {
  long long ret = 0, len = 0, dn = 0;
  const long long n = (long long)E * A;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    ret += ep_return[i];
    dn += ep_done[i];
    len += tick[i / A];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    ret += __shfl_down_sync(0xffffffffu, ret, off);
    len += __shfl_down_sync(0xffffffffu, len, off);
    dn += __shfl_down_sync(0xffffffffu, dn, off);
  }
  __shared__ long long s_part[3][8];
  if ((threadIdx.x & 31) == 0) {
    s_part[0][threadIdx.x >> 5] = ret;
    s_part[1][threadIdx.x >> 5] = len;
    s_part[2][threadIdx.x >> 5] = dn;
  }
  __syncthreads();
  if (threadIdx.x < 3) {  // one atomic per quantity and CTA (per-warp atomics on one address serialise)
    long long v = 0;
    for (int i = 0; i < 8; ++i) v += s_part[threadIdx.x][i];
    atomicAdd(out + threadIdx.x, (double)v);
  }
}

__global__ void k_init_headers(uint2* __restrict__ vrec, int E, int stride) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  vrec[(size_t)e * stride + 0] = make_uint2(0u, 0u);
  vrec[(size_t)e * stride + 1] = make_uint2(0xffffffffu, 0xffffffffu);
  vrec[(size_t)e * stride + 2] = make_uint2(0xffffffffu, 0xffffffffu);
}

__global__ void k_or_flags(const int32_t* __restrict__ flags, int n, int32_t* __restrict__ out) {
  int v = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v |= flags[i];
  v = __reduce_or_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && v) atomicOr(out, v);
}

void launch_raster_vec(rt_env* h, float* obs_maps, cudaStream_t st) {
  const EnvDev& d = h->dev;
  const long long n_chunks = (long long)d.E * (d.cells >> 3);
  const int threads = 256;
  long long blocks = (n_chunks + threads - 1) / threads;
  const long long cap_blocks = (long long)h->n_sm * h->raster_ctas;
  if (blocks > cap_blocks) blocks = cap_blocks;
  const size_t sm = h->lut_in_smem ? (size_t)(d.cap + 1) * 4 : 0;
#define RT_LAUNCH_RV(U)                                                                       \
  do {                                                                                        \
    if (h->lut_in_smem)                                                                       \
      k_raster_vec<true, U><<<(unsigned)blocks, threads, sm, st>>>(d, obs_maps, n_chunks);   \
    else                                                                                      \
      k_raster_vec<false, U><<<(unsigned)blocks, threads, 0, st>>>(d, obs_maps, n_chunks);   \
  } while (0)
  if (h->raster_unroll >= 4)
    RT_LAUNCH_RV(4);
  else if (h->raster_unroll >= 2)
    RT_LAUNCH_RV(2);
  else
    RT_LAUNCH_RV(1);
#undef RT_LAUNCH_RV
}

int launch_raster(rt_env* h, float* obs_maps, cudaStream_t st) {
  const EnvDev& d = h->dev;
  h->last_maps = obs_maps;
  if (h->sparse && h->small_raster) {
    // small batch: warp per env, zeros streamed while k_step still runs, cells patched after it (k_raster_small)
    RT_CUDA(launch_pdl(k_raster_small, (unsigned)((d.E + 7) / 8), 256u, (size_t)0, st, h->pdl, d, obs_maps));
  } else if (h->sparse && h->sparse_lsu) {
    // visit-table lookups are rare here (a few dozen per env), so the table is read through L1 from global
    // memory: staging it in shared memory per CTA (+ a CTA barrier) was measured slower (profiles/r02a)
    const size_t smb = raster_sl_smem_bytes(d.cells, 0);
    long long nb = (long long)h->n_sm * h->sparse_ctas;
    if (nb * kSlWarps > d.E) nb = (d.E + kSlWarps - 1) / kSlWarps;
    RT_CUDA(launch_pdl(k_raster_sparse_lsu<false>, (unsigned)nb, kSlThreads, smb, st, h->pdl, d, obs_maps,
                       h->sparse_hints & 16));
  } else if (h->sparse) {
    const size_t smb = raster_sparse_smem_bytes(d.cells, d.rec_prefix, h->lut_in_smem ? d.cap + 1 : 0);
    long long nb = (long long)h->n_sm;  // 16 tiles fill an SM's shared memory: one persistent CTA each
    if (nb * kSpWarps > d.E) nb = (d.E + kSpWarps - 1) / kSpWarps;
    if (h->lut_in_smem)
      k_raster_sparse<true><<<(unsigned)nb, kSpThreads, smb, st>>>(d, obs_maps, h->sparse_hints);
    else
      k_raster_sparse<false><<<(unsigned)nb, kSpThreads, smb, st>>>(d, obs_maps, h->sparse_hints);
  } else if ((d.cells & 7) == 0) {
    launch_raster_vec(h, obs_maps, st);
  } else {
    const long long n = (long long)d.E * d.cells;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)h->n_sm * 32) blocks = (long long)h->n_sm * 32;
    k_raster_any<<<(unsigned)blocks, 256, 0, st>>>(d, obs_maps, n);
  }
  h->launches += 1;
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

// full rasterisation, or — opted in and handed the tensor that already holds the previous observation —
// the incremental patch of the cells the step changed
int raster_or_patch(rt_env* h, float* obs_maps, cudaStream_t st) {
  const EnvDev& d = h->dev;
  if (h->incremental && obs_maps == h->last_maps) {
    k_patch_incremental<<<(unsigned)((d.E + 255) / 256), 256, 0, st>>>(d, obs_maps);
    h->launches += 1;
    RT_CUDA_LAUNCH_CHECK();
    return RT_OK;
  }
  return launch_raster(h, obs_maps, st);
}

int launch_step(rt_env* h, const int32_t* actions, int32_t* obs_xy, float* obs_maps, int32_t* reward,
                uint8_t* done, cudaStream_t st, bool raster_follows = false, bool publish_to_host = false) {
  EnvDev d = h->dev;
  d.host_seq = h->host_seq;
  if (publish_to_host) {  // small-batch host path: the kernel itself returns flag word and sequence number
    d.host_flag_out = static_cast<volatile int32_t*>(const_cast<void*>(device_alias(h->pinned + h->pinned_bytes - 16)));
    d.host_seq_out = static_cast<volatile unsigned long long*>(const_cast<void*>(device_alias(h->pinned + h->pinned_bytes - 8)));
  }
  const int epb = h->epb;
  const unsigned blocks = (unsigned)((d.E + epb - 1) / epb);
  RT_CUDA(launch_pdl(k_step, blocks, (unsigned)(epb * d.A), step_smem_bytes(epb, d.A), st, h->pdl, d, actions, obs_xy,
                     reward, done));
  h->launches += 1;
  RT_CUDA_LAUNCH_CHECK();
  if (obs_maps) return raster_or_patch(h, obs_maps, st);
  if (!raster_follows) h->last_maps = nullptr;  // a step without rasterisation leaves every tensor stale
  return RT_OK;
}

int launch_reset(rt_env* h, const uint8_t* mask, int32_t* obs_xy, float* obs_maps, cudaStream_t st) {
  const EnvDev& d = h->dev;
  k_reset_state<<<(unsigned)((d.E + 127) / 128), 128, 0, st>>>(d, mask, obs_xy);
  k_zero_maps<<<(unsigned)std::min<long long>(d.E, (long long)h->n_sm * 8), 256, 0, st>>>(d, mask);
  h->launches += 2;
  RT_CUDA_LAUNCH_CHECK();
  if (obs_maps) return launch_raster(h, obs_maps, st);
  h->last_maps = nullptr;  // state changed without a rasterisation: every tensor is stale
  return RT_OK;
}

}  // namespace

extern "C" {

int rt_abi_version(void) { return RT_ABI_VERSION; }

const char* rt_status_string(int status) {
  switch (status) {
    case RT_OK: return "ok";
    case RT_ERR_BAD_ARG: return "bad argument";
    case RT_ERR_CUDA: return rt_last_cuda_error();
    case RT_ERR_NOMEM: return "out of memory";
    case RT_ERR_BAD_ACTION: return "action outside 1..4";
    case RT_ERR_ASSERT: return "reference assertion";
    case RT_ERR_KEY: return "key does not exist";
    default: return "unknown status";
  }
}

int rt_env_create(const rt_env_cfg* cfg, rt_env** out) {
  if (!cfg || !out) return RT_ERR_BAD_ARG;
  *out = nullptr;
  const long long cap = (long long)cfg->max_steps * cfg->n_agents;
  if (cfg->n_envs < 1 || cfg->n_agents < 1 || cfg->n_agents > RT_MAX_AGENTS || cfg->grid < 2 ||
      cfg->grid > 255 || cfg->max_steps < 1 || cap < 2 || cap > 65534 || cfg->poly_min < 0 ||
      cfg->poly_max < cfg->poly_min || cfg->poly_max > RT_MAX_POLYS || cfg->lognorm_step < 1 ||
      !(cfg->cell_pitch_cm > 0.0) || !(cfg->min_dist_cm > 0.0) || !(cfg->transmission >= 0.0) ||
      (cfg->poly_max > 0 && (cfg->rect_min < 1 || cfg->rect_max < cfg->rect_min)) ||
      !(cfg->src_lo >= 0.0) || !(cfg->src_hi >= cfg->src_lo) || !(cfg->bkg_lo >= 0.0) || !(cfg->bkg_hi >= cfg->bkg_lo) ||
      !std::isfinite(cfg->src_hi) || !std::isfinite(cfg->bkg_hi))
    return RT_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    rt_note_cuda_error(cudaGetLastError(), "cudaGetDeviceCount");
    return RT_ERR_CUDA;  // no CPU fallback, by design
  }
  rt_env* h = new (std::nothrow) rt_env();
  if (!h) return RT_ERR_NOMEM;
  h->cfg = *cfg;
  RT_CUDA_OR(cudaGetDevice(&h->device), { delete h; });
  cudaDeviceProp prop;
  RT_CUDA_OR(cudaGetDeviceProperties(&prop, h->device), { delete h; });
  h->n_sm = prop.multiProcessorCount;

  const size_t E = (size_t)cfg->n_envs, A = (size_t)cfg->n_agents;
  const size_t cells = (size_t)cfg->grid * cfg->grid;
  // visited-cell records: 3 header slots + one record per distinct visited cell, rounded up to even
  const size_t max_rec = (size_t)cap < cells ? (size_t)cap : cells;
  const int rec_stride = (int)((3 + max_rec + 1) & ~(size_t)1);
  const size_t sz[RT_BUF__COUNT] = {
      /*AGENT_XY*/ E * A * 8,   /*PREV_DIST*/ E * A * 4, /*START_XY*/ E * A * 8, /*SRC_XY*/ E * 8,
      /*SRC_INT*/ E * 8,        /*BKG*/ E * 8,           /*N_POLY*/ E * 4,       /*EDGES*/ E * RT_MAX_EDGES * 16,
      /*TICK*/ E * 4,           /*EPISODE*/ E * 4,       /*VISIT*/ E * cells * 2, /*INTENSITY*/ E * cells * 4,
      /*CELLREC*/ E * cells * 8,   /*LOG*/ (size_t)cap * E * 32, /*unused*/ 0,       /*WELFORD*/ E * 48,
      /*WELFORD_N*/ E * 4,      /*EST_MAXMIN*/ E * 16,   /*FLAGS*/ E * 4,        /*LAST_COUNT*/ E * A * 4,
      /*LAST_LOS*/ E * A,       /*LAST_LAMBDA*/ E * A * 8, /*LAST_EST*/ E * A * 8, /*LAST_Z*/ E * A * 8,
      /*LAST_VALUE*/ E * A * 8, /*VISIT_LUT*/ ((size_t)cap + 1) * 4, /*EP_RETURN*/ E * A * 4,
      /*EP_DONE*/ E * A,        /*reserved*/ 0,  /*VREC*/ E * (size_t)rec_stride * 8};
  Carver cv;
  size_t off[RT_BUF__COUNT];
  for (int i = 0; i < RT_BUF__COUNT; ++i) off[i] = cv.take(sz[i]);
  // host path: obs_xy | reward | done packed back to back, so that ONE copy returns them when the caller's
  // buffers are laid out the same way
  const size_t o_act = cv.take(E * A * 4), o_xy = cv.take(E * A * 13), o_rw = o_xy + E * A * 8,
               o_dn = o_xy + E * A * 12, o_mask = cv.take(E), o_for = cv.take(32), o_prev = cv.take(E * 16);
  h->slab_bytes = cv.off;
  if (const char* ev = std::getenv("RT_B200_SLAB_COMPRESS")) h->slab_vmm = std::atoi(ev) != 0;
  if (h->slab_vmm) {
    const int rc = rt_obs_alloc(h->slab_bytes, &h->slab, nullptr);
    if (rc != RT_OK) {
      delete h;
      return rc;
    }
  } else if (cudaMalloc(&h->slab, h->slab_bytes) != cudaSuccess) {
    rt_note_cuda_error(cudaGetLastError(), "cudaMalloc(slab)");
    delete h;
    return RT_ERR_NOMEM;
  }
  RT_CUDA_OR(cudaMemset(h->slab, 0, h->slab_bytes), { rt_env_destroy(h); });
  unsigned char* base = (unsigned char*)h->slab;
  for (int i = 0; i < RT_BUF__COUNT; ++i) {
    h->buf_ptr[i] = sz[i] ? base + off[i] : nullptr;
    h->buf_bytes[i] = sz[i];
  }
  h->d_actions = (int32_t*)(base + o_act);
  h->d_obs_xy = (int32_t*)(base + o_xy);
  h->d_reward = (int32_t*)(base + o_rw);
  h->d_done = (uint8_t*)(base + o_dn);
  h->d_mask = (uint8_t*)(base + o_mask);
  h->d_flag_or = (int32_t*)(base + o_for);

  // pinned staging: actions | obs_xy | reward | done | mask
  h->pinned_bytes = ((E * A * 4 + E * A * 8 + E * A * 4 + E * A + E + 15) & ~(size_t)15) + 64;  // tail: flag word, sequence word
  RT_CUDA_OR(cudaHostAlloc((void**)&h->pinned, h->pinned_bytes, cudaHostAllocDefault), { rt_env_destroy(h); });
  std::memset(h->pinned + h->pinned_bytes - 64, 0, 64);

  EnvDev& d = h->dev;
  d.E = cfg->n_envs; d.A = cfg->n_agents; d.G = cfg->grid; d.cells = (int)cells;
  d.T = cfg->max_steps; d.cap = (int)cap;
  d.gid_base = cfg->env_id_base; d.seed = cfg->seed;
  d.pitch2 = cfg->cell_pitch_cm * cfg->cell_pitch_cm;
  d.dmin2 = cfg->min_dist_cm * cfg->min_dist_cm;
  d.transmission = cfg->transmission;
  d.src_lo = cfg->src_lo; d.src_span = cfg->src_hi - cfg->src_lo;
  d.bkg_lo = cfg->bkg_lo; d.bkg_span = cfg->bkg_hi - cfg->bkg_lo;
  d.poly_min = cfg->poly_min; d.poly_max = cfg->poly_max;
  d.rect_min = cfg->rect_min; d.rect_max = cfg->rect_max; d.fixed = cfg->fixed_scenario;
  d.agent_xy = (int32_t*)h->buf_ptr[RT_BUF_AGENT_XY];
  d.prev_dist = (int32_t*)h->buf_ptr[RT_BUF_PREV_DIST];
  d.start_xy = (int32_t*)h->buf_ptr[RT_BUF_START_XY];
  d.src_xy = (int32_t*)h->buf_ptr[RT_BUF_SRC_XY];
  d.src_int = (double*)h->buf_ptr[RT_BUF_SRC_INT];
  d.bkg = (double*)h->buf_ptr[RT_BUF_BKG];
  d.n_poly = (int32_t*)h->buf_ptr[RT_BUF_N_POLY];
  d.edges = (float*)h->buf_ptr[RT_BUF_EDGES];
  d.tick = (int32_t*)h->buf_ptr[RT_BUF_TICK];
  d.episode = (int32_t*)h->buf_ptr[RT_BUF_EPISODE];
  d.visit = (uint16_t*)h->buf_ptr[RT_BUF_VISIT];
  d.inten = (float*)h->buf_ptr[RT_BUF_INTENSITY];
  d.cellrec = (uint16_t*)h->buf_ptr[RT_BUF_CELLREC];
  d.log = (uint4*)h->buf_ptr[RT_BUF_LOG];
  d.welford = (double*)h->buf_ptr[RT_BUF_WELFORD];
  d.welford_n = (int32_t*)h->buf_ptr[RT_BUF_WELFORD_N];
  d.est_mm = (double*)h->buf_ptr[RT_BUF_EST_MAXMIN];
  d.flags = (int32_t*)h->buf_ptr[RT_BUF_FLAGS];
  d.last_count = (int32_t*)h->buf_ptr[RT_BUF_LAST_COUNT];
  d.last_los = (uint8_t*)h->buf_ptr[RT_BUF_LAST_LOS];
  d.last_lambda = (double*)h->buf_ptr[RT_BUF_LAST_LAMBDA];
  d.last_est = (double*)h->buf_ptr[RT_BUF_LAST_EST];
  d.last_z = (double*)h->buf_ptr[RT_BUF_LAST_Z];
  d.last_value = (double*)h->buf_ptr[RT_BUF_LAST_VALUE];
  d.lut = (float*)h->buf_ptr[RT_BUF_VISIT_LUT];
  d.ep_return = (int32_t*)h->buf_ptr[RT_BUF_EP_RETURN];
  d.ep_done = (uint8_t*)h->buf_ptr[RT_BUF_EP_DONE];
  d.vrec = (uint2*)h->buf_ptr[RT_BUF_VREC];
  d.rec_stride = rec_stride;
  d.rec_prefix = rec_stride < 128 ? rec_stride : 128;  // 1 KB per env covers 125 visited cells; the rest is read directly
  d.rollout = nullptr;
  d.step_flags = h->d_flag_or + 4;  // {flags of odd calls, of even calls, sequence number (8 B)}, 16-byte aligned
  d.host_seq = 0;
  d.host_flag_out = nullptr;
  d.host_seq_out = nullptr;
  d.cta_counter = reinterpret_cast<unsigned*>(h->d_flag_or + 2);
  d.prev_cells = (uint4*)(base + o_prev);
  d.inj = nullptr;
  d.inj_k = 0;

  // record headers before the first reset: no visited cells, no agents
  k_init_headers<<<(unsigned)((E + 255) / 256), 256>>>(d.vrec, (int)E, rec_stride);
  RT_CUDA_OR(cudaGetLastError(), { rt_env_destroy(h); });
  // episode counters start at -1 so the first reset opens episode 0; Welford std starts at 1
  {
    std::vector<int32_t> ep(E, -1);
    RT_CUDA_OR(cudaMemcpy(d.episode, ep.data(), E * 4, cudaMemcpyHostToDevice),
               { rt_env_destroy(h); });
    std::vector<double> w(E * 6, 0.0);
    for (size_t i = 0; i < E; ++i) w[i * 6 + 3] = 1.0;
    RT_CUDA_OR(cudaMemcpy(d.welford, w.data(), E * 48, cudaMemcpyHostToDevice), { rt_env_destroy(h); });
  }
  // a8: BensLogNormalizer table over every reachable visit count, max = T*A
  // (normalize.py:64-65,110).  math.log(x, base) is log(x)/log(base); LUT[0] = 0 keeps
  // never-visited cells dark.
  {
    std::vector<float> lut((size_t)cap + 1, 0.0f);
    const double mx = (double)cap, st = (double)cfg->lognorm_step, lb = std::log(mx);
    for (long long v = 1; v <= cap; ++v) {
      const double num = std::log(st + (double)v) / lb;
      const double den = std::log(st * mx) / lb;
      const double r = num * 1.0 / den;
      lut[(size_t)v] = (r >= 0.0 && r <= 1.0) ? (float)r : 1.0f;
    }
    RT_CUDA_OR(cudaMemcpy(d.lut, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice),
               { rt_env_destroy(h); });
  }
  h->lut_in_smem = ((size_t)cap + 1) * 4 <= 32 * 1024;
  // envs per CTA of k_step: <= 192 threads, and enough CTAs to cover the SMs twice when E allows
  int epb = 192 / cfg->n_agents;
  if (epb > 32) epb = 32;  // measured: 32 envs per CTA beats 64 and 16 at c3 (profiles/r01m)
  while (epb > 8 && (cfg->n_envs + epb - 1) / epb < 2 * h->n_sm) epb >>= 1;
  if (epb * cfg->n_agents < 32) epb = (32 + cfg->n_agents - 1) / cfg->n_agents;
  // Small batches (configs[1]: 4,096 x 1) are one latency chain per warp and leave most SMs idle: a warp runs the
  // UNION of its lanes' sampler passes and bucket hops, so fewer envs per warp shorten the chain — until the number
  // of CTAs itself costs more (measured at c2, CUDA-graph loop, envs per CTA -> us per step: 32 15.1, 16 14.6,
  // 8 13.8, 4 13.2, 2 17.8, 1 19.2; profiles/r2g_*): about a thousand CTAs.
  if ((long long)cfg->n_envs * cfg->n_agents <= 16384)
    epb = std::min(32, std::max(1, (cfg->n_envs + 1023) / 1024));
  h->epb = epb;
  if (const char* ev = std::getenv("RT_B200_EPB")) {  // tuning knob: envs per CTA of the step kernels
    const int v = std::atoi(ev);
    if (v >= 1 && v * cfg->n_agents <= 192) h->epb = v;
  }
  if (const char* ev = std::getenv("RT_B200_ZEROCOPY")) h->zerocopy = std::atoi(ev) != 0;
  if (const char* ev = std::getenv("RT_B200_ZC_OUT_KB")) h->zc_out_bytes = (size_t)std::max(0, std::atoi(ev)) << 10;

  if (const char* ev = std::getenv("RT_B200_RASTER_CTAS")) h->raster_ctas = std::max(1, std::atoi(ev));
  if (const char* ev = std::getenv("RT_B200_RASTER_UNROLL")) h->raster_unroll = std::max(1, std::atoi(ev));
  if (const char* ev = std::getenv("RT_B200_SPARSE_CTAS")) h->sparse_ctas = std::max(1, std::atoi(ev));
  if (const char* ev = std::getenv("RT_B200_SPARSE_HINTS")) h->sparse_hints = std::atoi(ev);
  if (const char* ev = std::getenv("RT_B200_PDL")) h->pdl = std::atoi(ev) != 0;
  if (const char* ev = std::getenv("RT_B200_RASTER")) {
    h->sparse = std::strcmp(ev, "dense") != 0;
    h->sparse_lsu = std::strcmp(ev, "sparse_tma") != 0;
  }
  {
    // the sparse rasteriser composes whole [3][G][G] tiles in shared memory: needs cells % 8 == 0, cell
    // indices that fit 16 bits and kSpStages tiles within the opt-in shared memory; otherwise dense
    const size_t smb = raster_sparse_smem_bytes((int)cells, d.rec_prefix, h->lut_in_smem ? (int)cap + 1 : 0);
    if ((cells & 7) != 0 || cells > 65535 || smb > (size_t)prop.sharedMemPerBlockOptin) {
      h->sparse = false;
    } else {
      RT_CUDA_OR(cudaFuncSetAttribute(k_raster_sparse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb),
                 { rt_env_destroy(h); });
      RT_CUDA_OR(cudaFuncSetAttribute(k_raster_sparse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb),
                 { rt_env_destroy(h); });
      const size_t sml = raster_sl_smem_bytes((int)cells, h->lut_in_smem ? (int)cap + 1 : 0);
      RT_CUDA_OR(cudaFuncSetAttribute(k_raster_sparse_lsu<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml),
                 { rt_env_destroy(h); });
      RT_CUDA_OR(cudaFuncSetAttribute(k_raster_sparse_lsu<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml),
                 { rt_env_destroy(h); });
    }
  }
  // small batches (observation tensor well inside the 126 MB L2): the warp-per-env zero-then-patch rasteriser
  h->small_raster = h->sparse && (cells & 3) == 0 && E * cells * 12 <= (size_t)64 << 20;
  if (const char* ev = std::getenv("RT_B200_RASTER")) {
    if (std::strcmp(ev, "small") == 0) h->small_raster = h->sparse && (cells & 3) == 0;
    else if (std::strcmp(ev, "auto") != 0) h->small_raster = false;  // an explicit choice of another variant
  }
  d.live_dense = h->sparse ? 0 : 1;
  RT_CUDA_OR(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking), { rt_env_destroy(h); });
  RT_CUDA_OR(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming), { rt_env_destroy(h); });
  *out = h;
  return RT_OK;
}

int rt_env_destroy(rt_env* h) {
  if (!h) return RT_OK;
  RtDeviceGuard on_device(h->device);
  if (h->aux) {
    cudaStreamSynchronize(h->aux);
    cudaStreamDestroy(h->aux);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->slab && h->slab_vmm) rt_obs_free(h->slab);
  else if (h->slab) cudaFree(h->slab);
  if (h->pinned) cudaFreeHost(h->pinned);
  delete h;
  return RT_OK;
}

int rt_env_set_scenario(rt_env* h, const int32_t* src_xy, const double* src_int, const double* bkg,
                        const int32_t* start_xy, const int32_t* n_poly, const float* edges) {
  if (!h) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  const size_t E = (size_t)h->dev.E, A = (size_t)h->dev.A;
  if (src_xy)
    for (size_t i = 0; i < E * 2; ++i)
      if (src_xy[i] < 0 || src_xy[i] >= h->dev.G) return RT_ERR_BAD_ARG;
  if (start_xy)
    for (size_t i = 0; i < E * A * 2; ++i)
      if (start_xy[i] < 0 || start_xy[i] >= h->dev.G) return RT_ERR_BAD_ARG;
  if (n_poly)
    for (size_t i = 0; i < E; ++i)
      if (n_poly[i] < 0 || n_poly[i] > RT_MAX_POLYS) return RT_ERR_BAD_ARG;
  if (src_xy) RT_CUDA(cudaMemcpy(h->dev.src_xy, src_xy, E * 8, cudaMemcpyHostToDevice));
  if (src_int) RT_CUDA(cudaMemcpy(h->dev.src_int, src_int, E * 8, cudaMemcpyHostToDevice));
  if (bkg) RT_CUDA(cudaMemcpy(h->dev.bkg, bkg, E * 8, cudaMemcpyHostToDevice));
  if (start_xy) RT_CUDA(cudaMemcpy(h->dev.start_xy, start_xy, E * A * 8, cudaMemcpyHostToDevice));
  if (n_poly) RT_CUDA(cudaMemcpy(h->dev.n_poly, n_poly, E * 4, cudaMemcpyHostToDevice));
  if (edges) RT_CUDA(cudaMemcpy(h->dev.edges, edges, E * RT_MAX_EDGES * 16, cudaMemcpyHostToDevice));
  return RT_OK;
}

int rt_env_reset(rt_env* h, const uint8_t* mask_dev, int32_t* obs_xy_dev, float* obs_maps_dev,
                 void* stream) {
  if (!h) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  RtRange nvtx_range("rt_env_reset");
  return launch_reset(h, mask_dev, obs_xy_dev, obs_maps_dev, (cudaStream_t)stream);
}

int rt_env_step(rt_env* h, const int32_t* actions_dev, int32_t* obs_xy_dev, float* obs_maps_dev,
                int32_t* reward_dev, uint8_t* done_dev, void* stream) {
  if (!h || !actions_dev || !obs_xy_dev || !reward_dev || !done_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  RtRange nvtx_range("rt_env_step");
  return launch_step(h, actions_dev, obs_xy_dev, obs_maps_dev, reward_dev, done_dev,
                     (cudaStream_t)stream);
}

int rt_env_reset_host(rt_env* h, const uint8_t* mask_host, int32_t* obs_xy_host, float* obs_maps_dev,
                      void* stream) {
  if (!h) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  RtRange nvtx_range("rt_env_reset_host");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t E = (size_t)h->dev.E, A = (size_t)h->dev.A;
  const uint8_t* mask_dev = nullptr;
  if (mask_host) {
    unsigned char* pm = h->pinned + E * A * 17;
    const void* srcp = mask_host;
    if (!is_pinned_or_device(mask_host)) {
      std::memcpy(pm, mask_host, E);
      srcp = pm;
    }
    RT_CUDA(cudaMemcpyAsync(h->d_mask, srcp, E, cudaMemcpyDefault, st));
    mask_dev = h->d_mask;
  }
  // k_reset_state writes the start positions; they travel back on the auxiliary stream while the map clears and the
  // rasteriser run on `st`, and the call returns when they have landed (as rt_env_step_host does)
  const EnvDev& d = h->dev;
  k_reset_state<<<(unsigned)((d.E + 127) / 128), 128, 0, st>>>(d, mask_dev, h->d_obs_xy);
  h->launches += 1;
  RT_CUDA_LAUNCH_CHECK();
  bool direct = false;
  if (obs_xy_host) {
    RT_CUDA(cudaEventRecord(h->ev_fork, st));
    RT_CUDA(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
    direct = is_pinned_or_device(obs_xy_host);
    void* dst = direct ? (void*)obs_xy_host : (void*)(h->pinned + E * A * 4);
    RT_CUDA(cudaMemcpyAsync(dst, h->d_obs_xy, E * A * 8, cudaMemcpyDefault, h->aux));
  }
  k_zero_maps<<<(unsigned)std::min<long long>(d.E, (long long)h->n_sm * 8), 256, 0, st>>>(d, mask_dev);
  h->launches += 1;
  RT_CUDA_LAUNCH_CHECK();
  if (obs_maps_dev) {
    const int rc = launch_raster(h, obs_maps_dev, st);
    if (rc != RT_OK) return rc;
  } else {
    h->last_maps = nullptr;
  }
  if (obs_xy_host) {
    RT_CUDA(cudaStreamSynchronize(h->aux));
    if (!direct) std::memcpy(obs_xy_host, h->pinned + E * A * 4, E * A * 8);
  }
  if (mask_host) RT_CUDA(cudaStreamSynchronize(st));  // the staged mask must stay put until the kernels have read it
  return RT_OK;
}

int rt_env_step_host(rt_env* h, const int32_t* actions_host, int32_t* obs_xy_host, int32_t* reward_host,
                     uint8_t* done_host, float* obs_maps_dev, void* stream) {
  if (!h || !actions_host || !obs_xy_host || !reward_host || !done_host) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  RtRange nvtx_range("rt_env_step_host");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)h->dev.E * h->dev.A;
  unsigned char* p_act = h->pinned;
  unsigned char* p_xy = h->pinned + n * 4;
  unsigned char* p_rw = h->pinned + n * 12;
  unsigned char* p_dn = h->pinned + n * 16;
  // tail of the pinned block: {flags of odd calls, flags of even calls, sequence number} as the device keeps them
  volatile int32_t* p_fl = reinterpret_cast<volatile int32_t*>(h->pinned + h->pinned_bytes - 16);
  volatile unsigned long long* p_seq = reinterpret_cast<volatile unsigned long long*>(h->pinned + h->pinned_bytes - 8);
  // Invalid actions are caught on the device: the env's step is voided (the reference raises before any
  // state change, simple_gridworld.py:138) and a flag word travels back with the results.
  // Pinned (mapped) caller memory is read by k_step directly over PCIe — no separate H2D copy in front
  // of the kernel; pageable memory is staged and copied.  RT_B200_ZEROCOPY=0 forces the copy.
  const int32_t* actions_dev = h->d_actions;
  const void* alias = h->zerocopy ? device_alias(actions_host) : nullptr;
  if (alias != nullptr) {
    actions_dev = static_cast<const int32_t*>(alias);
  } else {
    const void* a_src = actions_host;
    if (!is_pinned_or_device(actions_host)) {
      std::memcpy(p_act, actions_host, n * 4);
      a_src = p_act;
    }
    RT_CUDA(cudaMemcpyAsync(h->d_actions, a_src, n * 4, cudaMemcpyDefault, st));
  }
  // k_step (it stamps this call's sequence number next to the flag word and clears the word of the NEXT call: no
  // memset), then the small results travel back on the auxiliary stream WHILE the rasteriser streams on `st`.
  const unsigned long long seq = ++h->host_seq;
  const bool dx = is_pinned_or_device(obs_xy_host), dr = is_pinned_or_device(reward_host),
             dd = is_pinned_or_device(done_host);
  // Small batches: no copy engine at all.  k_step stores obs_xy / reward / done straight into mapped host memory (the
  // caller's pinned buffers, or the handle's staging) and its last CTA publishes the flag word and the sequence
  // number: two launches and a poll per step.  (A copy-engine request costs ~6 us of latency whatever its size —
  // profiles/r2b_k4b_probe: 16 bytes 6.2 us — so for a 53 KB result it IS the step; above ~256 KB the copy engine's
  // bandwidth wins and keeps the stores off k_step's critical path.)
  if (h->zerocopy && n * 13 <= h->zc_out_bytes) {
    int32_t* o_xy = static_cast<int32_t*>(const_cast<void*>(device_alias(dx ? (void*)obs_xy_host : (void*)p_xy)));
    int32_t* o_rw = static_cast<int32_t*>(const_cast<void*>(device_alias(dr ? (void*)reward_host : (void*)p_rw)));
    uint8_t* o_dn = static_cast<uint8_t*>(const_cast<void*>(device_alias(dd ? (void*)done_host : (void*)p_dn)));
    if (o_xy && o_rw && o_dn) {
      int rc = launch_step(h, actions_dev, o_xy, nullptr, o_rw, o_dn, st, obs_maps_dev != nullptr, true);
      if (rc != RT_OK) return rc;
      if (obs_maps_dev) {
        rc = raster_or_patch(h, obs_maps_dev, st);
        if (rc != RT_OK) return rc;
      }
      for (unsigned spins = 0; *p_seq != seq;) {
        rt_cpu_relax();
        if ((++spins & 0xfffffu) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) {
          if (*p_seq == seq) break;
          RT_CUDA(cudaStreamSynchronize(st));
          if (*p_seq != seq) {
            rt_note_error_text("rt_env_step_host: the results' sequence number did not arrive");
            return RT_ERR_CUDA;
          }
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      if (!dx) std::memcpy(obs_xy_host, p_xy, n * 8);
      if (!dr) std::memcpy(reward_host, p_rw, n * 4);
      if (!dd) std::memcpy(done_host, p_dn, n);
      if (p_fl[seq & 1] & RT_FLAG_BAD_ACTION) return RT_ERR_BAD_ACTION;
      return RT_OK;
    }
  }
  int rc = launch_step(h, actions_dev, h->d_obs_xy, nullptr, h->d_reward, h->d_done, st, obs_maps_dev != nullptr);
  if (rc != RT_OK) return rc;
  RT_CUDA(cudaEventRecord(h->ev_fork, st));
  RT_CUDA(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
  unsigned char* xy_b = reinterpret_cast<unsigned char*>(obs_xy_host);
  if (dx && dr && dd && reinterpret_cast<unsigned char*>(reward_host) == xy_b + n * 8 && done_host == xy_b + n * 12) {
    // the caller's three buffers are one pinned block laid out like the device's: ONE copy
    RT_CUDA(cudaMemcpyAsync(obs_xy_host, h->d_obs_xy, n * 13, cudaMemcpyDefault, h->aux));
  } else if (!dx && !dr && !dd) {
    RT_CUDA(cudaMemcpyAsync(p_xy, h->d_obs_xy, n * 13, cudaMemcpyDefault, h->aux));  // staging has the same layout
  } else {
    RT_CUDA(cudaMemcpyAsync(dx ? (void*)obs_xy_host : (void*)p_xy, h->d_obs_xy, n * 8, cudaMemcpyDefault, h->aux));
    RT_CUDA(cudaMemcpyAsync(dr ? (void*)reward_host : (void*)p_rw, h->d_reward, n * 4, cudaMemcpyDefault, h->aux));
    RT_CUDA(cudaMemcpyAsync(dd ? (void*)done_host : (void*)p_dn, h->d_done, n, cudaMemcpyDefault, h->aux));
  }
  // flags + sequence number, behind the results in stream order: when the number has landed, so has everything else
  RT_CUDA(cudaMemcpyAsync(const_cast<int32_t*>(p_fl), h->dev.step_flags, 16, cudaMemcpyDeviceToHost, h->aux));
  if (obs_maps_dev) {
    rc = raster_or_patch(h, obs_maps_dev, st);
    if (rc != RT_OK) return rc;
  }
  // Return when the step's results are in host memory: poll the sequence word (no driver call on the wait).  The
  // rasteriser keeps streaming on `st`: whatever the host does before its next call hides under it, and everything
  // enqueued on `st` later is ordered behind it.
  for (unsigned spins = 0; *p_seq != seq;) {
    rt_cpu_relax();
    if ((++spins & 0xfffffu) == 0 && cudaStreamQuery(h->aux) != cudaErrorNotReady) {  // finished or failed
      if (*p_seq == seq) break;
      RT_CUDA(cudaStreamSynchronize(h->aux));
      if (*p_seq != seq) {
        rt_note_error_text("rt_env_step_host: the results' sequence number did not arrive");
        return RT_ERR_CUDA;
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (!dx) std::memcpy(obs_xy_host, p_xy, n * 8);
  if (!dr) std::memcpy(reward_host, p_rw, n * 4);
  if (!dd) std::memcpy(done_host, p_dn, n);
  if (p_fl[seq & 1] & RT_FLAG_BAD_ACTION) return RT_ERR_BAD_ACTION;
  return RT_OK;
}

int rt_env_rasterize(rt_env* h, float* obs_maps_dev, void* stream) {
  if (!h || !obs_maps_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  RtRange nvtx_range("rt_env_rasterize");
  return launch_raster(h, obs_maps_dev, (cudaStream_t)stream);
}

int rt_env_set_incremental(rt_env* h, int32_t enable) {
  if (!h) return RT_ERR_BAD_ARG;
  h->incremental = enable != 0;
  h->last_maps = nullptr;
  return RT_OK;
}

int rt_env_inject_uniforms(rt_env* h, const double* u_dev, int32_t k_per_agent) {
  if (!h || (u_dev && k_per_agent < 1)) return RT_ERR_BAD_ARG;
  h->dev.inj = u_dev;
  h->dev.inj_k = u_dev ? k_per_agent : 0;
  return RT_OK;
}

int rt_env_set_rollout(rt_env* h, float* readings_dev) {
  if (!h) return RT_ERR_BAD_ARG;
  h->dev.rollout = readings_dev;
  return RT_OK;
}

int rt_env_episode_stats(rt_env* h, double* stats_dev, void* stream) {
  if (!h || !stats_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  RtRange nvtx_range("rt_env_episode_stats");
  cudaStream_t st = (cudaStream_t)stream;
  RT_CUDA(cudaMemsetAsync(stats_dev, 0, 3 * sizeof(double), st));
  long long blocks = ((long long)h->dev.E * h->dev.A + 255) / 256;
  if (blocks > h->n_sm * 8) blocks = h->n_sm * 8;
  k_episode_stats<<<(unsigned)blocks, 256, 0, st>>>(h->dev.ep_return, h->dev.ep_done, h->dev.tick, h->dev.E,
                                                  h->dev.A, stats_dev);
  h->launches += 1;
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

namespace {
struct StateHeader {  // 64 bytes in front of the slab image
  uint64_t magic;
  int32_t abi, E, A, G, T, reserved;
  uint64_t slab_bytes;
  unsigned char pad[24];
};
static_assert(sizeof(StateHeader) == 64, "state header is 64 bytes");
constexpr uint64_t kStateMagic = 0x3142544152545352ull;  // "RSTRATB1"
}  // namespace

int rt_env_state_bytes(const rt_env* h, size_t* bytes) {
  if (!h || !bytes) return RT_ERR_BAD_ARG;
  *bytes = sizeof(StateHeader) + h->slab_bytes;
  return RT_OK;
}

int rt_env_get_state(rt_env* h, void* dst, void* stream) {
  if (!h || !dst) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  StateHeader hd = {};
  hd.magic = kStateMagic;
  hd.abi = RT_ABI_VERSION;
  hd.E = h->dev.E; hd.A = h->dev.A; hd.G = h->dev.G; hd.T = h->dev.T;
  hd.slab_bytes = h->slab_bytes;
  // the header goes through a synchronous copy (it lives on this stack frame), the slab in stream order
  RT_CUDA(cudaStreamSynchronize(st));
  RT_CUDA(cudaMemcpy(dst, &hd, sizeof hd, cudaMemcpyDefault));
  RT_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(dst) + sizeof hd, h->slab, h->slab_bytes, cudaMemcpyDefault, st));
  return RT_OK;
}

int rt_env_set_state(rt_env* h, const void* src, void* stream) {
  if (!h || !src) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  StateHeader hd;
  RT_CUDA(cudaStreamSynchronize(st));
  RT_CUDA(cudaMemcpy(&hd, src, sizeof hd, cudaMemcpyDefault));
  if (hd.magic != kStateMagic || hd.abi != RT_ABI_VERSION || hd.E != h->dev.E || hd.A != h->dev.A ||
      hd.G != h->dev.G || hd.T != h->dev.T || hd.slab_bytes != h->slab_bytes)
    return RT_ERR_BAD_ARG;
  RT_CUDA(cudaMemcpyAsync(h->slab, static_cast<const unsigned char*>(src) + sizeof hd, h->slab_bytes,
                          cudaMemcpyDefault, st));
  // the slab also holds the host path's scratch words (voided-step flags, sequence stamp, CTA check-in counter):
  // they belong to the calls in flight, not to the env's state
  RT_CUDA(cudaMemsetAsync(h->d_flag_or, 0, 32, st));
  if (!is_pinned_or_device(src)) RT_CUDA(cudaStreamSynchronize(st));
  h->last_maps = nullptr;  // no tensor holds this state's observation
  return RT_OK;
}

int rt_env_buffer(rt_env* h, int which, void** dev_ptr, size_t* bytes) {
  if (!h || which < 0 || which >= RT_BUF__COUNT || !h->buf_ptr[which]) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  if ((which == RT_BUF_INTENSITY || which == RT_BUF_VISIT) && !h->dev.live_dense) {
    // sparse mode keeps visits / intensities in the records only: bring the dense views up to date (synchronises)
    RT_CUDA(cudaDeviceSynchronize());
    RT_CUDA(cudaMemset(h->dev.visit, 0, h->buf_bytes[RT_BUF_VISIT]));  // resets leave the dense views alone in sparse mode
    RT_CUDA(cudaMemset(h->dev.inten, 0, h->buf_bytes[RT_BUF_INTENSITY]));
    k_densify<<<(unsigned)((h->dev.E + 127) / 128), 128>>>(h->dev);
    h->launches += 1;
    RT_CUDA_LAUNCH_CHECK();
    RT_CUDA(cudaDeviceSynchronize());
  }
  if (dev_ptr) *dev_ptr = h->buf_ptr[which];
  if (bytes) *bytes = h->buf_bytes[which];
  return RT_OK;
}

int rt_env_errors(rt_env* h, int32_t* flags_host, void* stream) {
  if (!h || !flags_host) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  RT_CUDA(cudaMemsetAsync(h->d_flag_or, 0, 4, st));
  int blocks = (h->dev.E + 255) / 256;
  if (blocks > h->n_sm * 8) blocks = h->n_sm * 8;
  k_or_flags<<<blocks, 256, 0, st>>>(h->dev.flags, h->dev.E, h->d_flag_or);
  h->launches += 1;
  RT_CUDA_LAUNCH_CHECK();
  RT_CUDA(cudaMemcpyAsync(flags_host, h->d_flag_or, 4, cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaStreamSynchronize(st));
  return RT_OK;
}

int64_t rt_env_launch_count(const rt_env* h) { return h ? h->launches : 0; }

}  // extern "C"
