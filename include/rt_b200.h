/*
 * rt_b200.h — C ABI of librt_b200.so: the B200-native (sm_100a) batched
 * radiation-search environment step and observation pipeline.
 *
 * This is the drop-in boundary for RAD-TEAM's data-parallel hot path.  The
 * reference has no FFI (it is pure Python), so every entry point below names
 * the reference *interface* it stands behind (paths relative to the reference
 * repo root).  The Python facade in rad_team_b200/ binds these with ctypes and
 * keeps the reference's class and method names; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns an rt_status (0 = ok) and never throws;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - pointers suffixed _dev are DEVICE pointers owned by the caller (PyTorch
 *     tensors), _host are HOST pointers; the handle owns all persistent state;
 *   - nothing here synchronises the device unless its comment says so;
 *   - a handle belongs to the CUDA device that was current in rt_*_create; its
 *     other entry points run on that device whatever the calling thread has
 *     selected since (and restore the caller's selection).  A handle is not
 *     thread-safe: one caller, one stream at a time;
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point returns RT_ERR_CUDA.
 *
 * Layouts (row-major, E local envs, A agents, G grid side, C = 3 channels)
 *   actions  int32 [E][A]        1 UP(+y) 2 RIGHT(+x) 3 LEFT(-x) 4 DOWN(-y)
 *   obs_xy   int32 [E][A][2]     (x, y) grid indices
 *   obs_maps float [E][3][G][G]  channel 0 location, 1 intensity, 2 visits;
 *                                map[row = y][col = x]
 *   reward   int32 [E][A]        in {-1, 0, +1}
 *   done     uint8 [E][A]
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 1
#define RT_MAX_AGENTS 8
#define RT_MAX_POLYS 5
#define RT_EDGES_PER_POLY 4
#define RT_MAX_EDGES (RT_MAX_POLYS * RT_EDGES_PER_POLY) /* 20 edges = 320 B / env */
#define RT_MAP_CHANNELS 3

typedef enum rt_status {
  RT_OK = 0,
  RT_ERR_BAD_ARG = 1,    /* null pointer, size out of range, bad config      */
  RT_ERR_CUDA = 2,       /* a CUDA runtime call failed / no device           */
  RT_ERR_NOMEM = 3,      /* device or host allocation failed                 */
  RT_ERR_BAD_ACTION = 4, /* action outside 1..4 (reference: ValueError)      */
  RT_ERR_ASSERT = 5,     /* a reference `assert` would have fired            */
  RT_ERR_KEY = 6         /* estimator key missing (reference: ValueError)    */
} rt_status;

/* Sticky per-env error bits, OR-reduced by rt_env_errors(). */
#define RT_FLAG_BAD_ACTION 0x01 /* step skipped for that env (no state change)     */
#define RT_FLAG_LOG_FULL   0x02 /* more than max_steps steps since reset            */
#define RT_FLAG_LAMBDA     0x04 /* expected count outside [0, 2^30]                 */
#define RT_FLAG_DRAWS      0x08 /* Poisson sampler ran out of (injected) uniforms   */
#define RT_FLAG_NORMALIZE  0x10 /* a Normalizer assertion would have fired          */

typedef struct rt_env rt_env;           /* batched environment (opaque)          */
typedef struct rt_welford rt_welford;   /* per-stream Welford standardisers      */
typedef struct rt_moments rt_moments;   /* batch running moments (Chan-merged)   */
typedef struct rt_estimator rt_estimator; /* per-cell median resampling estimator */
typedef struct rt_comm rt_comm;         /* peer mailboxes of the ranks of one box */

#define RT_COMM_MAX_RANKS 16    /* ranks (GPUs) of one box                       */
#define RT_COMM_HANDLE_BYTES 64 /* exported mailbox handle (a CUDA IPC handle)   */

/* One POD config across the ABI.  Stands in for the constructor kwargs of
 * SimpleGrid(start, terminal, size=10, render_mode=None)
 * (src/envs/simple_gridworld.py:87-93) plus the scenario parameters the
 * north-star adds.  Zero-initialise, then fill. */
typedef struct rt_env_cfg {
  int32_t n_envs;       /* E: envs owned by this handle (this GPU's shard)        */
  int32_t n_agents;     /* A: 1..RT_MAX_AGENTS                                   */
  int32_t grid;         /* G: cells per side (reference `size`, default 10), 2..255 */
  int32_t max_steps;    /* T: steps per episode; reading log holds T*A entries;
                           T*A must be in [2, 65534]                             */
  int32_t poly_min;     /* obstructions per generated scenario: U{min..max}      */
  int32_t poly_max;     /*   0 <= min <= max <= RT_MAX_POLYS                     */
  int32_t rect_min;     /* generated rectangles: side ~ U{rect_min..rect_max}    */
  int32_t rect_max;     /*   cells                                               */
  int32_t lognorm_step; /* BensLogNormalizer step_size (normalize.py:80), 2      */
  int32_t fixed_scenario; /* 1: reset keeps the scenario set by rt_env_set_scenario
                             (SimpleGrid semantics: fixed start / terminal)      */
  int64_t env_id_base;  /* global id of local env 0 (multi-GPU sharding); RNG
                           streams are functions of the GLOBAL env id            */
  uint64_t seed;
  double cell_pitch_cm; /* distance between neighbouring cell centres            */
  double min_dist_cm;   /* inverse-square clamp: d := max(d, min_dist_cm)        */
  double transmission;  /* source term is multiplied by this once per obstruction
                           on the line of sight (0 => background only)           */
  double src_lo, src_hi; /* generated source strength  ~ U[lo, hi)  cps*cm^2     */
  double bkg_lo, bkg_hi; /* generated background rate  ~ U[lo, hi)  cps          */
} rt_env_cfg;

/* Names for rt_env_buffer(): raw device views of the handle's state, for
 * checkpoint / deterministic replay (get_state / set_state) and parity tests.
 * Shapes are in the comments; all arrays are dense row-major.  The arrays are
 * mutually consistent views of one state (e.g. VREC repeats the agents' cells
 * and the visited cells of VISIT / INTENSITY for the rasteriser): restore a
 * snapshot by copying ALL of them back, do not edit one in isolation. */
typedef enum rt_buf {
  RT_BUF_AGENT_XY = 0,   /* int32  [E][A][2]                               */
  RT_BUF_PREV_DIST = 1,  /* int32  [E][A]   agent_previous_dist (L1)       */
  RT_BUF_START_XY = 2,   /* int32  [E][A][2]                               */
  RT_BUF_SRC_XY = 3,     /* int32  [E][2]   source cell == reference `terminal` */
  RT_BUF_SRC_INT = 4,    /* double [E]                                     */
  RT_BUF_BKG = 5,        /* double [E]                                     */
  RT_BUF_N_POLY = 6,     /* int32  [E]                                     */
  RT_BUF_EDGES = 7,      /* float  [E][20][4]  (x0,y0,x1,y1) in cell units */
  RT_BUF_TICK = 8,       /* int32  [E]      steps since reset              */
  RT_BUF_EPISODE = 9,    /* int32  [E]      resets so far - 1              */
  RT_BUF_VISIT = 10,     /* uint16 [E][G*G] visit counts (dense view, see INTENSITY) */
  RT_BUF_INTENSITY = 11, /* float  [E][G*G] normalised intensity map.  With the default (sparse)
                            rasteriser the step keeps visits and intensities in CELLREC / VREC only
                            and the dense VISIT / INTENSITY views are brought up to date by each
                            rt_env_buffer(VISIT or INTENSITY) call (which synchronises the device);
                            with the dense rasterisers they are live */
  RT_BUF_CELLREC = 12,   /* uint16 [E][G*G][4] per-cell record of the step kernel: episode tag (low 16
                            bits of EPISODE when the record was written; a record whose tag differs from
                            the env's current episode is EMPTY — resets do not clear this array), first
                            bucket of the cell in LOG + 1 (0 = none), record index in VREC + 1
                            (0 = not visited this episode), unused                          */
  RT_BUF_LOG = 13,       /* uint32 [T*A][E][8] reading buckets, time-major: 7 Poisson counts of ONE cell in
                            ascending order (the first min(7, rest) are valid) + a link word: low 16 bits =
                            next bucket of the cell + 1, high 16 bits (first bucket of a cell only) = number
                            of readings of the cell; all buckets of a cell are full except the last and
                            ordered by value; a bucket's id is tick*A + agent of the visit that opened it */
  RT_BUF__RESERVED14 = 14,
  RT_BUF_WELFORD = 15,   /* double [E][6]   mean, M2, var, std, zmax, zmin */
  RT_BUF_WELFORD_N = 16, /* int32  [E]      count                          */
  RT_BUF_EST_MAXMIN = 17,/* double [E][2]   estimator _max, _min           */
  RT_BUF_FLAGS = 18,     /* int32  [E]      sticky RT_FLAG_* bits          */
  RT_BUF_LAST_COUNT = 19,/* int32  [E][A]   Poisson counts of the last step */
  RT_BUF_LAST_LOS = 20,  /* uint8  [E][A]   obstructions on the line of sight */
  RT_BUF_LAST_LAMBDA = 21,/* double [E][A]  expected counts of the last step */
  RT_BUF_LAST_EST = 22,  /* double [E][A]   median estimate at the agent cell */
  RT_BUF_LAST_Z = 23,    /* double [E][A]   standardised estimate            */
  RT_BUF_LAST_VALUE = 24,/* double [E][A]   normalised value written to the map */
  RT_BUF_VISIT_LUT = 25, /* float  [T*A+1]  BensLogNormalizer table          */
  RT_BUF_EP_RETURN = 26, /* int32  [E][A]   sum of rewards since reset       */
  RT_BUF_EP_DONE = 27,   /* uint8  [E][A]   1 once the agent has reached the source this episode */
  RT_BUF__RESERVED28 = 28,
  RT_BUF_VREC = 29,      /* uint32 [E][R][2] compact list of visited cells read by the sparse rasteriser:
                            [0] = {n_visited, 0}; [1..2] = agent cells, 8 x uint16 (0xFFFF = none);
                            [2+k] = {cell | visits << 16, fp32 bits of the intensity value};
                            R = (3 + min(T*A, G*G) + 1) & ~1                                  */
  RT_BUF__COUNT = 30
} rt_buf;

int rt_abi_version(void);
const char* rt_status_string(int status);

/* ---- environment: src/envs/simple_gridworld.py --------------------------- */

/* SimpleGrid.__init__ (simple_gridworld.py:87-112), batched.  Allocates all
 * persistent device state on the current CUDA device.  Synchronous. */
int rt_env_create(const rt_env_cfg* cfg, rt_env** out);
/* SimpleGrid.close (simple_gridworld.py:174-178). */
int rt_env_destroy(rt_env* env);

/* Fix the scenario instead of generating it: the batched form of the
 * reference's `start` / `terminal` constructor arguments
 * (simple_gridworld.py:96-98).  HOST pointers, any may be NULL to keep the
 * current value.  src_xy [E][2], src_int [E], bkg [E], start_xy [E][A][2],
 * n_poly [E], edges [E][20][4].  Synchronous (setup path). */
int rt_env_set_scenario(rt_env* env, const int32_t* src_xy_host, const double* src_int_host,
                        const double* bkg_host, const int32_t* start_xy_host,
                        const int32_t* n_poly_host, const float* edges_host);

/* SimpleGrid.reset (simple_gridworld.py:114-129), batched and masked:
 * envs with mask[e] != 0 (all when mask is NULL) start a new episode — a new
 * generated scenario unless cfg.fixed_scenario, agents back at their start
 * cells, agent_previous_dist recomputed, maps / estimator / Welford cleared.
 * Writes obs_xy [E][A][2] and, when not NULL, rasterises obs_maps. */
int rt_env_reset(rt_env* env, const uint8_t* mask_dev, int32_t* obs_xy_dev, float* obs_maps_dev,
                 void* stream);

/* SimpleGrid.step (simple_gridworld.py:131-167), batched, followed by the
 * measurement, the observation pipeline of src/math_utils (estimator.py:31-70
 * -> standardize.py:46-93 -> normalize.py:25-55, and normalize.py:79-115 for
 * visits) and the rasterisation of obs_maps (skipped when obs_maps_dev is NULL).
 * `truncated` is constantly False in the reference (:167) and is not returned. */
int rt_env_step(rt_env* env, const int32_t* actions_dev, int32_t* obs_xy_dev, float* obs_maps_dev,
                int32_t* reward_dev, uint8_t* done_dev, void* stream);

/* The same two calls with HOST buffers for everything the reference returns
 * to Python (obs, reward, done); the maps stay on the device (they are the
 * CNN's input tensor).  Pinned caller buffers are used directly, pageable ones
 * go through pinned staging owned by the handle.  The step kernel reads pinned
 * actions in place over PCIe.  The results return
 *   - results of at most 256 KB (small batches): stored by the step kernel
 *     itself into (mapped) host memory, its last CTA publishing a sequence
 *     word — no copy-engine request at all;
 *   - larger results: by the copy engine on an internal stream while the
 *     rasteriser runs — ONE request when obs_xy_host, reward_host and
 *     done_host are one pinned block laid out obs_xy | reward | done
 *     (reward_host == obs_xy_host + 8 E A bytes, done_host == + 12 E A bytes),
 *     three otherwise — followed by the sequence word;
 * and the call returns as soon as that word has landed (it is polled in
 * pinned memory, no driver call on the wait): the rasterisation of
 * obs_maps_dev may still be in flight on `stream` (it is ordered before
 * anything enqueued there later — the consumer network, the next step — so the
 * host's work between two steps hides under it).  rt_env_reset_host likewise
 * returns when the start positions have landed (with a mask: when `stream` has
 * drained).
 * rt_env_step_host returns RT_ERR_BAD_ACTION when an action was
 * outside 1..4: the steps of exactly those envs were voided (no state change,
 * as in the reference, which raises before touching state); every other env
 * advanced and all outputs are valid. */
int rt_env_reset_host(rt_env* env, const uint8_t* mask_host, int32_t* obs_xy_host,
                      float* obs_maps_dev, void* stream);
int rt_env_step_host(rt_env* env, const int32_t* actions_host, int32_t* obs_xy_host,
                     int32_t* reward_host, uint8_t* done_host, float* obs_maps_dev, void* stream);

/* SimpleGrid._get_obs (simple_gridworld.py:188-190) for the map observation:
 * rasterise the current state into obs_maps [E][3][G][G]. */
int rt_env_rasterize(rt_env* env, float* obs_maps_dev, void* stream);

/* Opt-in for loops that hand the SAME obs_maps tensor to every step (inference,
 * evaluation): when enabled and obs_maps is the tensor this env rasterised
 * into last time, rt_env_step only rewrites the <= 2A cells per env that the
 * step changed instead of the whole tensor.  The caller must not have modified
 * the tensor in between.  A different pointer, a reset, or a step without maps
 * falls back to a full rasterisation.  Off by default. */
int rt_env_set_incremental(rt_env* env, int32_t enable);

/* Test hook: replace the counter-based RNG of the Poisson sampler by injected
 * uniforms u [E][A][k_per_agent] (device, doubles in [0,1)); NULL switches
 * back to Philox.  The pointer must stay valid until replaced. */
int rt_env_inject_uniforms(rt_env* env, const double* u_dev, int32_t k_per_agent);

/* Rollout sink: when set, every step also stores each agent's Poisson count as
 * fp32 into readings_dev [T][E][A] at row `tick` (the env's step index in its
 * episode) — the buffer rt_moments_update() consumes at the end of a rollout.
 * NULL detaches.  The pointer must stay valid until replaced. */
int rt_env_set_rollout(rt_env* env, float* readings_dev);

/* Episode statistics of the current episodes, reduced over all (env, agent):
 * stats_dev double [3] = { sum of returns, sum of episode lengths (steps),
 * number of agents that reached the source }.  What each rank contributes,
 * next to its moments, to the once-per-rollout all-gather. */
int rt_env_episode_stats(rt_env* env, double* stats_dev, void* stream);

/* Checkpoint / deterministic replay for hosts that do not want to walk rt_buf:
 * the handle's whole persistent state (every rt_buf array, RNG positions are
 * functions of it) as ONE opaque blob.  rt_env_state_bytes gives its size;
 * rt_env_get_state copies it to `dst` and rt_env_set_state restores it from
 * `src` (host or device memory, cudaMemcpyDefault) in stream order; set_state
 * checks that the blob came from a handle of the same shape (else
 * RT_ERR_BAD_ARG) and SYNCHRONISES `stream` when `src` is pageable host memory.
 * The reference env has no such call ("expose a
 * state_dict-style export for deterministic replay"). */
int rt_env_state_bytes(const rt_env* env, size_t* bytes);
int rt_env_get_state(rt_env* env, void* dst, void* stream);
int rt_env_set_state(rt_env* env, const void* src, void* stream);

/* Raw device view of one state array (see rt_buf). */
int rt_env_buffer(rt_env* env, int which, void** dev_ptr, size_t* bytes);
/* OR of all envs' sticky flags -> *flags_host.  Synchronises `stream`. */
int rt_env_errors(rt_env* env, int32_t* flags_host, void* stream);
/* Number of kernels this handle has launched so far (bench bookkeeping). */
int64_t rt_env_launch_count(const rt_env* env);

/* ---- observation-tensor memory ---------------------------------------------
 * north_star: "writes directly into the CNN input tensor" — the tensor is the
 * caller's.  These two calls let the caller place it in COMPRESSIBLE device
 * memory (driver virtual-memory API, CU_MEM_ALLOCATION_COMP_GENERIC): the maps
 * are mostly zeros, and Blackwell's L2 compresses such lines on their way to
 * HBM, which shortens both the rasteriser's store stream and the consumer's
 * read pass.  The same holds for rollout buffers of Poisson readings (small
 * integers stored as float32, rt_env_set_rollout / rt_moments_update).  Any device memory works with rt_env_step; this is an option, not
 * a requirement.  The block lives on the current device; contents are
 * uninitialised.  *compressible_out (may be NULL) = 1 when compressible backing
 * was granted, 0 when the device (or RT_B200_OBS_COMPRESS=0) gave plain memory.
 * rt_obs_free synchronises that device, then unmaps and releases the block;
 * a pointer that did not come from rt_obs_alloc is RT_ERR_BAD_ARG. */
int rt_obs_alloc(size_t bytes, void** dev_ptr, int32_t* compressible_out);
int rt_obs_free(void* dev_ptr);

/* ---- WelfordsOnlineStandardization: src/math_utils/standardize.py:7-107 --- */

/* n independent streams, each with the reference's exact sequential state. */
int rt_welford_create(int64_t n_streams, rt_welford** out);
int rt_welford_destroy(rt_welford* w);
/* .reset() (standardize.py:36-44) for every stream. */
int rt_welford_reset(rt_welford* w, void* stream);
/* .update(reading) (standardize.py:46-80): one reading per stream.  A negative
 * reading leaves its stream untouched and sets err_dev[i] = RT_ERR_ASSERT. */
int rt_welford_update(rt_welford* w, const double* reading_dev, int32_t* err_dev, void* stream);
/* .standardize(reading) (standardize.py:82-93). */
int rt_welford_standardize(rt_welford* w, const double* reading_dev, double* z_dev,
                           int32_t* err_dev, void* stream);
/* State view: double [n][6] = mean, M2, sample_variance, std, _max, _min and
 * int32 [n] count. */
int rt_welford_buffers(rt_welford* w, void** state_dev, void** count_dev);

/* ---- batch running moments (cross-env normaliser, merged over NCCL) ------- */

int rt_moments_create(rt_moments** out);
int rt_moments_destroy(rt_moments* m);
int rt_moments_reset(rt_moments* m, void* stream);
/* Fold n fp32 readings into the running (count, mean, M2): warp-shuffle /
 * block / grid Chan merges in a fixed order (deterministic for a given n and
 * alignment).  x_dev: any contiguous float buffer (4-byte aligned; up to three
 * readings in front of the first 16-byte boundary are folded singly). */
int rt_moments_update(rt_moments* m, const float* x_dev, int64_t n, void* stream);
/* out = (x - mean) / max(sqrt(M2/(count-1)), 1)  (standardize.py:74,93). */
int rt_moments_standardize(rt_moments* m, const float* x_dev, float* out_dev, int64_t n,
                           void* stream);
/* Device view of the state: double [3] = count, mean, M2 (what each rank
 * contributes to the all-gather). */
int rt_moments_state(rt_moments* m, void** state_dev);
/* Replace the state by the Chan merge, in rank order 0..n_ranks-1, of the
 * gathered triples: rank r's {count, mean, M2} starts at
 * gathered_dev[r * stride_doubles] (device; stride >= 3). */
int rt_moments_merge(rt_moments* m, const double* gathered_dev, int32_t n_ranks,
                     int32_t stride_doubles, void* stream);

/* ---- multi-GPU: the once-per-rollout exchange -------
 * The reference is single-process (simple_gridworld.py holds no cross-instance
 * state, :58-65), so these entry points stand behind nothing in it: they are
 * what lets envs shard over the GPUs of one box, one process per GPU.  `step`
 * has no cross-GPU traffic; once per rollout every rank contributes six
 * doubles — the moments of ITS readings of this rollout (count, mean, M2 of the
 * rt_moments `local`) and its episode statistics (rt_env_episode_stats) — and
 * every rank merges the rows in rank order: `global` <- chan(`global`,
 * chan-merge over ranks of the locals), episode statistics summed, `local`
 * cleared for the next rollout.  Same IEEE operations in the same order on every
 * rank => bit-identical results.  `local` and `global` must be distinct handles
 * (a running state that was also the per-rollout accumulator would be counted
 * once per rank at the next merge).
 *
 * Two transports, same result:
 *   (1) rt_stats_allgather_merge — ONE kernel per rank over peer-mapped
 *       mailboxes (NVLink / NVSwitch): stores into every peer's mailbox, flag,
 *       wait, merge.  No library collective, no host round trip, CUDA-graph
 *       capturable.  Collective semantics: every rank must call it the same
 *       number of times; a peer that does not arrive within the timeout
 *       (RT_B200_COMM_TIMEOUT_MS, default 30000) sets a sticky error
 *       (rt_comm_errors) instead of hanging the device.
 *   (2) rt_stats_pack -> the caller's all-gather of 6 doubles per rank (NCCL
 *       through torch.distributed, MPI, ...) -> rt_stats_merge_gathered. */

/* Step 1 on every rank (current device = the rank's GPU): allocate this rank's
 * mailbox and export its handle (RT_COMM_HANDLE_BYTES bytes, host).  world == 1
 * needs no step 2. */
int rt_comm_create(int32_t rank, int32_t world, rt_comm** out, void* handle_out_host);
/* Step 2: the handles of ALL ranks in rank order (world x RT_COMM_HANDLE_BYTES,
 * host), exchanged by the caller by any means; maps every peer's mailbox. */
int rt_comm_connect(rt_comm* comm, const void* all_handles_host);
/* Steps 1 + 2 for hosts without a communication library: the ranks of one box
 * meet in the POSIX shared-memory object `name` ("/something", unique per job;
 * removed again once every rank has connected).  Blocks until all ranks arrive. */
int rt_comm_create_shm(const char* name, int32_t rank, int32_t world, rt_comm** out);
/* Orderly shutdown: every rank calls rt_comm_disconnect (unmaps the peers'
 * mailboxes; no further exchange), the caller synchronises the ranks by its own
 * means, then every rank calls rt_comm_destroy (frees its own mailbox) — an
 * exported mailbox must outlive its mappings in the other processes.
 * rt_comm_destroy alone does both for this rank. */
int rt_comm_disconnect(rt_comm* comm);
int rt_comm_destroy(rt_comm* comm);
/* Sticky error word -> *err_host (1 = a peer timed out).  Synchronises `stream`. */
int rt_comm_errors(rt_comm* comm, int32_t* err_host, void* stream);

/* Transport (1).  ep_stats_dev double [3] (may be NULL = zeros) is this rank's
 * rt_env_episode_stats output; ep_merged_dev double [3] (may be NULL) receives
 * the sums over ranks; gathered_out_dev double [world][6] (may be NULL) receives
 * the gathered rows for inspection. */
int rt_stats_allgather_merge(rt_comm* comm, rt_moments* local, rt_moments* global,
                             const double* ep_stats_dev, double* ep_merged_dev,
                             double* gathered_out_dev, void* stream);
/* Transport (2): out6_dev double [6] = {count, mean, M2, episode stats}; then,
 * after the caller's all-gather into gathered_dev double [n_ranks][6] (row r =
 * rank r), the same merge as above. */
int rt_stats_pack(rt_moments* local, const double* ep_stats_dev, double* out6_dev, void* stream);
int rt_stats_merge_gathered(rt_moments* local, rt_moments* global, const double* gathered_dev,
                            int32_t n_ranks, double* ep_merged_dev, void* stream);

/* ---- Normalizer / BensLogNormalizer: src/math_utils/normalize.py ---------- */

/* Normalizer.normalize (normalize.py:25-55), elementwise.  min_dev may be NULL
 * (reference `min=None`); err_dev[i] = RT_ERR_ASSERT where the reference
 * asserts, and out[i] is then NaN. */
int rt_normalize(const double* current_dev, const double* max_dev, const double* min_dev,
                 double* out_dev, int32_t* err_dev, int64_t n, void* stream);
/* BensLogNormalizer.normalize_incremental_logscale (normalize.py:79-115). */
int rt_lognormalize(const double* current_dev, double max, int32_t step_size, double* out_dev,
                    int32_t* err_dev, int64_t n, void* stream);

/* ---- IntensityResamplingEstimator: src/math_utils/estimator.py:6-109 ------ */

/* n independent estimators over `n_keys` key slots each, `capacity` readings
 * per estimator. */
int rt_estimator_create(int64_t n, int32_t n_keys, int32_t capacity, rt_estimator** out);
int rt_estimator_destroy(rt_estimator* est);
int rt_estimator_reset(rt_estimator* est, void* stream);
/* .update(key, value) (estimator.py:31-48) once per estimator; writes the new
 * median estimate of that key to est_dev[i] (may be NULL).  err_dev[i] =
 * RT_ERR_BAD_ARG for a key outside [0, n_keys), RT_ERR_NOMEM when the
 * estimator's `capacity` readings are used up (est_dev[i] is then NaN). */
int rt_estimator_update(rt_estimator* est, const int32_t* key_dev, const double* value_dev,
                        double* est_dev, int32_t* err_dev, void* stream);
/* .get_estimate(key) (estimator.py:61-70); RT_ERR_KEY in err_dev for a key
 * never updated. */
int rt_estimator_estimate(rt_estimator* est, const int32_t* key_dev, double* est_dev,
                          int32_t* err_dev, void* stream);
/* .get_buffer(key) (estimator.py:50-59): readings of key_dev[i] in insertion
 * order into out_dev [n][max_out]; n_out_dev[i] = buffer length. */
int rt_estimator_buffer(rt_estimator* est, const int32_t* key_dev, double* out_dev,
                        int32_t* n_out_dev, int32_t max_out, int32_t* err_dev, void* stream);
/* double [n][2] = _max, _min (estimator.py:72-88). */
int rt_estimator_maxmin(rt_estimator* est, void** maxmin_dev);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
