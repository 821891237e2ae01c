// rt_comm.cu — the ONE place GPUs talk: the once-per-rollout exchange of reading moments and episode
// statistics between the ranks of one box.
//
//   rt_comm_*                  peer mailboxes: every rank owns a small device buffer, exported through CUDA IPC
//                              and mapped by every other rank of the box (NVLink / NVSwitch peer access)
//   rt_stats_allgather_merge   ONE kernel launch per rank: pack this rank's six doubles, store them into its
//                              slot of every peer's mailbox with a sequence flag, wait until all slots of the own
//                              mailbox carry the flag, Chan-merge the rollout's per-rank moments in rank order,
//                              fold the result into the persistent global moments, sum the episode statistics and
//                              clear the per-rollout accumulator.  No host round trip, no library collective;
//                              CUDA-graph capturable (the sequence number lives in device memory).
//   rt_stats_pack / rt_stats_merge_gathered   the same exchange around a library all-gather (NCCL through the
//                              caller's communication layer): pack -> [caller's all-gather] -> merge.
//
// Every rank performs the same IEEE operations in the same order on the same gathered rows, so the merged
// moments are bit-identical everywhere.  Moments are never summed as (sum x, sum x^2).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

#include "rt_host.h"
#include "rt_moments.cuh"

using namespace rt;

namespace {

constexpr int kSlotDoubles = 8;  // 6 payload doubles + sequence word + pad = 64 B, one slot per (parity, rank)

struct CommDev {
  int rank, world;
  double* mailbox[RT_COMM_MAX_RANKS];  // mailbox of rank r as mapped in THIS process: [2][world][kSlotDoubles]
  unsigned long long* seq;             // exchanges completed so far (device)
  int* err;                            // sticky: 1 = a peer did not arrive within the timeout
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// rows[r] = rank r's {count, mean, M2, sum_return, sum_len, n_done}; lane 0 merges them in rank order
__device__ __forceinline__ void merge_rows(const double (&row)[6], int world, int lane, double* __restrict__ local,
                                           double* __restrict__ global, double* __restrict__ ep_out,
                                           double* __restrict__ gathered_out) {
  Mom delta = {0.0, 0.0, 0.0};
  double ep[3] = {0.0, 0.0, 0.0};
  for (int r = 0; r < world; ++r) {
    double v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = __shfl_sync(0xffffffffu, row[i], r);
    if (lane == 0) {
      delta = chan(delta, Mom{v[0], v[1], v[2]});
      ep[0] += v[3];  // integer-valued doubles well below 2^53: exact in any order, summed in rank order anyway
      ep[1] += v[4];
      ep[2] += v[5];
      if (gathered_out)
        for (int i = 0; i < 6; ++i) gathered_out[6 * r + i] = v[i];
    }
  }
  if (lane == 0) {
    const Mom g = chan(Mom{global[0], global[1], global[2]}, delta);
    global[0] = g.n;
    global[1] = g.mean;
    global[2] = g.m2;
    if (ep_out) {
      ep_out[0] = ep[0];
      ep_out[1] = ep[1];
      ep_out[2] = ep[2];
    }
    local[0] = local[1] = local[2] = 0.0;  // the per-rollout accumulator starts over
  }
}

__global__ void __launch_bounds__(32)
k_stats_exchange(const CommDev c, double* __restrict__ local, double* __restrict__ global,
                 const double* __restrict__ ep, double* __restrict__ ep_out, double* __restrict__ gathered_out) {
  const int lane = threadIdx.x;
  const unsigned long long seq = *c.seq + 1ull;
  const int par = (int)(seq & 1ull);  // two slots per rank: exchange k + 1 never overwrites what a peer still reads of k
  double mine[6];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    mine[i] = local[i];
    mine[3 + i] = ep ? ep[i] : 0.0;
  }
  // 1. lane r writes this rank's row into slot [par][rank] of rank r's mailbox (its own included), then the flag
  if (lane < c.world) {
    double* slot = c.mailbox[lane] + ((size_t)par * c.world + c.rank) * kSlotDoubles;
#pragma unroll
    for (int i = 0; i < 6; ++i) st_relaxed_sys(slot + i, mine[i]);
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(slot + 6), seq);
  }
  // 2. lane r waits for rank r's row in the own mailbox and reads it
  double row[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (lane < c.world) {
    const double* slot = c.mailbox[c.rank] + ((size_t)par * c.world + lane) * kSlotDoubles;
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(slot + 6);
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    while (ld_acquire_sys(flag) != seq) {
      if (globaltimer_ns() - t0 > c.timeout_ns) {  // a kernel must never spin forever: flag the failure and go on
        ok = false;
        atomicOr(c.err, 1);
        break;
      }
      __nanosleep(200);
    }
    if (ok) {
#pragma unroll
      for (int i = 0; i < 6; ++i) row[i] = ld_relaxed_sys(slot + i);
    }
  }
  __syncwarp();
  // 3. rank-ordered merge, identical on every rank
  merge_rows(row, c.world, lane, local, global, ep_out, gathered_out);
  if (lane == 0) *c.seq = seq;
}

__global__ void __launch_bounds__(32)
k_stats_pack(const double* __restrict__ local, const double* __restrict__ ep, double* __restrict__ out6) {
  const int i = threadIdx.x;
  if (i < 3) out6[i] = local[i];
  else if (i < 6) out6[i] = ep ? ep[i - 3] : 0.0;
}

__global__ void __launch_bounds__(32)
k_stats_merge_gathered(const double* __restrict__ gathered, int world, double* __restrict__ local,
                       double* __restrict__ global, double* __restrict__ ep_out) {
  const int lane = threadIdx.x;
  double row[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (lane < world)
    for (int i = 0; i < 6; ++i) row[i] = gathered[6 * lane + i];
  merge_rows(row, world, lane, local, global, ep_out, nullptr);
}

// Rendezvous block in POSIX shared memory (rt_comm_create_shm): ranks of one box, no communication library.
struct ShmBlock {
  std::atomic<int> ready[RT_COMM_MAX_RANKS];
  std::atomic<int> connected;
  unsigned char handle[RT_COMM_MAX_RANKS][RT_COMM_HANDLE_BYTES];
};

}  // namespace

struct rt_comm {
  int device = 0;
  int rank = 0, world = 1;
  double* mailbox = nullptr;                  // own mailbox (cudaMalloc): [2][world][kSlotDoubles] + seq + err
  double* peer[RT_COMM_MAX_RANKS] = {};       // mapped peers (peer[rank] == mailbox)
  bool opened[RT_COMM_MAX_RANKS] = {};
  bool connected = false;
  unsigned long long timeout_ns = 30000000000ull;  // 30 s: ranks may legitimately be seconds apart (logging, checkpoints)
  size_t mailbox_bytes = 0;
};

extern "C" {

int rt_comm_create(int32_t rank, int32_t world, rt_comm** out, void* handle_out) {
  if (!out || world < 1 || world > RT_COMM_MAX_RANKS || rank < 0 || rank >= world) return RT_ERR_BAD_ARG;
  *out = nullptr;
  static_assert(sizeof(cudaIpcMemHandle_t) <= RT_COMM_HANDLE_BYTES, "IPC handle does not fit the ABI's handle size");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    rt_note_cuda_error(cudaGetLastError(), "cudaGetDeviceCount");
    return RT_ERR_CUDA;
  }
  rt_comm* c = new (std::nothrow) rt_comm();
  if (!c) return RT_ERR_NOMEM;
  c->rank = rank;
  c->world = world;
  RT_CUDA_OR(cudaGetDevice(&c->device), { delete c; });
  if (const char* ev = std::getenv("RT_B200_COMM_TIMEOUT_MS")) {
    const long long ms = std::atoll(ev);
    if (ms > 0) c->timeout_ns = (unsigned long long)ms * 1000000ull;
  }
  // slots, then the sequence counter and the error word in the tail (64 B)
  c->mailbox_bytes = (size_t)2 * world * kSlotDoubles * sizeof(double) + 64;
  if (cudaMalloc((void**)&c->mailbox, c->mailbox_bytes) != cudaSuccess) {
    rt_note_cuda_error(cudaGetLastError(), "cudaMalloc(mailbox)");
    delete c;
    return RT_ERR_NOMEM;
  }
  RT_CUDA_OR(cudaMemset(c->mailbox, 0, c->mailbox_bytes), { rt_comm_destroy(c); });
  RT_CUDA_OR(cudaDeviceSynchronize(), { rt_comm_destroy(c); });
  c->peer[rank] = c->mailbox;
  if (handle_out) {
    std::memset(handle_out, 0, RT_COMM_HANDLE_BYTES);
    if (world > 1) {
      cudaIpcMemHandle_t h;
      RT_CUDA_OR(cudaIpcGetMemHandle(&h, c->mailbox), { rt_comm_destroy(c); });
      std::memcpy(handle_out, &h, sizeof h);
    }
  }
  c->connected = world == 1;
  *out = c;
  return RT_OK;
}

int rt_comm_connect(rt_comm* c, const void* all_handles_host) {
  if (!c) return RT_ERR_BAD_ARG;
  if (c->connected) return RT_OK;
  if (!all_handles_host) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(c->device);
  const unsigned char* hs = static_cast<const unsigned char*>(all_handles_host);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hs + (size_t)r * RT_COMM_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    RT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = static_cast<double*>(p);
    c->opened[r] = true;
  }
  c->connected = true;
  return RT_OK;
}

int rt_comm_create_shm(const char* name, int32_t rank, int32_t world, rt_comm** out) {
  if (!name || !out || name[0] != '/') return RT_ERR_BAD_ARG;
  unsigned char mine[RT_COMM_HANDLE_BYTES];
  int rc = rt_comm_create(rank, world, out, mine);
  if (rc != RT_OK || world == 1) return rc;
  rt_comm* c = *out;
  auto fail = [&](const char* what) {
    rt_note_error_text(what);
    rt_comm_destroy(c);
    *out = nullptr;
    return RT_ERR_CUDA;
  };
  const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
  if (fd < 0) return fail("rt_comm_create_shm: shm_open failed");
  if (ftruncate(fd, sizeof(ShmBlock)) != 0) {  // a fresh object is zero-filled: every flag starts at 0
    close(fd);
    return fail("rt_comm_create_shm: ftruncate failed");
  }
  void* mp = mmap(nullptr, sizeof(ShmBlock), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (mp == MAP_FAILED) return fail("rt_comm_create_shm: mmap failed");
  ShmBlock* sb = static_cast<ShmBlock*>(mp);
  std::memcpy(sb->handle[rank], mine, RT_COMM_HANDLE_BYTES);
  sb->ready[rank].store(1, std::memory_order_release);
  const auto t0 = std::chrono::steady_clock::now();
  bool all = false;
  while (!all) {
    all = true;
    for (int r = 0; r < world; ++r) all = all && sb->ready[r].load(std::memory_order_acquire) == 1;
    if (all) break;
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
      munmap(mp, sizeof(ShmBlock));
      return fail("rt_comm_create_shm: peers did not arrive within 120 s");
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  unsigned char all_handles[RT_COMM_MAX_RANKS * RT_COMM_HANDLE_BYTES];
  std::memcpy(all_handles, sb->handle, sizeof all_handles);
  rc = rt_comm_connect(c, all_handles);
  // the last rank to finish removes the name (the mapping itself lives until munmap)
  if (sb->connected.fetch_add(1, std::memory_order_acq_rel) + 1 == world) shm_unlink(name);
  munmap(mp, sizeof(ShmBlock));
  if (rc != RT_OK) {
    rt_comm_destroy(c);
    *out = nullptr;
  }
  return rc;
}

int rt_comm_disconnect(rt_comm* c) {
  if (!c) return RT_OK;
  RtDeviceGuard on_device(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r] && c->peer[r]) {
      cudaIpcCloseMemHandle(c->peer[r]);
      c->peer[r] = nullptr;
      c->opened[r] = false;
    }
  c->connected = c->world == 1;
  return RT_OK;
}

int rt_comm_destroy(rt_comm* c) {
  if (!c) return RT_OK;
  rt_comm_disconnect(c);
  RtDeviceGuard on_device(c->device);
  if (c->mailbox) cudaFree(c->mailbox);
  delete c;
  return RT_OK;
}

int rt_comm_errors(rt_comm* c, int32_t* err_host, void* stream) {
  if (!c || !err_host) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(c->device);
  const int* err = reinterpret_cast<const int*>(c->mailbox + (size_t)2 * c->world * kSlotDoubles + 1);
  RT_CUDA(cudaMemcpyAsync(err_host, err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  RT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return RT_OK;
}

int rt_stats_allgather_merge(rt_comm* c, rt_moments* local, rt_moments* global, const double* ep_stats_dev,
                             double* ep_merged_dev, double* gathered_out_dev, void* stream) {
  if (!c || !local || !global || local == global) return RT_ERR_BAD_ARG;
  if (!c->connected) {
    rt_note_error_text("rt_stats_allgather_merge: communicator not connected (rt_comm_connect)");
    return RT_ERR_BAD_ARG;
  }
  if (local->device != c->device || global->device != c->device) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(c->device);
  RtRange nvtx_range("rt_stats_allgather_merge");
  double *ls = local->state, *gs = global->state;
  CommDev d;
  d.rank = c->rank;
  d.world = c->world;
  for (int r = 0; r < RT_COMM_MAX_RANKS; ++r) d.mailbox[r] = r < c->world ? c->peer[r] : nullptr;
  double* tail = c->mailbox + (size_t)2 * c->world * kSlotDoubles;
  d.seq = reinterpret_cast<unsigned long long*>(tail);
  d.err = reinterpret_cast<int*>(tail + 1);
  d.timeout_ns = c->timeout_ns;
  k_stats_exchange<<<1, 32, 0, (cudaStream_t)stream>>>(d, ls, gs, ep_stats_dev, ep_merged_dev, gathered_out_dev);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_stats_pack(rt_moments* local, const double* ep_stats_dev, double* out6_dev, void* stream) {
  if (!local || !out6_dev) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(local->device);
  k_stats_pack<<<1, 32, 0, (cudaStream_t)stream>>>(local->state, ep_stats_dev, out6_dev);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

int rt_stats_merge_gathered(rt_moments* local, rt_moments* global, const double* gathered_dev, int32_t n_ranks,
                            double* ep_merged_dev, void* stream) {
  if (!local || !global || local == global || !gathered_dev || n_ranks < 1 || n_ranks > RT_COMM_MAX_RANKS)
    return RT_ERR_BAD_ARG;
  if (local->device != global->device) return RT_ERR_BAD_ARG;
  RtDeviceGuard on_device(local->device);
  k_stats_merge_gathered<<<1, 32, 0, (cudaStream_t)stream>>>(gathered_dev, n_ranks, local->state, global->state,
                                                            ep_merged_dev);
  RT_CUDA_LAUNCH_CHECK();
  return RT_OK;
}

}  // extern "C"
