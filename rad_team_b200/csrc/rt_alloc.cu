// rt_alloc.cu — observation-tensor memory (rt_obs_alloc / rt_obs_free).
//
// The observation maps are mostly zeros: a few dozen non-zero cells of 3 x G x G per env.  Blackwell's
// L2 compresses such lines on their way to HBM when the allocation was created compressible (driver
// virtual-memory API, CU_MEM_ALLOCATION_COMP_GENERIC); cudaMalloc memory never is.  Measured on B200
// (profiles/r02t): the rasteriser's 805 MB store stream takes 0.115 ms into compressible memory against
// 0.138 ms into cudaMalloc memory, and a consumer's read pass 0.109 ms against 0.136 ms.
//
// The driver entry points are resolved through the runtime (cudaGetDriverEntryPoint), so the library
// keeps linking against the static runtime only.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "rt_host.h"

namespace {

struct Block {
  CUmemGenericAllocationHandle handle;
  size_t size;
  int device;
};

std::mutex g_mu;
std::unordered_map<void*, Block> g_blocks;

struct Driver {
  CUresult (*getAttr)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*devGet)(CUdevice*, int) = nullptr;
  CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*propsOf)(CUmemAllocationProp*, CUmemGenericAllocationHandle) = nullptr;
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
  bool ok = false;
};

template <typename F>
bool resolve(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

const Driver& driver() {
  static const Driver d = [] {
    Driver r;
    r.ok = resolve("cuDeviceGetAttribute", r.getAttr) && resolve("cuDeviceGet", r.devGet) &&
           resolve("cuMemGetAllocationGranularity", r.granularity) && resolve("cuMemCreate", r.create) &&
           resolve("cuMemGetAllocationPropertiesFromHandle", r.propsOf) &&
           resolve("cuMemAddressReserve", r.reserve) && resolve("cuMemMap", r.map) &&
           resolve("cuMemSetAccess", r.setAccess) && resolve("cuMemUnmap", r.unmap) &&
           resolve("cuMemRelease", r.release) && resolve("cuMemAddressFree", r.addressFree);
    return r;
  }();
  return d;
}

int fail(CUresult e, const char* what, int status = RT_ERR_CUDA) {
  char msg[160];
  std::snprintf(msg, sizeof msg, "driver error %d in %s", (int)e, what);
  rt_note_error_text(msg);
  return status;
}

}  // namespace

extern "C" {

int rt_obs_alloc(size_t bytes, void** dev_ptr, int32_t* compressible_out) {
  if (!dev_ptr || bytes == 0) return RT_ERR_BAD_ARG;
  *dev_ptr = nullptr;
  if (compressible_out) *compressible_out = 0;
  int dev = 0;
  RT_CUDA(cudaFree(nullptr));  // make sure the primary context exists before talking to the driver
  RT_CUDA(cudaGetDevice(&dev));
  const Driver& D = driver();
  if (!D.ok) {
    rt_note_error_text("driver virtual-memory entry points not available");
    return RT_ERR_CUDA;
  }
  CUdevice cudev;
  CUresult e = D.devGet(&cudev, dev);
  if (e != CUDA_SUCCESS) return fail(e, "cuDeviceGet");
  int supported = 0;
  D.getAttr(&supported, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, cudev);
  bool want = supported != 0;
  if (const char* ev = std::getenv("RT_B200_OBS_COMPRESS")) want = want && std::atoi(ev) != 0;  // A/B knob

  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.allocFlags.compressionType = want ? CU_MEM_ALLOCATION_COMP_GENERIC : CU_MEM_ALLOCATION_COMP_NONE;
  size_t gran = 0;
  e = D.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (e != CUDA_SUCCESS || gran == 0) return fail(e, "cuMemGetAllocationGranularity");
  const size_t size = (bytes + gran - 1) / gran * gran;

  CUmemGenericAllocationHandle handle;
  e = D.create(&handle, size, &prop, 0);
  if (e != CUDA_SUCCESS && want) {  // compressible backing refused: plain device memory serves as well
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_NONE;
    e = D.create(&handle, size, &prop, 0);
  }
  if (e != CUDA_SUCCESS) return fail(e, "cuMemCreate", RT_ERR_NOMEM);
  CUmemAllocationProp got = {};
  const bool granted = D.propsOf(&got, handle) == CUDA_SUCCESS &&
                       got.allocFlags.compressionType == CU_MEM_ALLOCATION_COMP_GENERIC;

  CUdeviceptr ptr = 0;
  e = D.reserve(&ptr, size, 0, 0, 0);
  if (e != CUDA_SUCCESS) {
    D.release(handle);
    return fail(e, "cuMemAddressReserve", RT_ERR_NOMEM);
  }
  e = D.map(ptr, size, 0, handle, 0);
  if (e == CUDA_SUCCESS) {
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    e = D.setAccess(ptr, size, &acc, 1);
    if (e != CUDA_SUCCESS) D.unmap(ptr, size);
  }
  if (e != CUDA_SUCCESS) {
    D.addressFree(ptr, size);
    D.release(handle);
    return fail(e, "cuMemMap / cuMemSetAccess");
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_blocks[(void*)ptr] = Block{handle, size, dev};
  }
  *dev_ptr = (void*)ptr;
  if (compressible_out) *compressible_out = granted ? 1 : 0;
  return RT_OK;
}

int rt_obs_free(void* dev_ptr) {
  if (!dev_ptr) return RT_OK;
  Block b;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blocks.find(dev_ptr);
    if (it == g_blocks.end()) return RT_ERR_BAD_ARG;  // not one of ours
    b = it->second;
    g_blocks.erase(it);
  }
  RtDeviceGuard on_device(b.device);
  cudaDeviceSynchronize();  // nothing may still be writing the block when its mapping goes away
  const Driver& D = driver();
  const CUdeviceptr ptr = (CUdeviceptr)dev_ptr;
  CUresult e = D.unmap(ptr, b.size);
  if (e != CUDA_SUCCESS) return fail(e, "cuMemUnmap");
  D.release(b.handle);
  D.addressFree(ptr, b.size);
  return RT_OK;
}

}  // extern "C"
