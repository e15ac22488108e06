// libgsx: map exchange between the GPUs of one node through peer memory.
//
// The path shards over the batch axis, one process per GPU; the only traffic between GPUs is the exchange of the finished
// maps.  A map lives in its owner's store as B strided row blocks (row r of element b at  base + (b * capacity + r) * row
// bytes), so "all rows [0, n) of every element" is a pitched 2-D region.  Instead of staging that region into a
// contiguous send buffer and running a collective kernel (which takes SMs and HBM bandwidth from the fusion kernels of the
// next step), the owner publishes the store through a CUDA IPC handle and every peer PULLS the region straight into its
// own output store with one pitched device-to-device copy per row array: the copy engines move the bytes over NVLink, no
// SM runs a communication kernel, nothing is staged, nothing is padded to the longest map of the job.
//
// The calls below are plain C: export a pointer (handle of its allocation + offset), open a peer's handle (cached: the
// torch allocator hands out the same few segments step after step), and the pitched copy.  Ordering between the processes
// is the caller's job (gradslam_b200/parallel.py: the size all-gather that precedes the pulls orders them after the
// owners' fusion, a one-word collective after them releases the owners' stores).
#include <cuda.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "gsx_common.cuh"
#include "../../include/gsx.h"

namespace gsx {

typedef CUresult (*AddressRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);

static AddressRangeFn address_range_fn() {
  static AddressRangeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (AddressRangeFn)p;
    cudaGetLastError();
  }
  return fn;
}

static std::mutex g_peer_mutex;
// (device, handle bytes) -> mapped base address in this process
static std::map<std::pair<int, std::string>, void *> g_peer_open;

}  // namespace gsx

using namespace gsx;

static_assert(sizeof(cudaIpcMemHandle_t) == GSX_IPC_HANDLE_BYTES, "GSX_IPC_HANDLE_BYTES");

extern "C" int gsx_peer_export(const void *ptr, unsigned char *handle, int64_t *offset, int64_t *allocation_bytes) {
  GSX_CHECK_ARG(ptr && handle && offset, "gsx_peer_export: null argument");
  const AddressRangeFn range = address_range_fn();
  GSX_CHECK_ARG(range != nullptr, "gsx_peer_export: cuMemGetAddressRange is not available from this driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = range(&base, &size, (CUdeviceptr)(uintptr_t)ptr);
  GSX_CHECK_ARG(r == CUDA_SUCCESS, "gsx_peer_export: pointer %p is not device memory (cuMemGetAddressRange: %d)", ptr,
                (int)r);
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("gsx_peer_export: cudaIpcGetMemHandle: %s (allocations of an expandable-segments / VMM allocator cannot "
              "be exported this way)", cudaGetErrorString(e));
    return 2;
  }
  memcpy(handle, &h, sizeof(h));
  *offset = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
  if (allocation_bytes) *allocation_bytes = (int64_t)size;
  return 0;
}

extern "C" int gsx_peer_open(const unsigned char *handle, int64_t offset, void **ptr_out) {
  GSX_CHECK_ARG(handle && ptr_out && offset >= 0, "gsx_peer_open: bad argument");
  int dev = 0;
  cudaGetDevice(&dev);
  const std::pair<int, std::string> key(dev, std::string((const char *)handle, GSX_IPC_HANDLE_BYTES));
  std::lock_guard<std::mutex> lock(g_peer_mutex);
  auto it = g_peer_open.find(key);
  if (it == g_peer_open.end()) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *base = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("gsx_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
      return 2;
    }
    it = g_peer_open.emplace(key, base).first;
  }
  *ptr_out = (char *)it->second + offset;
  return 0;
}

extern "C" int gsx_peer_close_all(void) {
  std::lock_guard<std::mutex> lock(g_peer_mutex);
  int rc = 0;
  for (auto &kv : g_peer_open) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(kv.first.first);
    if (cudaIpcCloseMemHandle(kv.second) != cudaSuccess) {
      cudaGetLastError();
      rc = 2;
    }
    cudaSetDevice(cur);
  }
  g_peer_open.clear();
  if (rc) set_error("gsx_peer_close_all: cudaIpcCloseMemHandle failed for at least one mapping");
  return rc;
}

extern "C" int gsx_peer_copy_rows(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                                  int64_t width_bytes, int64_t n_blocks, void *stream) {
  GSX_CHECK_ARG(width_bytes >= 0 && n_blocks >= 0 && width_bytes <= dst_pitch_bytes && width_bytes <= src_pitch_bytes,
                "gsx_peer_copy_rows: width %lld exceeds a pitch (%lld, %lld)", (long long)width_bytes,
                (long long)dst_pitch_bytes, (long long)src_pitch_bytes);
  if (width_bytes == 0 || n_blocks == 0) return 0;
  GSX_CHECK_ARG(dst && src, "gsx_peer_copy_rows: null pointer");
  cudaError_t e = cudaSuccess;
  constexpr int64_t kMaxPitch = 0x7fffffffll;  // cudaDeviceProp::memPitch
  if (dst_pitch_bytes <= kMaxPitch && src_pitch_bytes <= kMaxPitch) {
    e = cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes,
                          (size_t)n_blocks, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  } else {  // stores beyond the 2-D copy's pitch limit: one linear copy per element
    for (int64_t b = 0; b < n_blocks && e == cudaSuccess; ++b)
      e = cudaMemcpyAsync((char *)dst + b * dst_pitch_bytes, (const char *)src + b * src_pitch_bytes, (size_t)width_bytes,
                          cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("gsx_peer_copy_rows: cudaMemcpy2DAsync: %s", cudaGetErrorString(e));
    return 2;
  }
  return 0;
}
