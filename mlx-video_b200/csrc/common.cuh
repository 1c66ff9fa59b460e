// Host-side plumbing shared by the C-ABI entry points: error reporting, argument checks,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency, so the
// library loads on a box without a GPU driver).
#pragma once

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include "../../include/ltxb.h"

namespace ltxb {

int set_error(int code, const char* fmt, ...);

#define LTXB_CHECK_ARG(cond, ...)                                 \
  do {                                                            \
    if (!(cond)) return ::ltxb::set_error(LTXB_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)
#define LTXB_CHECK_SUPPORTED(cond, ...)                               \
  do {                                                                \
    if (!(cond)) return ::ltxb::set_error(LTXB_ERR_UNSUPPORTED, __VA_ARGS__); \
  } while (0)
#define LTXB_CUDA(call)                                                                            \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return ::ltxb::set_error(LTXB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                               __FILE__, __LINE__);                                                \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// bf16 tensor map, 128-byte swizzle, zero OOB fill. dims/strides innermost-first; strides in bytes for
// dims 1..rank-1. box innermost extent must be 64 elements (128 B = the swizzle span).
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

// 2-D map without swizzle: TMA STORES of a row-major shared-memory tile, loads of raw byte tiles (elem_bytes 1 = bytes, 2 = bf16, 4 = f32): dims / box
// innermost-first, row stride in bytes.
int encode_tmap_plain_2d(CUtensorMap* map, const void* base, int elem_bytes, const uint64_t* dims, uint64_t row_stride_bytes,
                         const uint32_t* box);

int num_sms();  // of the CURRENT device (cached per device index)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE property of a kernel: one flag per (kernel
// instantiation, device), so a process that moves between GPUs configures each of them once.
constexpr int kMaxDeviceSlots = 16;
struct PerDeviceOnce {
  bool done[kMaxDeviceSlots] = {};
  // true the first time it is called on the current device (devices beyond the table: every time)
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDeviceSlots) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// Launch with programmatic stream serialization (PDL) so the kernel may begin while its predecessor in the
// stream drains; all kernels of this library call pdl_wait() before touching global memory.  LTXB_PDL=0 disables.
bool pdl_enabled();
extern std::atomic<long long> g_kernel_launches;  // every kernel this library has launched (ltxb_kernel_launches)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ltxb
