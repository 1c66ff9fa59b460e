// Library-level plumbing of the C ABI: last-error string, device check, TMA tensor-map encoding.
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ltxb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  auto fn = encode_fn();
  if (fn == nullptr) return set_error(LTXB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LTXB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)",
                     static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
                     static_cast<unsigned long long>(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1);
  return LTXB_OK;
}

int encode_tmap_plain_2d(CUtensorMap* map, const void* base, int elem_bytes, const uint64_t* dims, uint64_t row_stride_bytes,
                         const uint32_t* box) {
  auto fn = encode_fn();
  if (fn == nullptr) return set_error(LTXB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
  cuuint64_t gdim[2] = {dims[0], dims[1]};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t bdim[2] = {box[0], box[1]};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base),
                  gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  // loads of narrow byte tiles (32 / 64 bytes per row and k-block): a miss brings the row's next k-blocks too
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LTXB_ERR_CUDA, "cuTensorMapEncodeTiled (plain 2-D, %d-byte elements) failed with CUresult %d (dims %llu x %llu, box %u x %u)",
                     elem_bytes, static_cast<int>(r), static_cast<unsigned long long>(dims[0]), static_cast<unsigned long long>(dims[1]), box[0], box[1]);
  return LTXB_OK;
}

std::atomic<long long> g_kernel_launches{0};

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("LTXB_PDL"); return e == nullptr || atoi(e) != 0; }();
  return on;
}

int num_sms() {
  static int sms[kMaxDeviceSlots] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  const bool cached = dev >= 0 && dev < kMaxDeviceSlots;
  if (cached && sms[dev] > 0) return sms[dev];
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cached) sms[dev] = v;
  return v;
}

}  // namespace ltxb

extern "C" const char* ltxb_last_error(void) { return ltxb::g_err; }

extern "C" int ltxb_abi_version(void) { return 4; }  // 4: ltxb_epilogue grew flags + peer_sync, ltxb_gemm_qw_bf16, ltxb_attention_fwd_peers_sync

extern "C" int64_t ltxb_kernel_launches(void) { return ltxb::g_kernel_launches.load(std::memory_order_relaxed); }

extern "C" int ltxb_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return ltxb::set_error(LTXB_ERR_NO_DEVICE, "no CUDA device is visible");
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10)
    return ltxb::set_error(LTXB_ERR_NO_DEVICE, "device %d is compute capability %d.x; this library is sm_100a only", dev,
                           major);
  return LTXB_OK;
}

extern "C" int ltxb_set_device(int32_t device) {
  LTXB_CUDA(cudaSetDevice(device));
  return LTXB_OK;
}
