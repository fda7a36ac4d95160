// Library-level entry points and host helpers (error string, TMA descriptor encoding, counters).
#include "host_util.h"

#include <atomic>
#include <mutex>
#include <string.h>

namespace ovb {

static thread_local char g_last_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap_bf16_ex(out, gptr, rank, dims, strides_bytes, box, nullptr, 128);
}

int encode_tmap_bf16_ex(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box,
                        const uint32_t* elem_strides, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return OVB_ERR_CUDA;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = elem_strides != nullptr ? elem_strides[i] : 1;
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_32B;
  if (static_cast<int>(box[0]) * 2 > swizzle_bytes) {
    set_last_error("TMA inner box of %u bf16 exceeds the %d-byte swizzle span", box[0], swizzle_bytes);
    return OVB_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstrides[i] = strides_bytes[i];
    if (gstrides[i] % 16 != 0) {
      set_last_error("TMA stride %d = %llu bytes is not a multiple of 16", i,
                     static_cast<unsigned long long>(gstrides[i]));
      return OVB_ERR_INVALID_ARGUMENT;
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(gptr), gdims, gstrides, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
                   static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
                   static_cast<unsigned long long>(rank > 1 ? dims[1] : 0), box[0],
                   rank > 1 ? box[1] : 0);
    return OVB_ERR_CUDA;
  }
  return OVB_OK;
}

int encode_tmap_f32(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return OVB_ERR_CUDA;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstrides[i] = strides_bytes[i];
    if (gstrides[i] % 16 != 0) {
      set_last_error("TMA stride %d = %llu bytes is not a multiple of 16", i,
                     static_cast<unsigned long long>(gstrides[i]));
      return OVB_ERR_INVALID_ARGUMENT;
    }
  }
  if (box[0] * 4 > 128) {
    set_last_error("TMA inner box of %u fp32 exceeds the 128-byte swizzle span", box[0]);
    return OVB_ERR_INVALID_ARGUMENT;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(gptr), gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(fp32) failed with CUresult %d", static_cast<int>(r));
    return OVB_ERR_CUDA;
  }
  return OVB_OK;
}

bool once_per_device(bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;  // unknown: redo the setup
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

}  // namespace ovb

extern "C" int ovb_abi_version(void) { return OVB_ABI_VERSION; }

extern "C" const char* ovb_last_error(void) { return ovb::g_last_error; }

extern "C" int64_t ovb_launch_count(void) { return ovb::g_launches.load(); }
extern "C" void ovb_add_launch_count(int64_t n) { ovb::g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int ovb_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    ovb::set_last_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return OVB_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    ovb::set_last_error("device %d is sm_%d%d; libovb200 contains sm_100a code only", dev, major,
                        minor);
    return OVB_ERR_UNSUPPORTED_ARCH;
  }
  return OVB_OK;
}
