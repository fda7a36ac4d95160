// Host-side helpers shared by the C-ABI entry points: error reporting, TMA descriptor encoding
// (cuTensorMapEncodeTiled resolved at run time so the library links without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ovb200.h"

namespace ovb {

void set_last_error(const char* fmt, ...);

#define OVB_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::ovb::set_last_error(__VA_ARGS__); \
      return OVB_ERR_INVALID_ARGUMENT;    \
    }                                     \
  } while (0)

#define OVB_CHECK_CUDA(expr)                                                               \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::ovb::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                     \
      return OVB_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// bf16 tiled tensor map, 128-byte swizzle, zero fill out of bounds.
// dims/box are innermost-first; strides_bytes has rank-1 entries (stride of dims 1..rank-1).
int encode_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

// General form: element strides per dim (1 = dense) and swizzle span in bytes (128, 64 or 32).
int encode_tmap_bf16_ex(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box,
                        const uint32_t* elem_strides, int swizzle_bytes);

// fp32 tiled tensor map (128-byte swizzle: inner box <= 32 floats), zero fill / clipping out of bounds.
int encode_tmap_f32(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);

int device_sm_count();
// true the first time it is called for the CURRENT device with this flag array: per-device one-time
// setup (cudaFuncSetAttribute is per device; a process may drive several GPUs)
bool once_per_device(bool (&done)[64]);
void count_launch();

}  // namespace ovb
