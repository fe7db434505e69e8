// Host-side plumbing shared by every translation unit of libb200unet.so: error reporting and TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200unet.h"

namespace b200 {

enum { kErrInvalid = -1, kErrCuda = -2, kErrUnsupported = -3, kErrDriver = -4 };

int set_error(int code, const char* fmt, ...);  // defined in api.cu (thread-local message)

#define B200_CHECK_ARG(cond, ...)                                 \
  do {                                                            \
    if (!(cond)) return ::b200::set_error(::b200::kErrInvalid, __VA_ARGS__); \
  } while (0)

#define B200_CUDA(expr)                                                                                     \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return ::b200::set_error(::b200::kErrCuda, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                                   \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library in this process (api.cu)

#define B200_LAUNCH_CHECK(name)                                                                            \
  do {                                                                                                     \
    __atomic_fetch_add(&::b200::g_launches, 1ull, __ATOMIC_RELAXED);                                       \
    cudaError_t _e = cudaGetLastError();                                                                   \
    if (_e != cudaSuccess)                                                                                 \
      return ::b200::set_error(::b200::kErrCuda, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Build a tiled bf16 tensor map of `rank` dims (dim 0 innermost, stride 1).  strides_bytes[i] is the byte stride of
// dim i+1.  Returns 0 or a negative error.  Resolved through cudaGetDriverEntryPoint so the library has no link-time
// dependency on libcuda (it must load on a machine without a driver for the CPU-side symbol test).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swizzle);

int num_sms();

}  // namespace b200
