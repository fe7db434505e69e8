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

// Programmatic dependent launch (PDL).  One training step is a chain of ~230 dependent kernels, many of them 5-50 us
// long: between two of them the GPU otherwise drains kernel N completely, then launches, rasterises and runs the
// prologue (barrier init, TMEM allocation, descriptor prefetch) of kernel N+1.  Launched with the programmatic-stream-
// serialization attribute, kernel N+1's CTAs become resident as kernel N's CTAs retire and run their prologue under
// N's tail; `pdl_wait()` (griddepcontrol.wait, ptx.cuh) -- executed by EVERY thread before its first access to global
// memory -- holds them until kernel N has completed and flushed.  Every kernel launched through launch_k() therefore
// starts with pdl_launch_dependents() + (setup) + pdl_wait() (no-ops in a plain launch).
// MEASURED (round 2, batch 32, 512^2, same box, alternating runs): graph replay 21.69 / 21.63 ms with PDL against
// 21.65 / 21.83 without -- no gain: the persistent conv kernels fill every SM until their last CTA retires, so the
// next kernel gets resident no earlier -- and the eager step LOSES 0.5-1.4 ms to the heavier cudaLaunchKernelEx host
// path.  OFF by default; B200UNET_PDL=1 / b200unet_set_pdl(1) turns it on.
extern int g_pdl;  // api.cu
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Build a tiled bf16 tensor map of `rank` dims (dim 0 innermost, stride 1).  strides_bytes[i] is the byte stride of
// dim i+1.  Returns 0 or a negative error.  Resolved through cudaGetDriverEntryPoint so the library has no link-time
// dependency on libcuda (it must load on a machine without a driver for the CPU-side symbol test).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swizzle);

int num_sms();

}  // namespace b200
