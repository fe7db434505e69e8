// Library-level entry points: version, error string, device check, TMA descriptor encoder.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <mutex>

namespace b200 {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
int g_pdl = -1;  // -1: read B200UNET_PDL on first use

bool pdl_enabled() {
  int v = __atomic_load_n(&g_pdl, __ATOMIC_RELAXED);
  if (v < 0) {
    const char* e = getenv("B200UNET_PDL");
    v = (e && e[0] == '1') ? 1 : 0;  // OFF by default: measured no gain in graph replay and +0.5..1.4 ms of host time eager
    __atomic_store_n(&g_pdl, v, __ATOMIC_RELAXED);
  }
  return v != 0;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(kErrDriver, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(kErrInvalid, "TMA base address %p is not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) return set_error(kErrInvalid, "TMA stride %llu not a multiple of 16 bytes",
                                            (unsigned long long)gstr[i]);
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(kErrDriver,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] "
                     "stride0 %llu swizzle %d",
                     (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                     (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0],
                     rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0,
                     (unsigned long long)(rank > 1 ? gstr[0] : 0), (int)swizzle);
  }
  return 0;
}

// SMs the grids of this library are sized for: the device's count minus the reserved ones.  Data-parallel runs keep
// a few SMs free for the NCCL kernels of the gradient all-reduce WHILE BACKWARD RUNS (b200unet_set_reserved_sms, called
// by the fused backward when a reducer is attached), so that a persistent one-CTA-per-SM conv kernel never has a CTA
// waiting behind a communication kernel; forward kernels keep the whole device.  B200UNET_RESERVED_SMS_DEFAULT sets a
// process-wide default (developer knob).
static int g_reserved_sms = -1;  // -1: not initialised

static int device_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n = v;
  }
  return n;
}

int num_sms() {
  int r = __atomic_load_n(&g_reserved_sms, __ATOMIC_RELAXED);
  if (r < 0) {
    const char* e = getenv("B200UNET_RESERVED_SMS_DEFAULT");
    r = e ? atoi(e) : 0;
    if (r < 0) r = 0;
    __atomic_store_n(&g_reserved_sms, r, __ATOMIC_RELAXED);
  }
  const int v = device_sms();
  return (r > 0 && r < v / 2) ? v - r : v;
}

}  // namespace b200

extern "C" {

int b200unet_version(void) { return B200UNET_VERSION; }

long long b200unet_launch_count(void) {
  return static_cast<long long>(__atomic_load_n(&b200::g_launches, __ATOMIC_RELAXED));
}

const char* b200unet_last_error(void) { return b200::g_err; }

int b200unet_set_reserved_sms(int n) {
  const int v = b200::num_sms();  // forces initialisation
  (void)v;
  return __atomic_exchange_n(&b200::g_reserved_sms, n < 0 ? 0 : n, __ATOMIC_RELAXED);
}

int b200unet_set_pdl(int on) {
  const int prev = b200::pdl_enabled() ? 1 : 0;
  __atomic_store_n(&b200::g_pdl, on ? 1 : 0, __ATOMIC_RELAXED);
  return prev;
}

int b200unet_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return b200::set_error(b200::kErrCuda, "no CUDA device");
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return b200::set_error(b200::kErrCuda, "cannot query compute capability");
  return major == 10 ? 1 : 0;
}

}  // extern "C"
