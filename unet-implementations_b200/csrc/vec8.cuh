// Eight consecutive channels of one NHWC pixel in the activation STORAGE type: bf16 (the production path: one
// 16-byte access) or fp32 (the verification mode `precision="fp32"`: two 16-byte accesses).  The HBM-bound kernels
// (norm, resample, head) are templates over this type, so the fp32 mode runs the SAME kernel code with fp32 storage:
// what it proves against the reference at 1e-4 is the algorithm of the bf16 path, whose only difference is rounding
// where a tensor is stored.
#pragma once
#include "ptx.cuh"

namespace b200 {

template <typename T>
struct Vec8;

template <>
struct Vec8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ static Vec8 ld_stream(const __nv_bfloat16* p) {
    Vec8 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.r.x), "=r"(v.r.y), "=r"(v.r.z), "=r"(v.r.w)
                 : "l"(p));
    return v;
  }
  __device__ __forceinline__ static Vec8 ldg(const __nv_bfloat16* p) {
    Vec8 v;
    v.r = __ldg(reinterpret_cast<const uint4*>(p));
    return v;
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ static void st(__nv_bfloat16* p, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                              pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
};

template <>
struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ static Vec8 ld_stream(const float* p) { return ldg(p); }
  __device__ __forceinline__ static Vec8 ldg(const float* p) {
    Vec8 v;
    v.a = __ldg(reinterpret_cast<const float4*>(p));
    v.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return v;
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ static void st(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// Accumulator type of the long sums (conv dot products, InstanceNorm statistics and their backward sums): fp32 on
// the bf16 path; double in the fp32 verification mode, so that its rounding noise is below the reference's own fp32
// noise (this network amplifies rounding noise by ~1e5 at random init, tests/test_gpu_fp32_mode.py).
template <typename T>
struct AccT {
  using type = float;
};
template <>
struct AccT<float> {
  using type = double;
};

// scalar element conversion for the CUDA-core convolutions
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }

}  // namespace b200
