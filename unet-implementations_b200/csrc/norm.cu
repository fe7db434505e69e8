// InstanceNorm2d(affine) + LeakyReLU + SpatialDropout2d, forward and backward, as HBM-bound fused passes over
// NHWC bf16 tensors.  Reference: Our_UNet/models/unet.py:118-127 (IN, LeakyReLU), :22-35 (SpatialDropout2d).
//
// Forward:   stats partials (produced by the conv epilogue) --finalize--> per-(n,c) affine (a, b)
//            z = leaky_relu(a*y + b),  a = s*gamma*rstd,  b = s*(beta - mean*gamma*rstd),  s = dropout scale >= 0.
// Backward:  g = dz * lrelu'(a*y+b) * s,  xh = (y-mean)*rstd
//            dy = gamma*rstd*(g - mean_hw(g) - xh*mean_hw(g*xh)),  dgamma = sum g*xh,  dbeta = sum g.
// Each thread owns 8 consecutive channels of a pixel (one 16-byte access); per-image parameter vectors are staged
// in shared memory once per block.
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kNormThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__global__ void in_finalize_kernel(const float* __restrict__ stats, int P, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ drop, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ a,
                                   float* __restrict__ b, int N, int C, double inv_hw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  double s1 = 0.0, s2 = 0.0;
  const float2* sp = reinterpret_cast<const float2*>(stats) + static_cast<int64_t>(n) * P * C + c;
  for (int p = 0; p < P; ++p) {
    const float2 v = sp[static_cast<int64_t>(p) * C];
    s1 += v.x;
    s2 += v.y;
  }
  const double m = s1 * inv_hw;
  double var = s2 * inv_hw - m * m;
  if (var < 0.0) var = 0.0;
  const double r = 1.0 / sqrt(var + static_cast<double>(eps));
  const float s = drop ? drop[i] : 1.f;
  const double gr = static_cast<double>(gamma[c]) * r;
  mean[i] = static_cast<float>(m);
  rstd[i] = static_cast<float>(r);
  a[i] = static_cast<float>(s * gr);
  b[i] = static_cast<float>(s * (static_cast<double>(beta[c]) - m * gr));
}

// grid (blocks_per_image, N); shared: a[C], b[C]
__global__ void __launch_bounds__(kNormThreads) in_apply_kernel(const __nv_bfloat16* __restrict__ y, int64_t yp,
                                                                 const float* __restrict__ a,
                                                                 const float* __restrict__ b, float slope,
                                                                 __nv_bfloat16* __restrict__ z, int64_t zp, int64_t HW,
                                                                 int C) {
  extern __shared__ float sm[];
  float* sa = sm;
  float* sb = sm + C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    sa[c] = a[n * C + c];
    sb[c] = b[n * C + c];
  }
  __syncthreads();
  const int c8n = C >> 3;
  const int64_t items = HW * c8n;
  const __nv_bfloat16* yb = y + static_cast<int64_t>(n) * HW * yp;
  __nv_bfloat16* zb = z + static_cast<int64_t>(n) * HW * zp;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kNormThreads + threadIdx.x; i < items;
       i += static_cast<int64_t>(gridDim.x) * kNormThreads) {
    const int64_t px = i / c8n;
    const int c0 = static_cast<int>(i - px * c8n) << 3;
    float f[8];
    unpack8(ld_stream(yb + px * yp + c0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = fmaf(sa[c0 + j], f[j], sb[c0 + j]);
      f[j] = t > 0.f ? t : t * slope;
    }
    *reinterpret_cast<uint4*>(zb + px * zp + c0) = pack8(f);
  }
}

struct InBwdArgs {
  const __nv_bfloat16* dz;
  int64_t dzp;
  const __nv_bfloat16* dz2;
  int64_t dz2p;
  const __nv_bfloat16* y;
  int64_t yp;
  const float *a, *b, *mean, *rstd, *drop;
  float slope;
  int64_t HW;
  int C;
};

__device__ __forceinline__ void in_bwd_g_xh(const InBwdArgs& A, const float* sa, const float* sb, const float* sm_,
                                            const float* sr, const float* ss, int n, int64_t px, int c0,
                                            float (&g)[8], float (&xh)[8]) {
  float yv[8], d[8];
  unpack8(ld_stream(A.y + (static_cast<int64_t>(n) * A.HW + px) * A.yp + c0), yv);
  unpack8(ld_stream(A.dz + (static_cast<int64_t>(n) * A.HW + px) * A.dzp + c0), d);
  if (A.dz2) {
    float d2[8];
    unpack8(ld_stream(A.dz2 + (static_cast<int64_t>(n) * A.HW + px) * A.dz2p + c0), d2);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] += d2[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float pre = fmaf(sa[c0 + j], yv[j], sb[c0 + j]);
    g[j] = d[j] * (pre > 0.f ? 1.f : A.slope) * ss[c0 + j];
    xh[j] = (yv[j] - sm_[c0 + j]) * sr[c0 + j];
  }
}

// grid (P, N): block p of image n reduces pixels [p*chunk, (p+1)*chunk) for every channel.
// shared: 5 parameter vectors [C] + reduction scratch [kNormThreads][16]
__global__ void __launch_bounds__(kNormThreads) in_bwd_reduce_kernel(InBwdArgs A, float* __restrict__ part, int P,
                                                                      int64_t chunk) {
  extern __shared__ float sm[];
  const int C = A.C;
  float* sa = sm;
  float* sb = sa + C;
  float* smn = sb + C;
  float* sr = smn + C;
  float* ss = sr + C;
  float* red = ss + C;  // [kNormThreads][16]
  const int n = blockIdx.y, p = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    sa[c] = A.a[n * C + c];
    sb[c] = A.b[n * C + c];
    smn[c] = A.mean[n * C + c];
    sr[c] = A.rstd[n * C + c];
    ss[c] = A.drop ? A.drop[n * C + c] : 1.f;
  }
  __syncthreads();
  const int c8n = C >> 3;                 // threads per pixel
  const int lanes = kNormThreads / c8n;   // pixels per sweep (C <= 8*kNormThreads, checked on the host)
  const int my_c8 = threadIdx.x % c8n;
  const int my_lane = threadIdx.x / c8n;
  const int c0 = my_c8 << 3;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const int64_t lo = p * chunk;
  int64_t hi = lo + chunk;
  if (hi > A.HW) hi = A.HW;
  if (my_lane < lanes) {
    for (int64_t px = lo + my_lane; px < hi; px += lanes) {
      float g[8], xh[8];
      in_bwd_g_xh(A, sa, sb, smn, sr, ss, n, px, c0, g, xh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += g[j];
        s2[j] = fmaf(g[j], xh[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 16 + j] = s1[j];
    red[threadIdx.x * 16 + 8 + j] = s2[j];
  }
  __syncthreads();
  // thread t < C*2 sums column (c, k) over the pixel lanes in fixed order
  for (int t = threadIdx.x; t < C * 2; t += kNormThreads) {
    const int c = t >> 1, k = t & 1;
    const int c8 = c >> 3, j = c & 7;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(l * c8n + c8) * 16 + k * 8 + j];
    part[((static_cast<int64_t>(n) * P + p) * C + c) * 2 + k] = s;
  }
}

__global__ void in_bwd_finalize_kernel(const float* __restrict__ part, int P, const float* __restrict__ gamma,
                                       const float* __restrict__ rstd, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef, int N, int C,
                                       double inv_hw) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dg = 0.0, db = 0.0;
  for (int n = 0; n < N; ++n) {
    double s1 = 0.0, s2 = 0.0;
    const float2* sp = reinterpret_cast<const float2*>(part) + static_cast<int64_t>(n) * P * C + c;
    for (int p = 0; p < P; ++p) {
      const float2 v = sp[static_cast<int64_t>(p) * C];
      s1 += v.x;
      s2 += v.y;
    }
    db += s1;
    dg += s2;
    float* co = coef + (static_cast<int64_t>(n) * C + c) * 3;
    co[0] = gamma[c] * rstd[n * C + c];
    co[1] = static_cast<float>(s1 * inv_hw);
    co[2] = static_cast<float>(s2 * inv_hw);
  }
  dgamma[c] = static_cast<float>(dg);
  dbeta[c] = static_cast<float>(db);
}

// grid (blocks_per_image, N); shared: 5 parameter vectors [C] + coef [C][3]
__global__ void __launch_bounds__(kNormThreads) in_bwd_apply_kernel(InBwdArgs A, const float* __restrict__ coef,
                                                                     __nv_bfloat16* __restrict__ dy, int64_t dyp) {
  extern __shared__ float sm[];
  const int C = A.C;
  float* sa = sm;
  float* sb = sa + C;
  float* smn = sb + C;
  float* sr = smn + C;
  float* ss = sr + C;
  float* sc = ss + C;  // [C][3]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    sa[c] = A.a[n * C + c];
    sb[c] = A.b[n * C + c];
    smn[c] = A.mean[n * C + c];
    sr[c] = A.rstd[n * C + c];
    ss[c] = A.drop ? A.drop[n * C + c] : 1.f;
    sc[c * 3 + 0] = coef[(static_cast<int64_t>(n) * C + c) * 3 + 0];
    sc[c * 3 + 1] = coef[(static_cast<int64_t>(n) * C + c) * 3 + 1];
    sc[c * 3 + 2] = coef[(static_cast<int64_t>(n) * C + c) * 3 + 2];
  }
  __syncthreads();
  const int c8n = C >> 3;
  const int64_t items = A.HW * c8n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kNormThreads + threadIdx.x; i < items;
       i += static_cast<int64_t>(gridDim.x) * kNormThreads) {
    const int64_t px = i / c8n;
    const int c0 = static_cast<int>(i - px * c8n) << 3;
    float g[8], xh[8], o[8];
    in_bwd_g_xh(A, sa, sb, smn, sr, ss, n, px, c0, g, xh);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* k = sc + (c0 + j) * 3;
      o[j] = k[0] * (g[j] - k[1] - xh[j] * k[2]);
    }
    *reinterpret_cast<uint4*>(dy + (static_cast<int64_t>(n) * A.HW + px) * dyp + c0) = pack8(o);
  }
}

static int elementwise_blocks(int64_t items, int N) {
  // enough blocks for ~8 waves over the chip, each thread doing several 16-byte items
  int64_t b = ceil_div64(items, static_cast<int64_t>(kNormThreads) * 4);
  const int64_t cap = ceil_div64(static_cast<int64_t>(num_sms()) * 16, N);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

static int check_nhwc(const char* what, int C, int64_t p0, int64_t p1, int64_t p2) {
  if (C % 8 != 0 || C <= 0 || C > 2048) return set_error(kErrInvalid, "%s: C=%d must be a multiple of 8 in (0,2048]", what, C);
  if (p0 % 8 || p1 % 8 || p2 % 8) return set_error(kErrInvalid, "%s: pitches must be multiples of 8 elements", what);
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200unet_in_finalize(const float* stats, int P, const float* gamma, const float* beta,
                                    const float* drop_scale, float eps, float* mean, float* rstd, float* a, float* b,
                                    int N, int C, int64_t HW, void* stream) {
  B200_CHECK_ARG(stats && gamma && beta && mean && rstd && a && b, "in_finalize: null pointer");
  B200_CHECK_ARG(P > 0 && HW > 0, "in_finalize: bad sizes");
  in_finalize_kernel<<<ceil_div(N * C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, P, gamma, beta, drop_scale, eps, mean, rstd, a, b, N, C, 1.0 / static_cast<double>(HW));
  B200_LAUNCH_CHECK("in_finalize_kernel");
  return 0;
}

extern "C" int b200unet_in_apply(const void* y, int64_t y_pitch, const float* a, const float* b, float slope, void* z,
                                 int64_t z_pitch, int N, int64_t HW, int C, void* stream) {
  B200_CHECK_ARG(y && a && b && z, "in_apply: null pointer");
  int rc = check_nhwc("in_apply", C, y_pitch, z_pitch, 0);
  if (rc) return rc;
  const int blocks = elementwise_blocks(HW * (C / 8), N);
  in_apply_kernel<<<dim3(blocks, N), kNormThreads, 2 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), y_pitch, a, b, slope, static_cast<__nv_bfloat16*>(z), z_pitch, HW, C);
  B200_LAUNCH_CHECK("in_apply_kernel");
  return 0;
}

extern "C" int b200unet_in_bwd_partials(int64_t HW, int C) {
  (void)C;
  int64_t p = HW / 1024;
  if (p < 1) p = 1;
  if (p > 64) p = 64;
  return static_cast<int>(p);
}

static InBwdArgs make_bwd_args(const void* dz, int64_t dz_pitch, const void* dz2, int64_t dz2_pitch, const void* y,
                               int64_t y_pitch, const float* a, const float* b, const float* mean, const float* rstd,
                               const float* drop, float slope, int64_t HW, int C) {
  InBwdArgs A;
  A.dz = static_cast<const __nv_bfloat16*>(dz);
  A.dzp = dz_pitch;
  A.dz2 = static_cast<const __nv_bfloat16*>(dz2);
  A.dz2p = dz2_pitch;
  A.y = static_cast<const __nv_bfloat16*>(y);
  A.yp = y_pitch;
  A.a = a;
  A.b = b;
  A.mean = mean;
  A.rstd = rstd;
  A.drop = drop;
  A.slope = slope;
  A.HW = HW;
  A.C = C;
  return A;
}

extern "C" int b200unet_in_bwd_reduce(const void* dz, int64_t dz_pitch, const void* dz2, int64_t dz2_pitch,
                                      const void* y, int64_t y_pitch, const float* a, const float* b,
                                      const float* mean, const float* rstd, const float* drop_scale, float slope,
                                      float* part, int N, int64_t HW, int C, void* stream) {
  B200_CHECK_ARG(dz && y && a && b && mean && rstd && part, "in_bwd_reduce: null pointer");
  int rc = check_nhwc("in_bwd_reduce", C, dz_pitch, dz2 ? dz2_pitch : 0, y_pitch);
  if (rc) return rc;
  const int P = b200unet_in_bwd_partials(HW, C);
  InBwdArgs A = make_bwd_args(dz, dz_pitch, dz2, dz2_pitch, y, y_pitch, a, b, mean, rstd, drop_scale, slope, HW, C);
  const size_t smem = (5 * C + kNormThreads * 16) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    B200_CUDA(cudaFuncSetAttribute(in_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  in_bwd_reduce_kernel<<<dim3(P, N), kNormThreads, smem, static_cast<cudaStream_t>(stream)>>>(A, part, P,
                                                                                              ceil_div64(HW, P));
  B200_LAUNCH_CHECK("in_bwd_reduce_kernel");
  return 0;
}

extern "C" int b200unet_in_bwd_finalize(const float* part, int P, const float* gamma, const float* rstd, float* dgamma,
                                        float* dbeta, float* coef, int N, int C, int64_t HW, void* stream) {
  B200_CHECK_ARG(part && gamma && rstd && dgamma && dbeta && coef, "in_bwd_finalize: null pointer");
  in_bwd_finalize_kernel<<<ceil_div(C, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      part, P, gamma, rstd, dgamma, dbeta, coef, N, C, 1.0 / static_cast<double>(HW));
  B200_LAUNCH_CHECK("in_bwd_finalize_kernel");
  return 0;
}

extern "C" int b200unet_in_bwd_apply(const void* dz, int64_t dz_pitch, const void* dz2, int64_t dz2_pitch,
                                     const void* y, int64_t y_pitch, const float* a, const float* b, const float* mean,
                                     const float* rstd, const float* drop_scale, const float* coef, float slope,
                                     void* dy, int64_t dy_pitch, int N, int64_t HW, int C, void* stream) {
  B200_CHECK_ARG(dz && y && a && b && mean && rstd && coef && dy, "in_bwd_apply: null pointer");
  int rc = check_nhwc("in_bwd_apply", C, dz_pitch, dz2 ? dz2_pitch : 0, y_pitch);
  if (rc) return rc;
  B200_CHECK_ARG(dy_pitch % 8 == 0, "in_bwd_apply: dy pitch must be a multiple of 8");
  InBwdArgs A = make_bwd_args(dz, dz_pitch, dz2, dz2_pitch, y, y_pitch, a, b, mean, rstd, drop_scale, slope, HW, C);
  const int blocks = elementwise_blocks(HW * (C / 8), N);
  in_bwd_apply_kernel<<<dim3(blocks, N), kNormThreads, 8 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      A, coef, static_cast<__nv_bfloat16*>(dy), dy_pitch);
  B200_LAUNCH_CHECK("in_bwd_apply_kernel");
  return 0;
}
