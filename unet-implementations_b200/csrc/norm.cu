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
#include "vec8.cuh"
#include <stdlib.h>

namespace b200 {

constexpr int kNormThreads = 256;

// ---------------------------------------------------------------------------------------------- thread mapping
// A block owns a contiguous pixel range of one image.  Thread t owns the FIXED channel octet c0 = 8*(t % c8n) for
// every pixel it visits (lane = t / c8n strides over pixels), so all per-channel parameters live in registers:
// the kernels issue no shared-memory loads per element and are bound by HBM, not by the LSU.
struct PixelMap {
  int c8n;      // threads per pixel = C / 8
  int threads;  // c8n * lanes  (<= kNormThreads)
  int lanes;    // pixels per sweep
};
static PixelMap make_map(int C) {
  PixelMap m;
  m.c8n = C >> 3;
  m.lanes = kNormThreads / m.c8n;
  if (m.lanes < 1) m.lanes = 1;
  m.threads = m.lanes * m.c8n;
  return m;
}
// pixels per block so that the whole launch is ~8 blocks per SM, rounded to a multiple of 4 sweeps
static int blocks_per_sm() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("B200UNET_NORM_BPS");  // developer knob
    v = e ? atoi(e) : 8;
    if (v < 1) v = 8;
  }
  return v;
}
// bps = blocks per SM over the whole launch.  Measured on B200 (tools/norm_bench.py): the streaming apply kernels want
// many small blocks (32/SM: 6.1 TB/s vs 5.5 at 8/SM); the reduce kernel wants fewer (8/SM) because every block adds a
// partial that the finalize kernel has to sum.
static int64_t pixels_per_block(int64_t HW, int N, const PixelMap& m, int bps) {
  int64_t per_img = ceil_div64(static_cast<int64_t>(num_sms()) * bps, N);
  if (per_img < 1) per_img = 1;
  int64_t chunk = ceil_div64(HW, per_img);
  const int64_t q = static_cast<int64_t>(m.lanes) * 4;
  chunk = ceil_div64(chunk, q) * q;
  return chunk;
}

__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 u = *reinterpret_cast<const float4*>(p);
  const float4 v = *reinterpret_cast<const float4*>(p + 4);
  f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w; f[4] = v.x; f[5] = v.y; f[6] = v.z; f[7] = v.w;
}

// ---------------------------------------------------------------------------------------------- forward
// grid (N, C/32), 256 threads: thread (pg = t/32, c = t%32) sums partials p = pg, pg+8, ... in double; fixed-order
// combine across the 8 groups.  Coalesced: 32 channels x (sum, sumsq) = 256 contiguous bytes per partial.
__global__ void __launch_bounds__(256) in_finalize_kernel(const float* __restrict__ stats, int P,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ drop, float eps,
                                                           float* __restrict__ mean, float* __restrict__ rstd,
                                                           float* __restrict__ a, float* __restrict__ b, int C,
                                                           double inv_hw) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  __shared__ double red[8][32][2];
  const int n = blockIdx.x;
  const int cl = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    const float2* sp = reinterpret_cast<const float2*>(stats) + static_cast<int64_t>(n) * P * C + c;
    for (int p = pg; p < P; p += 32) {  // four independent loads in flight: this kernel is pure latency
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (p + 8 * u < P) ? __ldg(sp + static_cast<int64_t>(p + 8 * u) * C) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s1 += v[u].x;
        s2 += v[u].y;
      }
    }
  }
  red[pg][cl][0] = s1;
  red[pg][cl][1] = s2;
  __syncthreads();
  if (pg != 0 || c >= C) return;
  s1 = s2 = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1 += red[k][cl][0];
    s2 += red[k][cl][1];
  }
  const int i = n * C + c;
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

// grid (blocks_per_image, N)
template <typename T>
__global__ void __launch_bounds__(kNormThreads) in_apply_kernel(const T* __restrict__ y, int64_t yp,
                                                                 const float* __restrict__ a,
                                                                 const float* __restrict__ b, float slope,
                                                                 T* __restrict__ z, int64_t zp, int64_t HW,
                                                                 int C, int c8n, int lanes, int64_t chunk) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  // images in REVERSE launch order: the producing conv wrote image N-1 last, so the first blocks find their input in
  // L2; this kernel then leaves image 0 in L2 for the next conv, which starts there
  const int n = gridDim.y - 1 - blockIdx.y;
  const int c0 = (threadIdx.x % c8n) << 3;
  const int lane = threadIdx.x / c8n;
  float ra[8], rb[8];
  ld8f(a + n * C + c0, ra);
  ld8f(b + n * C + c0, rb);
  const int64_t lo = blockIdx.x * chunk;
  int64_t hi = lo + chunk;
  if (hi > HW) hi = HW;
  const T* yb = y + static_cast<int64_t>(n) * HW * yp + c0;
  T* zb = z + static_cast<int64_t>(n) * HW * zp + c0;
  for (int64_t px = lo + lane; px < hi; px += 4 * lanes) {
    Vec8<T> v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (px + u * lanes < hi) v[u] = Vec8<T>::ld_stream(yb + (px + u * lanes) * yp);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (px + u * lanes >= hi) break;
      float f[8];
      v[u].unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(ra[j], f[j], rb[j]);
        f[j] = t > 0.f ? t : t * slope;
      }
      Vec8<T>::st(zb + (px + u * lanes) * zp, f);
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward
template <typename T>
struct InBwdK {
  const T* dz;
  int64_t dzp;
  const T* dz2;
  int64_t dz2p;
  const T* y;
  int64_t yp;
  const float *a, *b, *mean;
  float slope;
  int64_t HW;
  int C, c8n, lanes;
  int64_t chunk;
  int n0;  // first image of this launch
};

// T1 = sum dz*m, T2 = sum dz*m*(y - mean) over the block's pixels, m = lrelu'(a*y+b).   grid (P, images)
template <typename T, bool HAS2>
__global__ void __launch_bounds__(kNormThreads, (HAS2 || sizeof(T) != 2) ? 2 : 3) in_bwd_reduce_kernel(InBwdK<T> K, float* __restrict__ part, int P) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  using Acc = typename AccT<T>::type;
  extern __shared__ __align__(8) unsigned char red_raw[];
  Acc* red = reinterpret_cast<Acc*>(red_raw);  // [lanes][c8n][16]
  // reverse image order (the producer of dz wrote the last image last: L2 hits); the apply pass then runs forward
  // and finds the images this pass read last still in L2
  const int n = K.n0 + (gridDim.y - 1 - blockIdx.y);
  const int c8 = threadIdx.x % K.c8n;
  const int c0 = c8 << 3;
  const int lane = threadIdx.x / K.c8n;
  float ra[8], rb[8], rm[8];
  ld8f(K.a + n * K.C + c0, ra);
  ld8f(K.b + n * K.C + c0, rb);
  ld8f(K.mean + n * K.C + c0, rm);
  Acc t1[8], t2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t1[j] = t2[j] = 0;
  const int64_t lo = blockIdx.x * K.chunk;
  int64_t hi = lo + K.chunk;
  if (hi > K.HW) hi = K.HW;
  const T* yb = K.y + static_cast<int64_t>(n) * K.HW * K.yp + c0;
  const T* db = K.dz + static_cast<int64_t>(n) * K.HW * K.dzp + c0;
  const T* d2b = HAS2 ? K.dz2 + static_cast<int64_t>(n) * K.HW * K.dz2p + c0 : nullptr;
  const int lanes = K.lanes;
  constexpr int U = 4;  // pixels in flight per thread: 8-12 independent 16-byte loads cover the HBM latency
  for (int64_t px = lo + lane; px < hi; px += U * lanes) {
    Vec8<T> vy[U], vd[U], vd2[HAS2 ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (px + u * lanes < hi) {
        vy[u] = Vec8<T>::ld_stream(yb + (px + u * lanes) * K.yp);
        vd[u] = Vec8<T>::ld_stream(db + (px + u * lanes) * K.dzp);
        if (HAS2) vd2[HAS2 ? u : 0] = Vec8<T>::ld_stream(d2b + (px + u * lanes) * K.dz2p);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (px + u * lanes >= hi) break;
      float yv[8], d[8];
      vy[u].unpack(yv);
      vd[u].unpack(d);
      if (HAS2) {
        float d2[8];
        vd2[HAS2 ? u : 0].unpack(d2);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += d2[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = fmaf(ra[j], yv[j], rb[j]);
        const float gm = pre > 0.f ? d[j] : d[j] * K.slope;
        t1[j] += gm;
        t2[j] += static_cast<Acc>(gm) * static_cast<Acc>(yv[j] - rm[j]);
      }
    }
  }
  Acc* mine = red + static_cast<size_t>(threadIdx.x) * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j] = t1[j];
    mine[8 + j] = t2[j];
  }
  __syncthreads();
  // thread t < 2C sums value (c, k) over the pixel lanes in fixed order
  for (int t = threadIdx.x; t < K.C * 2; t += blockDim.x) {
    const int c = t >> 1, k = t & 1;
    const int cc8 = c >> 3, j = c & 7;
    Acc s = 0;
    for (int l = 0; l < lanes; ++l) s += red[(l * K.c8n + cc8) * 16 + k * 8 + j];
    part[((static_cast<int64_t>(n) * P + blockIdx.x) * K.C + c) * 2 + k] = static_cast<float>(s);
  }
}

// per (n, c):  S1 = s*T1 = sum g,  S2 = s*rstd*T2 = sum g*xh;  dy = A1*m*dz - A2*(y - mean) - A3 with
//   A1 = gamma*rstd*s,  A2 = gamma*rstd*rstd*S2/HW,  A3 = gamma*rstd*S1/HW.    grid (images, C/32), 256 threads
// `part2` (optional): a second set of partials (the skip-connection operand reduced by its own producer).  `raw_mean`
// (optional): the partials hold T2raw = sum gm * y instead of sum gm * (y - mean) (producer-side sums,
// b200unet_in_bwd_args.ext_part): T2 = T2raw - mean * T1, in double.
__global__ void __launch_bounds__(256) in_bwd_finalize_kernel(const float* __restrict__ part, int P,
                                                               const float* __restrict__ part2, int P2,
                                                               const float* __restrict__ raw_mean,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ drop,
                                                               float* __restrict__ coef, float* __restrict__ imgsum,
                                                               int C, int n0, double inv_hw) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  __shared__ double red[8][32][2];
  const int n = n0 + blockIdx.x;
  const int cl = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cl;
  double t1 = 0.0, t2 = 0.0;
  if (c < C) {
    for (int set = 0; set < 2; ++set) {
      const float* pp = set ? part2 : part;
      const int PP = set ? P2 : P;
      if (!pp) continue;
      const float2* sp = reinterpret_cast<const float2*>(pp) + static_cast<int64_t>(n) * PP * C + c;
      for (int p = pg; p < PP; p += 32) {  // four independent loads in flight
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          v[u] = (p + 8 * u < PP) ? sp[static_cast<int64_t>(p + 8 * u) * C] : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          t1 += v[u].x;
          t2 += v[u].y;
        }
      }
    }
  }
  red[pg][cl][0] = t1;
  red[pg][cl][1] = t2;
  __syncthreads();
  if (pg != 0 || c >= C) return;
  t1 = t2 = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    t1 += red[k][cl][0];
    t2 += red[k][cl][1];
  }
  const int i = n * C + c;
  if (raw_mean) t2 -= static_cast<double>(raw_mean[i]) * t1;
  const double s = drop ? static_cast<double>(drop[i]) : 1.0;
  const double r = rstd[i];
  const double S1 = s * t1, S2 = s * r * t2;
  const double gr = static_cast<double>(gamma[c]) * r;
  float4 k4;
  k4.x = static_cast<float>(gr * s);
  k4.y = static_cast<float>(gr * r * S2 * inv_hw);
  k4.z = static_cast<float>(gr * S1 * inv_hw);
  k4.w = 0.f;
  reinterpret_cast<float4*>(coef)[i] = k4;
  imgsum[i * 2 + 0] = static_cast<float>(S1);
  imgsum[i * 2 + 1] = static_cast<float>(S2);
}

// dgamma[c] = sum_n S2, dbeta[c] = sum_n S1 (fixed order).  (Summing the block partials of the whole batch inside the
// finalize launch instead was tried: one launch less, but the extra block row is a 150 .. 2000-load serial chain on
// the critical path.)
// grid (C/32), 256 threads: thread (ng = t/32, c = t%32) sums images n = ng, ng+8, ... in double (independent loads),
// fixed-order combine across the 8 groups -- the kernel sits between the apply pass and the next data gradient
__global__ void __launch_bounds__(256) in_bwd_param_kernel(const float* __restrict__ imgsum, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int N, int C) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  __shared__ double red[8][32][2];
  const int cl = threadIdx.x & 31, ng = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double dg = 0.0, db = 0.0;
  if (c < C) {
    const float2* sp = reinterpret_cast<const float2*>(imgsum) + c;
#pragma unroll 4
    for (int n = ng; n < N; n += 8) {
      const float2 v = sp[static_cast<int64_t>(n) * C];
      db += v.x;
      dg += v.y;
    }
  }
  red[ng][cl][0] = db;
  red[ng][cl][1] = dg;
  __syncthreads();
  if (ng != 0 || c >= C) return;
  db = dg = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    db += red[k][cl][0];
    dg += red[k][cl][1];
  }
  dgamma[c] = static_cast<float>(dg);
  dbeta[c] = static_cast<float>(db);
}

// grid (blocks_per_image, images)
template <typename T, bool HAS2>
__global__ void __launch_bounds__(kNormThreads, 2) in_bwd_apply_kernel(InBwdK<T> K, const float* __restrict__ coef,
                                                                        T* __restrict__ dy, int64_t dyp) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  const int n = K.n0 + blockIdx.y;
  const int c0 = (threadIdx.x % K.c8n) << 3;
  const int lane = threadIdx.x / K.c8n;
  // dy = k1*gm - k2*(y - mean) - k3 = k1*gm - k2*y + k3',  k3' = k2*mean - k3: 40 parameter registers per thread, so
  // that two 256-thread blocks fit an SM (the first version held 56 + 48 in-flight and ran at 12 % occupancy)
  float ra[8], rb[8], k1[8], k2[8], k3[8];
  ld8f(K.a + n * K.C + c0, ra);
  ld8f(K.b + n * K.C + c0, rb);
  {
    float rm[8];
    ld8f(K.mean + n * K.C + c0, rm);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 k = reinterpret_cast<const float4*>(coef)[n * K.C + c0 + j];
      k1[j] = k.x;
      k2[j] = k.y;
      k3[j] = fmaf(k.y, rm[j], -k.z);
    }
  }
  const int64_t lo = blockIdx.x * K.chunk;
  int64_t hi = lo + K.chunk;
  if (hi > K.HW) hi = K.HW;
  const T* yb = K.y + static_cast<int64_t>(n) * K.HW * K.yp + c0;
  const T* db = K.dz + static_cast<int64_t>(n) * K.HW * K.dzp + c0;
  const T* d2b = HAS2 ? K.dz2 + static_cast<int64_t>(n) * K.HW * K.dz2p + c0 : nullptr;
  T* ob = dy + static_cast<int64_t>(n) * K.HW * dyp + c0;
  const int lanes = K.lanes;
  constexpr int U = 4;
  for (int64_t px = lo + lane; px < hi; px += U * lanes) {
    Vec8<T> vy[U], vd[U], vd2[HAS2 ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (px + u * lanes < hi) {
        vy[u] = Vec8<T>::ld_stream(yb + (px + u * lanes) * K.yp);
        vd[u] = Vec8<T>::ld_stream(db + (px + u * lanes) * K.dzp);
        if (HAS2) vd2[HAS2 ? u : 0] = Vec8<T>::ld_stream(d2b + (px + u * lanes) * K.dz2p);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (px + u * lanes >= hi) break;
      float yv[8], d[8], o[8];
      vy[u].unpack(yv);
      vd[u].unpack(d);
      if (HAS2) {
        float d2[8];
        vd2[HAS2 ? u : 0].unpack(d2);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += d2[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = fmaf(ra[j], yv[j], rb[j]);
        const float gm = pre > 0.f ? d[j] : d[j] * K.slope;
        o[j] = fmaf(k1[j], gm, fmaf(-k2[j], yv[j], k3[j]));
      }
      Vec8<T>::st(ob + (px + u * lanes) * dyp, o);
    }
  }
}

// Small levels (32^2 and 16^2: five of the 22 units) in ONE kernel: a CTA owns one image x CG channels, whose whole
// (dz, [dz2], y) slab fits in shared memory -- 3 tensor passes over HBM instead of 5 and one launch instead of three.
// GPU time (CUDA-graph replay, tools/norm_small_bench.py): 43.8 -> 35.9 us at 32^2 x 512, 21.0 -> 11.7 us at 16^2 x 512.  Phase 1
// streams the slab into shared memory while accumulating T1, T2; the block reduces them in fixed order, 32/64 threads
// evaluate the per-channel coefficients exactly as in_bwd_finalize_kernel does; phase 2 re-reads the slab from shared
// memory and writes dy.  grid (C / CG, images), 256 threads, bf16 storage only.
template <bool HAS2>
__global__ void __launch_bounds__(256, 2) in_bwd_fused_kernel(InBwdK<__nv_bfloat16> K, const float* __restrict__ gamma,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ drop,
                                                               __nv_bfloat16* __restrict__ dy, int64_t dyp,
                                                               float* __restrict__ imgsum, int cg8, double inv_hw) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  using T = __nv_bfloat16;
  extern __shared__ __align__(16) unsigned char fz_raw[];
  const int HW = static_cast<int>(K.HW);
  uint4* ty = reinterpret_cast<uint4*>(fz_raw);               // [HW][cg8]
  uint4* td = ty + static_cast<size_t>(HW) * cg8;
  uint4* td2 = td + static_cast<size_t>(HW) * cg8;            // only with HAS2
  float* red = reinterpret_cast<float*>(td + static_cast<size_t>(HW) * cg8 * (HAS2 ? 2 : 1));  // [256][16]
  float* coef = red + 256 * 16;                                // [CG][4]: k1, k2, k3'
  const int n = blockIdx.y;
  const int CG = cg8 << 3;
  const int cbase = blockIdx.x * CG;
  const int c8 = threadIdx.x % cg8, lane = threadIdx.x / cg8, lanes = 256 / cg8;
  const int c0 = cbase + (c8 << 3);
  float ra[8], rb[8], rm[8];
  ld8f(K.a + n * K.C + c0, ra);
  ld8f(K.b + n * K.C + c0, rb);
  ld8f(K.mean + n * K.C + c0, rm);
  float t1[8], t2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t1[j] = t2[j] = 0.f;
  const T* yb = K.y + static_cast<int64_t>(n) * HW * K.yp + c0;
  const T* db = K.dz + static_cast<int64_t>(n) * HW * K.dzp + c0;
  const T* d2b = HAS2 ? K.dz2 + static_cast<int64_t>(n) * HW * K.dz2p + c0 : nullptr;
  constexpr int U = 4;
  for (int px = lane; px < HW; px += U * lanes) {
    Vec8<T> vy[U], vd[U], vd2[HAS2 ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (px + u * lanes < HW) {
        vy[u] = Vec8<T>::ld_stream(yb + static_cast<int64_t>(px + u * lanes) * K.yp);
        vd[u] = Vec8<T>::ld_stream(db + static_cast<int64_t>(px + u * lanes) * K.dzp);
        if (HAS2) vd2[HAS2 ? u : 0] = Vec8<T>::ld_stream(d2b + static_cast<int64_t>(px + u * lanes) * K.dz2p);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = px + u * lanes;
      if (q >= HW) break;
      ty[q * cg8 + c8] = vy[u].r;
      td[q * cg8 + c8] = vd[u].r;
      float yv[8], d[8];
      vy[u].unpack(yv);
      vd[u].unpack(d);
      if (HAS2) {
        td2[q * cg8 + c8] = vd2[HAS2 ? u : 0].r;
        float d2[8];
        vd2[HAS2 ? u : 0].unpack(d2);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += d2[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = fmaf(ra[j], yv[j], rb[j]);
        const float gm = pre > 0.f ? d[j] : d[j] * K.slope;
        t1[j] += gm;
        t2[j] += gm * (yv[j] - rm[j]);
      }
    }
  }
  float* mine = red + threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j] = t1[j];
    mine[8 + j] = t2[j];
  }
  __syncthreads();
  // thread t < CG: channel cbase + t -- sums over the pixel lanes in fixed order, then the coefficients of
  // in_bwd_finalize_kernel (same expressions, double)
  if (threadIdx.x < CG) {
    const int cl = threadIdx.x, cc8 = cl >> 3, j = cl & 7;
    float s1 = 0.f, s2 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      s1 += red[(l * cg8 + cc8) * 16 + j];
      s2 += red[(l * cg8 + cc8) * 16 + 8 + j];
    }
    const int c = cbase + cl;
    const int i = n * K.C + c;
    const double sc = drop ? static_cast<double>(drop[i]) : 1.0;
    const double r = rstd[i];
    const double S1 = sc * static_cast<double>(s1), S2 = sc * r * static_cast<double>(s2);
    const double gr = static_cast<double>(gamma[c]) * r;
    const float k1 = static_cast<float>(gr * sc);
    const float k2 = static_cast<float>(gr * r * S2 * inv_hw);
    const float k3 = static_cast<float>(gr * S1 * inv_hw);
    coef[cl * 4 + 0] = k1;
    coef[cl * 4 + 1] = k2;
    coef[cl * 4 + 2] = fmaf(k2, K.mean[i], -k3);
    imgsum[i * 2 + 0] = static_cast<float>(S1);
    imgsum[i * 2 + 1] = static_cast<float>(S2);
  }
  __syncthreads();
  float k1[8], k2[8], k3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    k1[j] = coef[((c8 << 3) + j) * 4 + 0];
    k2[j] = coef[((c8 << 3) + j) * 4 + 1];
    k3[j] = coef[((c8 << 3) + j) * 4 + 2];
  }
  T* ob = dy + static_cast<int64_t>(n) * HW * dyp + c0;
  for (int q = lane; q < HW; q += lanes) {  // this thread's own shared-memory entries
    Vec8<T> vy, vd;
    vy.r = ty[q * cg8 + c8];
    vd.r = td[q * cg8 + c8];
    float yv[8], d[8], o[8];
    vy.unpack(yv);
    vd.unpack(d);
    if (HAS2) {
      Vec8<T> v2;
      v2.r = td2[q * cg8 + c8];
      float d2[8];
      v2.unpack(d2);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] += d2[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float pre = fmaf(ra[j], yv[j], rb[j]);
      const float gm = pre > 0.f ? d[j] : d[j] * K.slope;
      o[j] = fmaf(k1[j], gm, fmaf(-k2[j], yv[j], k3[j]));
    }
    Vec8<T>::st(ob + static_cast<int64_t>(q) * dyp, o);
  }
}

// channels per CTA of the fused form, 0 = not applicable (slab + 17 KB of scratch must fit 220 KB of shared memory)
static int bwd_fused_group(int64_t HW, int C, bool has2) {
  static const bool off = [] { const char* e = getenv("B200UNET_NO_NORM_FUSED"); return e && e[0] == '1'; }();  // A/B knob
  if (off || HW > 4096) return 0;
  if (has2 && HW > 256) return 0;  // measured at 32^2 x 512: 61.6 us fused against 50.0 us for the three kernels
  const int64_t per_ch = HW * 2 * (has2 ? 3 : 2);
  const int64_t budget = 96 * 1024;  // two CTAs per SM: the phases of one overlap the other's
  for (int cg = 64; cg >= 16; cg >>= 1)
    if (C % cg == 0 && per_ch * cg <= budget && HW >= 256 / (cg / 8)) return cg;
  return 0;
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
  B200_CHECK_ARG(P > 0 && HW > 0 && N > 0 && C > 0, "in_finalize: bad sizes");
  launch_k(in_finalize_kernel, dim3(N, ceil_div(C, 32)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      stats, P, gamma, beta, drop_scale, eps, mean, rstd, a, b, C, 1.0 / static_cast<double>(HW));
  B200_LAUNCH_CHECK("in_finalize_kernel");
  return 0;
}

template <typename T>
static int in_apply_impl(const void* y, int64_t y_pitch, const float* a, const float* b, float slope, void* z,
                         int64_t z_pitch, int N, int64_t HW, int C, void* stream) {
  B200_CHECK_ARG(y && a && b && z, "in_apply: null pointer");
  int rc = check_nhwc("in_apply", C, y_pitch, z_pitch, 0);
  if (rc) return rc;
  const PixelMap m = make_map(C);
  const int64_t chunk = pixels_per_block(HW, N, m, 4 * blocks_per_sm());
  launch_k(in_apply_kernel<T>, dim3((unsigned)ceil_div64(HW, chunk), N), dim3(m.threads), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const T*>(y), y_pitch, a, b, slope, static_cast<T*>(z), z_pitch, HW, C, m.c8n, m.lanes, chunk);
  B200_LAUNCH_CHECK("in_apply_kernel");
  return 0;
}

extern "C" int b200unet_in_apply(const void* y, int64_t y_pitch, const float* a, const float* b, float slope, void* z,
                                 int64_t z_pitch, int N, int64_t HW, int C, void* stream) {
  return in_apply_impl<__nv_bfloat16>(y, y_pitch, a, b, slope, z, z_pitch, N, HW, C, stream);
}
extern "C" int b200unet_in_apply_f32(const void* y, int64_t y_pitch, const float* a, const float* b, float slope,
                                     void* z, int64_t z_pitch, int N, int64_t HW, int C, void* stream) {
  return in_apply_impl<float>(y, y_pitch, a, b, slope, z, z_pitch, N, HW, C, stream);
}

// Images per launch group.  Splitting the batch into L2-sized groups (so the apply pass re-reads dz and y from L2)
// was measured SLOWER on B200 at 512^2 x 32 channels: one image per group means ~10 us kernels whose launch gaps and
// ramp-up/tail cost more than the saved HBM reads (13.7 ms vs 9.2 ms per step over the 22 layers).  The whole batch
// therefore goes in one group; the partial sums are to move into the producers' epilogues instead.
static int bwd_images_per_chunk(int N, int64_t HW, int C, bool has_dz2) {
  (void)HW;
  (void)C;
  (void)has_dz2;
  return N;
}
static int bwd_partials(int N, int64_t HW, int C, bool has_dz2) {
  const PixelMap m = make_map(C);
  const int ipc = bwd_images_per_chunk(N, HW, C, has_dz2);
  const int64_t chunk = pixels_per_block(HW, ipc, m, blocks_per_sm());
  return static_cast<int>(ceil_div64(HW, chunk));
}

extern "C" int64_t b200unet_in_backward_workspace(int N, int64_t HW, int C) {
  // worst case over has_dz2: partials [N][P][C][2] + coef [N][C][4] + per-image sums [N][C][2]
  const int P0 = bwd_partials(N, HW, C, false), P1 = bwd_partials(N, HW, C, true);
  const int P = P0 > P1 ? P0 : P1;
  return (static_cast<int64_t>(N) * P * C * 2 + static_cast<int64_t>(N) * C * 6) * 4;
}

template <typename T>
static int in_backward_impl(const b200unet_in_bwd_args* A, void* stream) {
  B200_CHECK_ARG(A && A->dz && A->y && A->a && A->b && A->mean && A->rstd && A->gamma && A->dy && A->dgamma &&
                     A->dbeta && A->workspace,
                 "in_backward: null pointer");
  int rc = check_nhwc("in_backward", A->C, A->dz_pitch, A->dz2 ? A->dz2_pitch : 0, A->y_pitch);
  if (rc) return rc;
  B200_CHECK_ARG(A->dy_pitch % 8 == 0, "in_backward: dy pitch must be a multiple of 8");
  const int N = A->N, C = A->C;
  const int64_t HW = A->HW;
  B200_CHECK_ARG(A->workspace_bytes >= b200unet_in_backward_workspace(N, HW, C), "in_backward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool has2 = A->dz2 != nullptr;
  const PixelMap m = make_map(C);
  const int ipc = bwd_images_per_chunk(N, HW, C, has2);
  const int64_t chunk = pixels_per_block(HW, ipc, m, blocks_per_sm());
  // the backward apply re-loads ~56 per-channel parameters per thread at block start: it keeps the coarser grid
  const int64_t chunk_apply = chunk;
  const int P = static_cast<int>(ceil_div64(HW, chunk));
  // workspace layout: per-image sums [N][C][2] first (b200unet_in_bwd_params finds them without knowing P), then the
  // apply coefficients [N][C][4], then the block partials [N][P][C][2]
  float* imgsum = A->workspace;
  float* coef = imgsum + static_cast<int64_t>(N) * C * 2;
  float* part = coef + static_cast<int64_t>(N) * C * 4;
  InBwdK<T> K;
  K.dz = static_cast<const T*>(A->dz);
  K.dzp = A->dz_pitch;
  K.dz2 = static_cast<const T*>(A->dz2);
  K.dz2p = A->dz2_pitch;
  K.y = static_cast<const T*>(A->y);
  K.yp = A->y_pitch;
  K.a = A->a;
  K.b = A->b;
  K.mean = A->mean;
  K.slope = A->slope;
  K.HW = HW;
  K.C = C;
  K.c8n = m.c8n;
  K.lanes = m.lanes;
  K.chunk = chunk;
  const size_t red_bytes = static_cast<size_t>(m.threads) * 16 * sizeof(typename AccT<T>::type);
  const double inv_hw = 1.0 / static_cast<double>(HW);
  if constexpr (sizeof(T) == 2) {
    const bool ext_given = A->ext_part != nullptr && (!has2 || A->ext_part2 != nullptr);
    const int cg = ext_given ? 0 : bwd_fused_group(HW, C, has2);
    if (cg) {
      const size_t smem = static_cast<size_t>(HW) * cg * 2 * (has2 ? 3 : 2) + (256 * 16 + 64 * 4) * sizeof(float);
      auto kern = has2 ? in_bwd_fused_kernel<true> : in_bwd_fused_kernel<false>;
      static bool attr_set[2] = {false, false};
      if (!attr_set[has2 ? 1 : 0]) {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set[has2 ? 1 : 0] = true;
      }
      K.n0 = 0;
      launch_k(kern, dim3(C / cg, N), dim3(256), smem, st, K, A->gamma, A->rstd, A->drop_scale, static_cast<T*>(A->dy), A->dy_pitch,
                                              imgsum, cg / 8, inv_hw);
      B200_LAUNCH_CHECK("in_bwd_fused_kernel");
      if (!A->defer_params) {
        launch_k(in_bwd_param_kernel, dim3(ceil_div(C, 32)), dim3(256), 0, st, imgsum, A->dgamma, A->dbeta, N, C);
        B200_LAUNCH_CHECK("in_bwd_param_kernel");
      }
      return 0;
    }
  }
  // producer-side sums: the kernel(s) that wrote dz (and dz2) already reduced gm and gm * y per image over their own
  // blocks -- no pass over (dz, y) is needed before the apply pass
  const bool ext = A->ext_part != nullptr && (!has2 || A->ext_part2 != nullptr);
  for (int n0 = 0; n0 < N; n0 += ipc) {
    const int nn = (N - n0 < ipc) ? N - n0 : ipc;
    K.n0 = n0;
    if (ext) {
      launch_k(in_bwd_finalize_kernel, dim3(nn, ceil_div(C, 32)), dim3(256), 0, st, A->ext_part, A->ext_P, has2 ? A->ext_part2 : nullptr,
                                                                       A->ext_P2, A->mean, A->gamma, A->rstd, A->drop_scale,
                                                                       coef, imgsum, C, n0, inv_hw);
    } else {
      if (has2) launch_k(in_bwd_reduce_kernel<T, true>, dim3(P, nn), dim3(m.threads), red_bytes, st, K, part, P);
      else launch_k(in_bwd_reduce_kernel<T, false>, dim3(P, nn), dim3(m.threads), red_bytes, st, K, part, P);
      B200_LAUNCH_CHECK("in_bwd_reduce_kernel");
      launch_k(in_bwd_finalize_kernel, dim3(nn, ceil_div(C, 32)), dim3(256), 0, st, part, P, nullptr, 0, nullptr, A->gamma, A->rstd,
                                                                       A->drop_scale, coef, imgsum, C, n0, inv_hw);
    }
    B200_LAUNCH_CHECK("in_bwd_finalize_kernel");
    K.chunk = chunk_apply;
    if (has2)
      launch_k(in_bwd_apply_kernel<T, true>, dim3((unsigned)ceil_div64(HW, chunk_apply), nn), dim3(m.threads), 0, st, 
          K, coef, static_cast<T*>(A->dy), A->dy_pitch);
    else
      launch_k(in_bwd_apply_kernel<T, false>, dim3((unsigned)ceil_div64(HW, chunk_apply), nn), dim3(m.threads), 0, st, 
          K, coef, static_cast<T*>(A->dy), A->dy_pitch);
    K.chunk = chunk;
    B200_LAUNCH_CHECK("in_bwd_apply_kernel");
  }
  if (!A->defer_params) {
    launch_k(in_bwd_param_kernel, dim3(ceil_div(C, 32)), dim3(256), 0, st, imgsum, A->dgamma, A->dbeta, N, C);
    B200_LAUNCH_CHECK("in_bwd_param_kernel");
  }
  return 0;
}

// dgamma / dbeta of a b200unet_in_backward call made with defer_params: sums the per-image (sum g, sum g * x_hat) that call
// left in its workspace.  Off the critical path of backward (dy does not depend on it): the host runs it on the
// weight-gradient side stream.
extern "C" int b200unet_in_bwd_params(const float* workspace, int N, int64_t HW, int C, float* dgamma, float* dbeta,
                                      void* stream) {
  B200_CHECK_ARG(workspace && dgamma && dbeta && N > 0 && HW > 0 && C > 0, "in_bwd_params: null pointer or bad sizes");
  (void)HW;
  const float* imgsum = workspace;  // first block of the workspace layout (in_backward_impl)
  launch_k(in_bwd_param_kernel, dim3(ceil_div(C, 32)), dim3(256), 0, static_cast<cudaStream_t>(stream), imgsum, dgamma, dbeta, N, C);
  B200_LAUNCH_CHECK("in_bwd_param_kernel");
  return 0;
}

extern "C" int b200unet_in_backward(const b200unet_in_bwd_args* A, void* stream) {
  return in_backward_impl<__nv_bfloat16>(A, stream);
}
extern "C" int b200unet_in_backward_f32(const b200unet_in_bwd_args* A, void* stream) {
  return in_backward_impl<float>(A, stream);
}
