// Segmentation head (1x1 conv, unet.py:374-381) and SimpleLoss (Our_UNet/models/losses.py:24-121).
//
// Loss forward is ONE pass over (logits fp32 NCHW, target int64): per pixel softmax over 3 classes, accumulating
//   n_c      = #valid pixels of class c                       (class weights, losses.py:36-60)
//   S_c      = sum_{valid, t=c} -log p_t                       (weighted CE numerator, nn.CrossEntropyLoss)
//   I_bc, P_bc = sum p_c [t=c], sum p_c over valid pixels      (Dice, losses.py:98-113)
// per block, then a one-block finalize reduces in fixed order (double) and emits the loss and the small tables the
// elementwise backward needs.  Nothing synchronises with the host (the reference's three `if class_pixels[c] == 0`
// host syncs at losses.py:53 become a device-side clamp).
#include "common.cuh"
#include "ptx.cuh"
#include "vec8.cuh"

namespace b200 {

constexpr int kNC = 3;            // classes (losses.py:40 hard-codes 3)
constexpr int kLossThreads = 256;
constexpr int kLossPxPerThread = 8;
constexpr int kLossVals = 4 * kNC;  // per block: cnt[3], S[3], I[3], P[3]

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  // deterministic: shuffle tree inside the warp, then warp 0 sums the warp results in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += scratch[w];
  return s;  // valid on thread 0
}

// Four consecutive pixels of one image per thread and step: three 16-byte logit loads (one per class plane) and one
// 32-byte (int64) or 4-byte (uint8) target load.  `vec` = HW is a multiple of 4 (every plane and image then starts
// 16-byte aligned); otherwise the same code runs on guarded scalar loads.
template <typename TT>
__device__ __forceinline__ void load_px4(const float* __restrict__ z0, const TT* __restrict__ tg, int64_t HW, int64_t px,
                                         bool vec, float (&a)[3][4], int (&t)[4]) {
  if (vec) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(z0 + c * HW + px));
      a[c][0] = v.x; a[c][1] = v.y; a[c][2] = v.z; a[c][3] = v.w;
    }
    if (sizeof(TT) == 8) {
      const longlong2 u = __ldg(reinterpret_cast<const longlong2*>(tg + px));
      const longlong2 w = __ldg(reinterpret_cast<const longlong2*>(tg + px) + 1);
      // labels outside [0, 255] can match neither a class nor a (byte-sized) ignore index: map them to a value that
      // matches nothing so that the 32-bit compare below equals the reference's 64-bit one
      const long long q[4] = {u.x, u.y, w.x, w.y};
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = (q[j] < -(1ll << 30) || q[j] > (1ll << 30)) ? (1 << 30) + 1 : static_cast<int>(q[j]);
    } else {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(tg + px));
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = static_cast<int>((u >> (8 * j)) & 0xffu);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = px + j < HW;
#pragma unroll
      for (int c = 0; c < 3; ++c) a[c][j] = in ? z0[c * HW + px + j] : 0.f;
      long long q = in ? static_cast<long long>(tg[px + j]) : -1;
      t[j] = (q < -(1ll << 30) || q > (1ll << 30)) ? (1 << 30) + 1 : static_cast<int>(q);
      if (!in) t[j] = -2;  // outside the image: counted nowhere (handled by the caller through `in`)
    }
  }
}

// grid (blocks_per_image, N).  One pass; the twelve block sums are ONE round of warp shuffle trees + one fixed-order
// sum over the eight warps (the first version ran twelve serial block reductions with two barriers each and scalar
// 4-byte loads: 2.1 TB/s).
template <typename TT>
__global__ void __launch_bounds__(kLossThreads) loss_fwd_kernel(const float* __restrict__ logits,
                                                                 const TT* __restrict__ target, int ignore_index,
                                                                 float* __restrict__ part, int64_t HW) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  __shared__ float scratch[kLossThreads / 32][kLossVals];
  const int n = blockIdx.y;
  const float* z0 = logits + static_cast<int64_t>(n) * kNC * HW;
  const TT* tg = target + static_cast<int64_t>(n) * HW;
  const bool vec = (HW & 3) == 0;
  float acc[kLossVals];  // cnt[3], S[3], I[3], P[3]
#pragma unroll
  for (int i = 0; i < kLossVals; ++i) acc[i] = 0.f;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kLossThreads * kLossPxPerThread;
  constexpr int kIters = kLossPxPerThread / 4;
  float a[kIters][3][4];
  int t[kIters][4];
  bool live[kIters];
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    const int64_t px = base + (static_cast<int64_t>(k) * kLossThreads + threadIdx.x) * 4;
    live[k] = px < HW;
    if (live[k]) load_px4<TT>(z0, tg, HW, px, vec, a[k], t[k]);
  }
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    if (!live[k]) continue;
    const int64_t px = base + (static_cast<int64_t>(k) * kLossThreads + threadIdx.x) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tj = t[k][j];
      if (tj == ignore_index || (!vec && px + j >= HW)) continue;
      const float a0 = a[k][0][j], a1 = a[k][1][j], a2 = a[k][2][j];
      const float m = fmaxf(a0, fmaxf(a1, a2));
      const float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = expf(a2 - m);
      const float se = e0 + e1 + e2;
      const float inv = 1.f / se;
      const float p0 = e0 * inv, p1 = e1 * inv, p2 = e2 * inv;
      const float lse = m + logf(se);
      acc[9] += p0;
      acc[10] += p1;
      acc[11] += p2;
      if (tj == 0) { acc[0] += 1.f; acc[3] += lse - a0; acc[6] += p0; }
      else if (tj == 1) { acc[1] += 1.f; acc[4] += lse - a1; acc[7] += p1; }
      else if (tj == 2) { acc[2] += 1.f; acc[5] += lse - a2; acc[8] += p2; }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kLossVals; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) scratch[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kLossVals) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += scratch[w][threadIdx.x];
    part[(static_cast<int64_t>(n) * gridDim.x + blockIdx.x) * kLossVals + threadIdx.x] = s;
  }
}

// tables layout: [0..2] = w_c / W (CE weight over normaliser); then per image b: A[b][3], Bq[b][3]
//   dDice/dp_{ic} = A_bc*[t_i=c] + Bq_bc  for valid pixels, with
//   A = -(2/(C*B)) / (U+eps),  Bq = (1/(C*B)) * (2I+eps) / (U+eps)^2,  U = P + T
__global__ void loss_finalize_kernel(const float* __restrict__ part, int blocks, const float* __restrict__ class_w,
                                     int dynamic, float weight_ce, float weight_dice, float smooth,
                                     float* __restrict__ loss_out, float* __restrict__ tables, int N) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  // one block of 32 warps; warp w reduces values i = w, w + 32, ... (value v of image b): lanes stride over the
  // partials, fixed-order shuffle tree (the serial version took 46 us on the critical path between forward and backward)
  extern __shared__ double sred[];  // [N][12] sums, then [N][3] per-(image, class) Dice ratios
  double* dterm = sred + N * kLossVals;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // four values per warp at a time: their loads are independent (one value at a time was a chain of dependent L2 round
  // trips, 12 per warp); each value keeps its own fixed-order sum
  constexpr int UV = 4;
  const int total = N * kLossVals;
  for (int i = warp * UV; i < total; i += nwarps * UV) {
    double s[UV];
    const float* src[UV];
#pragma unroll
    for (int u = 0; u < UV; ++u) {
      const int iu = min(i + u, total - 1);
      const int b = iu / kLossVals, v = iu - b * kLossVals;
      src[u] = part + static_cast<int64_t>(b) * blocks * kLossVals + v;
      s[u] = 0.0;
    }
    for (int k = lane; k < blocks; k += 32) {
      float x[UV];
#pragma unroll
      for (int u = 0; u < UV; ++u) x[u] = src[u][static_cast<int64_t>(k) * kLossVals];
#pragma unroll
      for (int u = 0; u < UV; ++u) s[u] += x[u];
    }
#pragma unroll
    for (int u = 0; u < UV; ++u) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
      if (lane == 0 && i + u < total) sred[i + u] = s[u];
    }
  }
  __syncthreads();
  // the per-(image, class) Dice terms and backward tables in parallel (96 double divisions were a serial tail)
  if (threadIdx.x < N * kNC) {
    const int b = threadIdx.x / kNC, c = threadIdx.x - b * kNC;
    const double T = sred[b * kLossVals + c];
    const double I = sred[b * kLossVals + 2 * kNC + c];
    const double Pc = sred[b * kLossVals + 3 * kNC + c];
    const double U = Pc + T + smooth;
    const double two_i = 2.0 * I + smooth;
    dterm[b * kNC + c] = two_i / U;
    tables[kNC + (b * 2 + 0) * kNC + c] = static_cast<float>(-(2.0 / (kNC * static_cast<double>(N))) / U);
    tables[kNC + (b * 2 + 1) * kNC + c] = static_cast<float>((1.0 / (kNC * static_cast<double>(N))) * two_i / (U * U));
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double cnt[kNC] = {0, 0, 0}, S[kNC] = {0, 0, 0};
  for (int b = 0; b < N; ++b)
    for (int c = 0; c < kNC; ++c) {
      cnt[c] += sred[b * kLossVals + c];
      S[c] += sred[b * kLossVals + kNC + c];
    }
  // class weights (fp32 arithmetic mirrors losses.py:44-60)
  float w[kNC];
  if (class_w && !dynamic) {
    for (int c = 0; c < kNC; ++c) w[c] = class_w[c];
  } else if (dynamic) {
    const float total = static_cast<float>(cnt[0] + cnt[1] + cnt[2]);
    float wsum = 0.f;
    for (int c = 0; c < kNC; ++c) {
      const float px = cnt[c] == 0.0 ? 1.f : static_cast<float>(cnt[c]);
      w[c] = total / px;
      wsum += w[c];
    }
    for (int c = 0; c < kNC; ++c) w[c] = w[c] * (static_cast<float>(kNC) / wsum);
  } else {
    for (int c = 0; c < kNC; ++c) w[c] = 1.f;
  }
  double num = 0.0, den = 0.0;
  for (int c = 0; c < kNC; ++c) {
    num += static_cast<double>(w[c]) * S[c];
    den += static_cast<double>(w[c]) * cnt[c];
  }
  const double ce = num / den;  // 0/0 -> NaN, like nn.CrossEntropyLoss on an all-ignored batch
  double dice_loss = 0.0;
  for (int c = 0; c < kNC; ++c) {
    double dsum = 0.0;
    for (int b = 0; b < N; ++b) dsum += dterm[b * kNC + c];  // same terms, same order as the serial form
    dice_loss += 1.0 - dsum / N;
  }
  dice_loss /= kNC;
  for (int c = 0; c < kNC; ++c) tables[c] = static_cast<float>(w[c] / den);
  loss_out[0] = static_cast<float>(weight_ce * ce + weight_dice * dice_loss);
  loss_out[1] = static_cast<float>(ce);
  loss_out[2] = static_cast<float>(dice_loss);
}

template <typename TT>
__global__ void __launch_bounds__(kLossThreads) loss_bwd_kernel(const float* __restrict__ logits,
                                                                 const TT* __restrict__ target,
                                                                 const float* __restrict__ tables,
                                                                 const float* __restrict__ grad_out, float weight_ce,
                                                                 float weight_dice, int ignore_index,
                                                                 float* __restrict__ dlogits, int64_t HW) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  const int n = blockIdx.y;
  const float gs = grad_out ? grad_out[0] : 1.f;
  const float wn0 = tables[0], wn1 = tables[1], wn2 = tables[2];
  const float* tb = tables + kNC + n * 2 * kNC;
  const float A0 = tb[0], A1 = tb[1], A2 = tb[2], B0 = tb[3], B1 = tb[4], B2 = tb[5];
  const float* z0 = logits + static_cast<int64_t>(n) * kNC * HW;
  float* d0 = dlogits + static_cast<int64_t>(n) * kNC * HW;
  const TT* tg = target + static_cast<int64_t>(n) * HW;
  const bool vec = (HW & 3) == 0;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kLossThreads * kLossPxPerThread;
  constexpr int kIters = kLossPxPerThread / 4;
  float a[kIters][3][4];
  int t[kIters][4];
  bool live[kIters];
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    const int64_t px = base + (static_cast<int64_t>(k) * kLossThreads + threadIdx.x) * 4;
    live[k] = px < HW;
    if (live[k]) load_px4<TT>(z0, tg, HW, px, vec, a[k], t[k]);
  }
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    if (!live[k]) continue;
    const int64_t px = base + (static_cast<int64_t>(k) * kLossThreads + threadIdx.x) * 4;
    float g[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tj = t[k][j];
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (tj != ignore_index) {
        const float a0 = a[k][0][j], a1 = a[k][1][j], a2 = a[k][2][j];
        const float m = fmaxf(a0, fmaxf(a1, a2));
        const float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = expf(a2 - m);
        const float inv = 1.f / (e0 + e1 + e2);
        const float p0 = e0 * inv, p1 = e1 * inv, p2 = e2 * inv;
        const float wt = tj == 0 ? wn0 : (tj == 1 ? wn1 : (tj == 2 ? wn2 : 0.f));
        // dDice/dp_c
        const float G0 = B0 + (tj == 0 ? A0 : 0.f), G1 = B1 + (tj == 1 ? A1 : 0.f), G2 = B2 + (tj == 2 ? A2 : 0.f);
        const float gp = G0 * p0 + G1 * p1 + G2 * p2;
        g0 = weight_ce * wt * (p0 - (tj == 0 ? 1.f : 0.f)) + weight_dice * p0 * (G0 - gp);
        g1 = weight_ce * wt * (p1 - (tj == 1 ? 1.f : 0.f)) + weight_dice * p1 * (G1 - gp);
        g2 = weight_ce * wt * (p2 - (tj == 2 ? 1.f : 0.f)) + weight_dice * p2 * (G2 - gp);
      }
      g[0][j] = g0 * gs;
      g[1][j] = g1 * gs;
      g[2][j] = g2 * gs;
    }
    if (vec) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        *reinterpret_cast<float4*>(d0 + c * HW + px) = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (px + j < HW) {
          d0[px + j] = g[0][j];
          d0[HW + px + j] = g[1][j];
          d0[2 * HW + px + j] = g[2][j];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------ head
constexpr int kHeadMaxC = 64;
constexpr int kHeadMaxK = 4;

// logits[k] = b[k] + sum_c W[k][c] * z[c].  C/8 threads per pixel: thread (px, c8) loads ONE 16-byte octet -- a warp
// reads 512 contiguous bytes per instruction (the one-thread-per-pixel version touched 32 half-used sectors per
// instruction: 3.7 TB/s) --, forms its partial dot products, a butterfly over the C/8 lanes completes them and lane
// c8 == k stores class k (8 consecutive pixels = one full sector per plane and warp).  kPix pixels per thread in flight.
// Optional fused producer (na, nb): z is then the RAW conv output of the last unit and the activation
// leaky_relu(na[n,c] * y + nb[n,c]) is applied on the fly (that unit's apply pass never runs; backward recomputes it).
template <typename T, int C, int K>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ z, int64_t zp,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ logits, int64_t HW,
                                                        const float* __restrict__ na, const float* __restrict__ nb,
                                                        float slope) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  constexpr int C8N = C / 8, L = 256 / C8N, kPix = 4;
  static_assert(C8N >= K && C8N <= 32 && (C8N & (C8N - 1)) == 0, "C/8 must be a power of two in [K, 32]");
  const int n = blockIdx.y;
  const int c8 = threadIdx.x % C8N, c0 = c8 * 8;
  float wr[K][8], av[8], bv[8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[k][i] = __ldg(w + k * C + c0 + i);
  if (na) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      av[i] = __ldg(na + n * C + c0 + i);
      bv[i] = __ldg(nb + n * C + c0 + i);
    }
  }
  const float my_bias = c8 < K ? __ldg(bias + c8) : 0.f;
  const T* zn = z + static_cast<int64_t>(n) * HW * zp + c0;
  float* ln = logits + static_cast<int64_t>(n) * K * HW;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * (L * kPix) + threadIdx.x / C8N;
  Vec8<T> v[kPix];
#pragma unroll
  for (int u = 0; u < kPix; ++u)
    if (base + u * L < HW) v[u] = Vec8<T>::ld_stream(zn + (base + u * L) * zp);
#pragma unroll
  for (int u = 0; u < kPix; ++u) {
    const int64_t px = base + u * L;
    if (px >= HW) break;
    float zf[8], acc[K];
    v[u].unpack(zf);
    if (na) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = fmaf(av[i], zf[i], bv[i]);
        zf[i] = t > 0.f ? t : t * slope;
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a = fmaf(wr[k][i], zf[i], a);
      acc[k] = a;
    }
#pragma unroll
    for (int m = 1; m < C8N; m <<= 1)
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], m);
    float mine = acc[0];
#pragma unroll
    for (int k = 1; k < K; ++k) mine = c8 == k ? acc[k] : mine;
    if (c8 < K) ln[c8 * HW + px] = mine + my_bias;
  }
}

// grid (blocks per image, N): thread (px, c8) computes dz for 8 channels of its pixels of image n and accumulates its 8
// columns of dW[K][C] (+ db on the c8 == 0 thread) -- 27 accumulators instead of K*C = 96 per thread (the
// one-thread-per-pixel version needed 254 registers: 12 % occupancy, 2.4 TB/s).  Block reduction: butterfly over the
// pixel lanes of a warp, then the 8 warps in fixed order; one partial row per block (deterministic).
// STATS (only with the fused producer na/nb): the block also emits the InstanceNorm-backward partial sums of the unit
// whose raw output y it is reading -- T1 = sum gm, T2raw = sum gm * y with gm = dz_stored * lrelu'(na*y+nb) -- so that
// that unit's norm backward needs no reduction pass over (dz, y) (b200unet_in_bwd_args.ext_part).
template <typename T, int C, int K, bool STATS>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 2 : 1) head_bwd_kernel(const float* __restrict__ dl, const T* __restrict__ z,
                                                        int64_t zp, const float* __restrict__ w,
                                                        T* __restrict__ dz, int64_t dzp,
                                                        float* __restrict__ partial, int64_t HW,
                                                        const float* __restrict__ na, const float* __restrict__ nb,
                                                        float slope, float* __restrict__ tpart) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  constexpr int C8N = C / 8;
  constexpr int L = 256 / C8N;  // pixel lanes of a block
  static_assert(C8N >= 1 && C8N <= 32 && (C8N & (C8N - 1)) == 0, "C/8 must be a power of two <= 32");
  __shared__ float red[8][K * C + K + (STATS ? 2 * C : 0)];
  const int n = blockIdx.y;
  const int c8 = threadIdx.x % C8N, c0 = c8 * 8;
  float wr[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[k][i] = w[k * C + c0 + i];
  float av[8], bv[8];
  if (na) {  // per-image affine of the fused producer: z = leaky_relu(na * y + nb)
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(na + n * C + c0)), a1 = __ldg(reinterpret_cast<const float4*>(na + n * C + c0) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(nb + n * C + c0)), b1 = __ldg(reinterpret_cast<const float4*>(nb + n * C + c0) + 1);
    av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
  }
  float aw[K][8], ab[K], t1[STATS ? 8 : 1], t2[STATS ? 8 : 1];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    ab[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[k][i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < (STATS ? 8 : 1); ++i) t1[i] = t2[i] = 0.f;
  const float* dln = dl + static_cast<int64_t>(n) * K * HW;
  const T* zn = z + static_cast<int64_t>(n) * HW * zp + c0;
  T* dzn = dz + static_cast<int64_t>(n) * HW * dzp + c0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L;
  constexpr int U = STATS ? 3 : 4;  // pixels in flight per thread (the sums cost 16 registers)
  for (int64_t g0 = static_cast<int64_t>(blockIdx.x) * L + threadIdx.x / C8N; g0 < HW; g0 += U * stride) {
    float d[U][K];
    Vec8<T> zv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t px = g0 + u * stride;
      if (px < HW) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[u][k] = __ldg(dln + k * HW + px);
        zv[u] = Vec8<T>::ld_stream(zn + px * zp);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t px = g0 + u * stride;
      if (px >= HW) break;
      float zf[8], o[8];
      zv[u].unpack(zf);
#pragma unroll
      for (int k = 0; k < K; ++k) ab[k] += d[u][k];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float yraw = zf[i];
        float zact = yraw, sl = 1.f;
        if (na) {
          const float t = fmaf(av[i], yraw, bv[i]);
          sl = t > 0.f ? 1.f : slope;
          zact = t * sl;
        }
        float sacc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          sacc = fmaf(wr[k][i], d[u][k], sacc);
          aw[k][i] = fmaf(d[u][k], zact, aw[k][i]);
        }
        o[i] = sacc;
        if (STATS) {
          const float gm = to_f32(from_f32<T>(sacc)) * sl;  // the value the norm backward will read back
          t1[i] += gm;
          t2[i] = fmaf(gm, yraw, t2[i]);
        }
      }
      Vec8<T>::st(dzn + px * dzp, o);
    }
  }
  // pixel lanes of a warp that share a channel octet: butterfly over the lane bits above log2(C8N)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int m = C8N; m < 32; m <<= 1) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) aw[k][i] += __shfl_xor_sync(0xffffffffu, aw[k][i], m);
      ab[k] += __shfl_xor_sync(0xffffffffu, ab[k], m);
    }
    if (STATS) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t1[i] += __shfl_xor_sync(0xffffffffu, t1[i], m);
        t2[i] += __shfl_xor_sync(0xffffffffu, t2[i], m);
      }
    }
  }
  if (lane < C8N) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][k * C + c0 + i] = aw[k][i];
      if (c8 == 0) red[warp][K * C + k] = ab[k];
    }
    if (STATS) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red[warp][K * C + K + 2 * (c0 + i)] = t1[i];
        red[warp][K * C + K + 2 * (c0 + i) + 1] = t2[i];
      }
    }
  }
  __syncthreads();
  const int64_t row = static_cast<int64_t>(n) * gridDim.x + blockIdx.x;
  for (int i = threadIdx.x; i < K * C + K + (STATS ? 2 * C : 0); i += 256) {
    float sacc = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sacc += red[wv][i];
    if (i < K * C + K) partial[row * (K * C + K) + i] = sacc;
    else tpart[row * (2 * C) + (i - (K * C + K))] = sacc;  // [N][P][C][2]
  }
}

// one block of 32 warps: warp w reduces values i = w, w + 32, ... over the `blocks` partial rows (lanes stride over the
// rows, fixed-order shuffle tree)
__global__ void __launch_bounds__(1024) head_bwd_finalize_kernel(const float* __restrict__ partial, int blocks, int KC,
                                                                  int K, float* __restrict__ dw, float* __restrict__ db) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < KC + K; i += 32) {
    double s = 0.0;
    for (int b = lane; b < blocks; b += 32) s += partial[static_cast<int64_t>(b) * (KC + K) + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      if (i < KC) dw[i] = static_cast<float>(s);
      else db[i - KC] = static_cast<float>(s);
    }
  }
}

// ------------------------------------------------------------------------------- reconstruction head + MSE (autoencoder)
// AE_pretrained/reconstruction/models/autoencoder.py:374-387: reconstruction_output = Conv2d(32 -> 3, 3x3, pad 1, bias)
// + Sigmoid.  The 3x3 conv itself runs on the conv kernels (output channels zero-padded to the kernels' granularity);
// these kernels are its epilogue -- bias + sigmoid into the fp32 NCHW boundary tensor -- and the matching backward
// prologue: dpre = dout * out * (1 - out) as the (padded) NHWC gradient operand of dgrad/wgrad, plus db = sum dpre.
template <typename T>
__global__ void __launch_bounds__(256) recon_head_fwd_kernel(const T* __restrict__ y, int64_t yp,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int64_t HW, int K) {
  const int n = blockIdx.y;
  const int64_t px = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (px >= HW) return;
  const T* src = y + (static_cast<int64_t>(n) * HW + px) * yp;
  for (int k = 0; k < K; ++k) {
    const float v = to_f32(src[k]) + bias[k];
    out[(static_cast<int64_t>(n) * K + k) * HW + px] = 1.f / (1.f + expf(-v));
  }
}

// grid (blocks, N); every block writes one partial row [K] of db.  One thread per pixel: K coalesced 4-byte reads per
// operand, and the Cpad-channel NHWC row (K real values, then zeros) written as 16-byte stores (the first version stored
// the row one element at a time: 32 two-byte stores per pixel, 4.4 ms at batch 64 against 0.3 ms of traffic).
template <typename T>
__global__ void __launch_bounds__(256) recon_head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                              T* __restrict__ dpre, int64_t dp, int Cpad,
                                                              float* __restrict__ partial, int64_t HW, int K) {
  __shared__ float scratch[8];
  const int n = blockIdx.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int vecs = Cpad * static_cast<int>(sizeof(T)) / 16;  // 16-byte stores per row (Cpad % 8 == 0)
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; px < HW; px += static_cast<int64_t>(gridDim.x) * 256) {
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < K) {
        const int64_t i = (static_cast<int64_t>(n) * K + k) * HW + px;
        const float o = __ldg(out + i);
        g[k] = __ldg(dout + i) * o * (1.f - o);
        acc[k] += g[k];
      }
    uint4* dst = reinterpret_cast<uint4*>(dpre + (static_cast<int64_t>(n) * HW + px) * dp);
    if (sizeof(T) == 2) {
      dst[0] = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), 0u, 0u);
    } else {
      dst[0] = make_uint4(__float_as_uint(g[0]), __float_as_uint(g[1]), __float_as_uint(g[2]), __float_as_uint(g[3]));
    }
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int v = 1; v < vecs; ++v) dst[v] = z;
  }
  for (int k = 0; k < K && k < 4; ++k) {
    const float v = block_sum(acc[k], scratch);
    if (threadIdx.x == 0) partial[(static_cast<int64_t>(n) * gridDim.x + blockIdx.x) * 4 + k] = v;
  }
}

__global__ void recon_head_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int K, float* __restrict__ db) {
  const int k = threadIdx.x;
  if (k >= K) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += partial[static_cast<int64_t>(r) * 4 + k];
  db[k] = static_cast<float>(s);
}

// nn.MSELoss(reduction='mean') (AE_pretrained/reconstruction/src/train.py:431): loss = mean((a - b)^2)
__global__ void __launch_bounds__(256) mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ partial, int64_t n) {
  __shared__ float scratch[8];
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float d = a[i] - b[i];
    acc = fmaf(d, d, acc);
  }
  const float v = block_sum(acc, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
__global__ void mse_finalize_kernel(const float* __restrict__ partial, int blocks, double inv_n, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < blocks; ++i) s += partial[i];
  out[0] = static_cast<float>(s * inv_n);
}
__global__ void __launch_bounds__(256) mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       const float* __restrict__ gout, float scale,
                                                       float* __restrict__ da, int64_t n) {
  const float g = (gout ? gout[0] : 1.f) * scale;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256)
    da[i] = g * (a[i] - b[i]);
}



}  // namespace b200

using namespace b200;

extern "C" int64_t b200unet_loss_workspace(int N, int64_t HW) {
  const int64_t blocks = ceil_div64(HW, kLossThreads * kLossPxPerThread);
  return static_cast<int64_t>(N) * blocks * kLossVals * 4;
}

template <typename TT>
static int loss_fwd_impl(const float* logits_nchw, const TT* target, const float* class_weights, int dynamic,
                         float weight_ce, float weight_dice, int ignore_index, float smooth, float* loss_out,
                         float* tables, float* workspace, int64_t workspace_bytes, int N, int64_t HW, void* stream) {
  B200_CHECK_ARG(logits_nchw && target && loss_out && tables && workspace, "loss_fwd: null pointer");
  B200_CHECK_ARG(N > 0 && HW > 0, "loss_fwd: empty batch");
  B200_CHECK_ARG(workspace_bytes >= b200unet_loss_workspace(N, HW), "loss_fwd: workspace too small");
  const int blocks = static_cast<int>(ceil_div64(HW, kLossThreads * kLossPxPerThread));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_k(loss_fwd_kernel<TT>, dim3(blocks, N), dim3(kLossThreads), 0, st, logits_nchw, target, ignore_index, workspace, HW);
  B200_LAUNCH_CHECK("loss_fwd_kernel");
  const size_t smem = static_cast<size_t>(N) * (kLossVals + kNC) * sizeof(double);
  B200_CHECK_ARG(N * kNC <= 1024, "loss_fwd: batch %d too large for the finalize kernel", N);
  B200_CHECK_ARG(smem <= 48 * 1024, "loss_fwd: batch %d too large for the finalize kernel", N);
  launch_k(loss_finalize_kernel, dim3(1), dim3(1024), smem, st, workspace, blocks, class_weights, dynamic, weight_ce, weight_dice, smooth,
                                             loss_out, tables, N);
  B200_LAUNCH_CHECK("loss_finalize_kernel");
  return 0;
}

template <typename TT>
static int loss_bwd_impl(const float* logits_nchw, const TT* target, const float* tables, const float* grad_out,
                         float weight_ce, float weight_dice, int ignore_index, float* dlogits_nchw, int N, int64_t HW,
                         void* stream) {
  B200_CHECK_ARG(logits_nchw && target && tables && dlogits_nchw, "loss_bwd: null pointer");
  const int blocks = static_cast<int>(ceil_div64(HW, kLossThreads * kLossPxPerThread));
  launch_k(loss_bwd_kernel<TT>, dim3(blocks, N), dim3(kLossThreads), 0, static_cast<cudaStream_t>(stream), 
      logits_nchw, target, tables, grad_out, weight_ce, weight_dice, ignore_index, dlogits_nchw, HW);
  B200_LAUNCH_CHECK("loss_bwd_kernel");
  return 0;
}

extern "C" int b200unet_loss_fwd(const float* logits_nchw, const int64_t* target, const float* class_weights,
                                 int dynamic, float weight_ce, float weight_dice, int ignore_index, float smooth,
                                 float* loss_out, float* tables, float* workspace, int64_t workspace_bytes, int N,
                                 int64_t HW, void* stream) {
  return loss_fwd_impl<int64_t>(logits_nchw, target, class_weights, dynamic, weight_ce, weight_dice, ignore_index, smooth,
                                loss_out, tables, workspace, workspace_bytes, N, HW, stream);
}
extern "C" int b200unet_loss_fwd_u8(const float* logits_nchw, const uint8_t* target, const float* class_weights,
                                    int dynamic, float weight_ce, float weight_dice, int ignore_index, float smooth,
                                    float* loss_out, float* tables, float* workspace, int64_t workspace_bytes, int N,
                                    int64_t HW, void* stream) {
  return loss_fwd_impl<uint8_t>(logits_nchw, target, class_weights, dynamic, weight_ce, weight_dice, ignore_index, smooth,
                                loss_out, tables, workspace, workspace_bytes, N, HW, stream);
}

extern "C" int b200unet_loss_bwd(const float* logits_nchw, const int64_t* target, const float* tables,
                                 const float* grad_out, float weight_ce, float weight_dice, int ignore_index,
                                 float* dlogits_nchw, int N, int64_t HW, void* stream) {
  return loss_bwd_impl<int64_t>(logits_nchw, target, tables, grad_out, weight_ce, weight_dice, ignore_index, dlogits_nchw,
                                N, HW, stream);
}
extern "C" int b200unet_loss_bwd_u8(const float* logits_nchw, const uint8_t* target, const float* tables,
                                    const float* grad_out, float weight_ce, float weight_dice, int ignore_index,
                                    float* dlogits_nchw, int N, int64_t HW, void* stream) {
  return loss_bwd_impl<uint8_t>(logits_nchw, target, tables, grad_out, weight_ce, weight_dice, ignore_index, dlogits_nchw,
                                N, HW, stream);
}

template <typename T>
static int head_fwd_impl(const void* z, int64_t z_pitch, const float* w, const float* bias, float* logits_nchw, int N,
                         int64_t HW, int C, int K, void* stream, const float* na = nullptr, const float* nb = nullptr,
                         float slope = 0.f) {
  B200_CHECK_ARG(z && w && bias && logits_nchw, "head_fwd: null pointer");
  B200_CHECK_ARG(z_pitch % 8 == 0, "head_fwd: pitch must be a multiple of 8");
  if (!(C == 32 && K == 3)) return set_error(kErrUnsupported, "head_fwd: only C=32, K=3 is built (got C=%d K=%d)", C, K);
  dim3 grid((unsigned)ceil_div64(HW, 256), N);  // 256 threads = 64 pixel lanes x 4 pixels per thread
  launch_k(head_fwd_kernel<T, 32, 3>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const T*>(z), z_pitch, w,
                                                                                  bias, logits_nchw, HW, na, nb, slope);
  B200_LAUNCH_CHECK("head_fwd_kernel");
  return 0;
}
extern "C" int b200unet_head_fwd(const void* z, int64_t z_pitch, const float* w, const float* bias, float* logits_nchw,
                                 int N, int64_t HW, int C, int K, void* stream) {
  return head_fwd_impl<__nv_bfloat16>(z, z_pitch, w, bias, logits_nchw, N, HW, C, K, stream);
}
extern "C" int b200unet_head_fwd_f32(const void* z, int64_t z_pitch, const float* w, const float* bias,
                                     float* logits_nchw, int N, int64_t HW, int C, int K, void* stream) {
  return head_fwd_impl<float>(z, z_pitch, w, bias, logits_nchw, N, HW, C, K, stream);
}

// blocks per image of the head backward = partial-sum slots per image of its norm-backward sums ([N][P][C][2])
extern "C" int b200unet_head_bwd_stat_slots(int N, int64_t HW) {
  if (N <= 0 || HW <= 0) return 0;
  // ONE wave: the kernel holds two blocks per SM (128 registers); round DOWN so that N * per <= the resident blocks
  // (19 blocks per image at batch 32 were 608 blocks for 296 places: a third, almost empty wave)
  int64_t per = (static_cast<int64_t>(num_sms()) * 2) / N;
  const int64_t mx = ceil_div64(HW, 64);  // at least one sweep of 64 pixel lanes per block
  if (per > mx) per = mx;
  return static_cast<int>(per < 1 ? 1 : per);
}

extern "C" int64_t b200unet_head_bwd_workspace(int N, int64_t HW, int C, int K) {
  return static_cast<int64_t>(b200unet_head_bwd_stat_slots(N, HW)) * N * (K * C + K) * 4;
}

template <typename T>
static int head_bwd_impl(const float* dlogits_nchw, const void* z, int64_t z_pitch, const float* w, void* dz,
                         int64_t dz_pitch, float* dw, float* db, float* workspace, int64_t workspace_bytes, int N,
                         int64_t HW, int C, int K, void* stream, const float* na = nullptr, const float* nb = nullptr,
                         float slope = 0.f, float* tpart = nullptr) {
  B200_CHECK_ARG(dlogits_nchw && z && w && dz && dw && db && workspace, "head_bwd: null pointer");
  B200_CHECK_ARG(z_pitch % 8 == 0 && dz_pitch % 8 == 0, "head_bwd: pitches must be multiples of 8");
  B200_CHECK_ARG(N > 0 && N <= 65535, "head_bwd: batch must fit the grid");
  if (!(C == 32 && K == 3)) return set_error(kErrUnsupported, "head_bwd: only C=32, K=3 is built (got C=%d K=%d)", C, K);
  B200_CHECK_ARG(!tpart || (na && nb), "head_bwd: the norm-backward partial sums need the fused producer (a, b)");
  const int bpi = b200unet_head_bwd_stat_slots(N, HW);
  B200_CHECK_ARG(workspace_bytes >= b200unet_head_bwd_workspace(N, HW, C, K), "head_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tpart)
    launch_k(head_bwd_kernel<T, 32, 3, true>, dim3(bpi, N), dim3(256), 0, st, dlogits_nchw, static_cast<const T*>(z), z_pitch, w,
                                                                  static_cast<T*>(dz), dz_pitch, workspace, HW, na, nb,
                                                                  slope, tpart);
  else
    launch_k(head_bwd_kernel<T, 32, 3, false>, dim3(bpi, N), dim3(256), 0, st, dlogits_nchw, static_cast<const T*>(z), z_pitch, w,
                                                                   static_cast<T*>(dz), dz_pitch, workspace, HW, na, nb,
                                                                   slope, nullptr);
  B200_LAUNCH_CHECK("head_bwd_kernel");
  launch_k(head_bwd_finalize_kernel, dim3(1), dim3(1024), 0, st, workspace, bpi * N, K * C, K, dw, db);
  B200_LAUNCH_CHECK("head_bwd_finalize_kernel");
  return 0;
}

extern "C" int b200unet_head_bwd(const float* dlogits_nchw, const void* z, int64_t z_pitch, const float* w, void* dz,
                                 int64_t dz_pitch, float* dw, float* db, float* workspace, int64_t workspace_bytes,
                                 int N, int64_t HW, int C, int K, void* stream) {
  return head_bwd_impl<__nv_bfloat16>(dlogits_nchw, z, z_pitch, w, dz, dz_pitch, dw, db, workspace, workspace_bytes, N,
                                      HW, C, K, stream);
}
extern "C" int b200unet_head_bwd_f32(const float* dlogits_nchw, const void* z, int64_t z_pitch, const float* w,
                                     void* dz, int64_t dz_pitch, float* dw, float* db, float* workspace,
                                     int64_t workspace_bytes, int N, int64_t HW, int C, int K, void* stream) {
  return head_bwd_impl<float>(dlogits_nchw, z, z_pitch, w, dz, dz_pitch, dw, db, workspace, workspace_bytes, N, HW, C,
                              K, stream);
}

static int recon_blocks(int64_t HW, int N) {
  int64_t per = ceil_div64(static_cast<int64_t>(num_sms()) * 8, N);
  const int64_t mx = ceil_div64(HW, 256);
  if (per > mx) per = mx;
  return static_cast<int>(per < 1 ? 1 : per);
}

template <typename T>
static int recon_head_fwd_impl(const void* y, int64_t y_pitch, const float* bias, float* out_nchw, int N, int64_t HW,
                               int K, void* stream) {
  B200_CHECK_ARG(y && bias && out_nchw, "recon_head_fwd: null pointer");
  B200_CHECK_ARG(K >= 1 && K <= 4 && y_pitch >= K && N <= 65535, "recon_head_fwd: K must be in [1,4]");
  recon_head_fwd_kernel<T><<<dim3((unsigned)ceil_div64(HW, 256), N), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(y), y_pitch, bias, out_nchw, HW, K);
  B200_LAUNCH_CHECK("recon_head_fwd_kernel");
  return 0;
}
extern "C" int b200unet_recon_head_fwd(const void* y, int64_t y_pitch, const float* bias, float* out_nchw, int N,
                                       int64_t HW, int K, void* stream) {
  return recon_head_fwd_impl<__nv_bfloat16>(y, y_pitch, bias, out_nchw, N, HW, K, stream);
}
extern "C" int b200unet_recon_head_fwd_f32(const void* y, int64_t y_pitch, const float* bias, float* out_nchw, int N,
                                           int64_t HW, int K, void* stream) {
  return recon_head_fwd_impl<float>(y, y_pitch, bias, out_nchw, N, HW, K, stream);
}

extern "C" int64_t b200unet_recon_head_bwd_workspace(int N, int64_t HW) {
  return static_cast<int64_t>(N) * recon_blocks(HW, N) * 4 * 4;
}

template <typename T>
static int recon_head_bwd_impl(const float* dout_nchw, const float* out_nchw, void* dpre, int64_t dpre_pitch, int Cpad,
                               float* db, float* workspace, int64_t workspace_bytes, int N, int64_t HW, int K,
                               void* stream) {
  B200_CHECK_ARG(dout_nchw && out_nchw && dpre && db && workspace, "recon_head_bwd: null pointer");
  B200_CHECK_ARG(K >= 1 && K <= 4 && Cpad >= K && dpre_pitch >= Cpad && N <= 65535, "recon_head_bwd: bad channel counts");
  B200_CHECK_ARG(Cpad % 8 == 0 && dpre_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(dpre) & 15) == 0,
                 "recon_head_bwd: Cpad and the pitch must be multiples of 8, dpre 16-byte aligned");
  B200_CHECK_ARG(workspace_bytes >= b200unet_recon_head_bwd_workspace(N, HW), "recon_head_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = recon_blocks(HW, N);
  recon_head_bwd_kernel<T><<<dim3(blocks, N), 256, 0, st>>>(dout_nchw, out_nchw, static_cast<T*>(dpre), dpre_pitch, Cpad,
                                                            workspace, HW, K);
  B200_LAUNCH_CHECK("recon_head_bwd_kernel");
  recon_head_bwd_finalize_kernel<<<1, 32, 0, st>>>(workspace, N * blocks, K, db);
  B200_LAUNCH_CHECK("recon_head_bwd_finalize_kernel");
  return 0;
}
extern "C" int b200unet_recon_head_bwd(const float* dout_nchw, const float* out_nchw, void* dpre, int64_t dpre_pitch,
                                       int Cpad, float* db, float* workspace, int64_t workspace_bytes, int N, int64_t HW,
                                       int K, void* stream) {
  return recon_head_bwd_impl<__nv_bfloat16>(dout_nchw, out_nchw, dpre, dpre_pitch, Cpad, db, workspace, workspace_bytes,
                                            N, HW, K, stream);
}
extern "C" int b200unet_recon_head_bwd_f32(const float* dout_nchw, const float* out_nchw, void* dpre, int64_t dpre_pitch,
                                           int Cpad, float* db, float* workspace, int64_t workspace_bytes, int N,
                                           int64_t HW, int K, void* stream) {
  return recon_head_bwd_impl<float>(dout_nchw, out_nchw, dpre, dpre_pitch, Cpad, db, workspace, workspace_bytes, N, HW,
                                    K, stream);
}

static int mse_blocks(int64_t n) {
  const int64_t mx = ceil_div64(n, 256 * 8);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  return static_cast<int>(mx < 1 ? 1 : (mx > cap ? cap : mx));
}
extern "C" int64_t b200unet_mse_workspace(int64_t n) { return static_cast<int64_t>(mse_blocks(n)) * 4; }

extern "C" int b200unet_mse_fwd(const float* a, const float* b, float* loss_out, float* workspace,
                                int64_t workspace_bytes, int64_t n, void* stream) {
  B200_CHECK_ARG(a && b && loss_out && workspace && n > 0, "mse_fwd: null pointer or empty input");
  B200_CHECK_ARG(workspace_bytes >= b200unet_mse_workspace(n), "mse_fwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = mse_blocks(n);
  mse_fwd_kernel<<<blocks, 256, 0, st>>>(a, b, workspace, n);
  B200_LAUNCH_CHECK("mse_fwd_kernel");
  mse_finalize_kernel<<<1, 32, 0, st>>>(workspace, blocks, 1.0 / static_cast<double>(n), loss_out);
  B200_LAUNCH_CHECK("mse_finalize_kernel");
  return 0;
}

extern "C" int b200unet_mse_bwd(const float* a, const float* b, const float* grad_out, float* da, int64_t n,
                                void* stream) {
  B200_CHECK_ARG(a && b && da && n > 0, "mse_bwd: null pointer or empty input");
  mse_bwd_kernel<<<mse_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, grad_out,
                                                                               static_cast<float>(2.0 / static_cast<double>(n)), da, n);
  B200_LAUNCH_CHECK("mse_bwd_kernel");
  return 0;
}

// Head with the last unit's InstanceNorm/LeakyReLU/dropout apply fused in: y is that unit's RAW conv output, (a, b) its
// folded per-(n, c) affine (b200unet_in_finalize), slope the LeakyReLU slope.
extern "C" int b200unet_head_norm_fwd(const void* y, int64_t y_pitch, const float* a, const float* b, float slope,
                                      const float* w, const float* bias, float* logits_nchw, int N, int64_t HW, int C,
                                      int K, void* stream) {
  B200_CHECK_ARG(a && b, "head_norm_fwd: null affine");
  return head_fwd_impl<__nv_bfloat16>(y, y_pitch, w, bias, logits_nchw, N, HW, C, K, stream, a, b, slope);
}
extern "C" int b200unet_head_norm_fwd_f32(const void* y, int64_t y_pitch, const float* a, const float* b, float slope,
                                          const float* w, const float* bias, float* logits_nchw, int N, int64_t HW,
                                          int C, int K, void* stream) {
  B200_CHECK_ARG(a && b, "head_norm_fwd: null affine");
  return head_fwd_impl<float>(y, y_pitch, w, bias, logits_nchw, N, HW, C, K, stream, a, b, slope);
}
extern "C" int b200unet_head_norm_bwd(const float* dlogits_nchw, const void* y, int64_t y_pitch, const float* a,
                                      const float* b, float slope, const float* w, void* dz, int64_t dz_pitch, float* dw,
                                      float* db, float* workspace, int64_t workspace_bytes, int N, int64_t HW, int C,
                                      int K, void* stream) {
  B200_CHECK_ARG(a && b, "head_norm_bwd: null affine");
  return head_bwd_impl<__nv_bfloat16>(dlogits_nchw, y, y_pitch, w, dz, dz_pitch, dw, db, workspace, workspace_bytes, N,
                                      HW, C, K, stream, a, b, slope);
}
extern "C" int b200unet_head_norm_bwd_stats(const float* dlogits_nchw, const void* y, int64_t y_pitch, const float* a,
                                            const float* b, float slope, const float* w, void* dz, int64_t dz_pitch,
                                            float* dw, float* db, float* workspace, int64_t workspace_bytes,
                                            float* bwd_part, int N, int64_t HW, int C, int K, void* stream) {
  B200_CHECK_ARG(a && b && bwd_part, "head_norm_bwd_stats: null affine or partial buffer");
  return head_bwd_impl<__nv_bfloat16>(dlogits_nchw, y, y_pitch, w, dz, dz_pitch, dw, db, workspace, workspace_bytes, N,
                                      HW, C, K, stream, a, b, slope, bwd_part);
}
extern "C" int b200unet_head_norm_bwd_stats_f32(const float* dlogits_nchw, const void* y, int64_t y_pitch, const float* a,
                                                const float* b, float slope, const float* w, void* dz, int64_t dz_pitch,
                                                float* dw, float* db, float* workspace, int64_t workspace_bytes,
                                                float* bwd_part, int N, int64_t HW, int C, int K, void* stream) {
  B200_CHECK_ARG(a && b && bwd_part, "head_norm_bwd_stats: null affine or partial buffer");
  return head_bwd_impl<float>(dlogits_nchw, y, y_pitch, w, dz, dz_pitch, dw, db, workspace, workspace_bytes, N, HW, C, K,
                              stream, a, b, slope, bwd_part);
}
extern "C" int b200unet_head_norm_bwd_f32(const float* dlogits_nchw, const void* y, int64_t y_pitch, const float* a,
                                          const float* b, float slope, const float* w, void* dz, int64_t dz_pitch,
                                          float* dw, float* db, float* workspace, int64_t workspace_bytes, int N,
                                          int64_t HW, int C, int K, void* stream) {
  B200_CHECK_ARG(a && b, "head_norm_bwd: null affine");
  return head_bwd_impl<float>(dlogits_nchw, y, y_pitch, w, dz, dz_pitch, dw, db, workspace, workspace_bytes, N, HW, C, K,
                              stream, a, b, slope);
}

