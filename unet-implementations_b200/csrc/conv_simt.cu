// CUDA-core convolutions: (1) the Cin = 3 stem of encoder stage 0, which is HBM-bound and too narrow for the
// tensor-core path, and (2) slow direct convolutions with the tensor-core kernels' exact argument structs,
// used by the tests as an on-device cross-check and for channel counts outside the tcgen05 envelope.
// Reference call sites: nn.Conv2d in ConvBlock, Our_UNet/models/unet.py:106-115.
#include "common.cuh"
#include "ptx.cuh"
#include "vec8.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------ weight packing
template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout,
                                    int Cin) {
  const int64_t total = static_cast<int64_t>(Cout) * Cin * 9;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i enumerates the fprop pack [co][tap][ci] so its writes are coalesced
  const int ci = static_cast<int>(i % Cin);
  const int64_t r = i / Cin;
  const int tap = static_cast<int>(r % 9);
  const int co = static_cast<int>(r / 9);
  const float v = w[(static_cast<int64_t>(co) * Cin + ci) * 9 + tap];
  const T b = from_f32<T>(v);
  wf[i] = b;
  if (wd) wd[(static_cast<int64_t>(ci) * 9 + tap) * Cout + co] = b;
}

// ------------------------------------------------------------------------------------------ generic stats
// partial (sum, sumsq) of a bf16 NHWC tensor: block (p, n) covers pixels [p*chunk, (p+1)*chunk) of image n
template <typename T>
__global__ void stats_partial_kernel(const T* __restrict__ y, int64_t pitch, float* __restrict__ stats,
                                     int P, int64_t HW, int C, int64_t chunk) {
  const int p = blockIdx.x, n = blockIdx.y;
  const int64_t lo = p * chunk;
  int64_t hi = lo + chunk;
  if (hi > HW) hi = HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    typename AccT<T>::type s1 = 0, s2 = 0;
    for (int64_t px = lo; px < hi; ++px) {
      const typename AccT<T>::type v = to_f32(y[(n * HW + px) * pitch + c]);
      s1 += v;
      s2 += v * v;
    }
    float* d = stats + ((static_cast<int64_t>(n) * P + p) * C + c) * 2;
    d[0] = static_cast<float>(s1);
    d[1] = static_cast<float>(s2);
  }
}

// ------------------------------------------------------------------------------------------ direct convs
template <typename T>
__global__ void conv_fprop_simt_kernel(const T* __restrict__ x, int64_t xp, const T* __restrict__ w,
                                       T* __restrict__ y, int64_t yp,
                                       int N, int H, int W, int Cin, int Cout, int s, int OH, int OW) {
  const int64_t total = static_cast<int64_t>(N) * OH * OW * Cout;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % Cout);
  int64_t r = i / Cout;
  const int ow = static_cast<int>(r % OW);
  r /= OW;
  const int oh = static_cast<int>(r % OH);
  const int n = static_cast<int>(r / OH);
  using Acc = typename AccT<T>::type;
  Acc acc = 0;
  for (int kh = 0; kh < 3; ++kh) {
    const int ih = oh * s + kh - 1;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int iw = ow * s + kw - 1;
      if (iw < 0 || iw >= W) continue;
      const T* xr = x + ((static_cast<int64_t>(n) * H + ih) * W + iw) * xp;
      const T* wr = w + (static_cast<int64_t>(co) * 9 + kh * 3 + kw) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc += static_cast<Acc>(to_f32(xr[ci])) * static_cast<Acc>(to_f32(wr[ci]));
    }
  }
  y[((static_cast<int64_t>(n) * OH + oh) * OW + ow) * yp + co] = from_f32<T>(static_cast<float>(acc));
}

template <typename T>
__global__ void conv_dgrad_simt_kernel(const T* __restrict__ dy, int64_t dyp, const T* __restrict__ wt,
                                       T* __restrict__ dx, int64_t dxp, int N, int H, int W, int Cin, int Cout, int s,
                                       int OH, int OW) {
  const int64_t total = static_cast<int64_t>(N) * H * W * Cin;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = static_cast<int>(i % Cin);
  int64_t r = i / Cin;
  const int iw = static_cast<int>(r % W);
  r /= W;
  const int ih = static_cast<int>(r % H);
  const int n = static_cast<int>(r / H);
  using Acc = typename AccT<T>::type;
  Acc acc = 0;
  for (int kh = 0; kh < 3; ++kh) {
    const int th = ih + 1 - kh;
    if (th < 0 || th % s != 0) continue;
    const int oh = th / s;
    if (oh >= OH) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int tw = iw + 1 - kw;
      if (tw < 0 || tw % s != 0) continue;
      const int ow = tw / s;
      if (ow >= OW) continue;
      const T* dr = dy + ((static_cast<int64_t>(n) * OH + oh) * OW + ow) * dyp;
      const T* wr = wt + (static_cast<int64_t>(ci) * 9 + kh * 3 + kw) * Cout;
      for (int co = 0; co < Cout; ++co) acc += static_cast<Acc>(to_f32(dr[co])) * static_cast<Acc>(to_f32(wr[co]));
    }
  }
  dx[((static_cast<int64_t>(n) * H + ih) * W + iw) * dxp + ci] = from_f32<T>(static_cast<float>(acc));
}

template <typename T>
__global__ void conv_wgrad_simt_kernel(const T* __restrict__ x, int64_t xp, const T* __restrict__ dy, int64_t dyp,
                                       float* __restrict__ dw, int N, int H, int W, int Cin, int Cout, int s, int OH,
                                       int OW) {
  // one warp per (co, ci, tap); lanes stride over pixels, shuffle-reduce
  const int64_t gw = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(Cout) * Cin * 9;
  if (gw >= total) return;
  const int tap = static_cast<int>(gw % 9);
  const int64_t r = gw / 9;
  const int ci = static_cast<int>(r % Cin);
  const int co = static_cast<int>(r / Cin);
  const int kh = tap / 3, kw = tap % 3;
  using Acc = typename AccT<T>::type;
  Acc acc = 0;
  const int64_t npx = static_cast<int64_t>(N) * OH * OW;
  for (int64_t px = lane; px < npx; px += 32) {
    const int ow = static_cast<int>(px % OW);
    const int64_t q = px / OW;
    const int oh = static_cast<int>(q % OH);
    const int n = static_cast<int>(q / OH);
    const int ih = oh * s + kh - 1, iw = ow * s + kw - 1;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
    acc += static_cast<Acc>(to_f32(dy[px * dyp + co])) *
           static_cast<Acc>(to_f32(x[((static_cast<int64_t>(n) * H + ih) * W + iw) * xp + ci]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) dw[gw] = static_cast<float>(acc);  // gw == (co*Cin + ci)*9 + tap, the OIHW offset
}

// ------------------------------------------------------------------------------------------ stem (Cin=3 -> 32)
// K = 27 is too small for a UMMA tile and the fp32 NCHW image would need a layout pass first, so the stem runs on the
// CUDA cores -- with the packed FFMA2 (two fp32 FMAs per instruction, sm_100) and two pixels per thread so that every
// shared-memory weight fetch feeds four FMAs.  HBM: reads the image once, writes 64 B per pixel.
constexpr int kStemCo = 32;
constexpr int kStemThreads = 128;
constexpr int kStemPx = 2;  // adjacent pixels (along w) per thread

__global__ void __launch_bounds__(kStemThreads) stem_fprop_kernel(const float* __restrict__ img,
                                                                   const float* __restrict__ w_oihw,
                                                                   __nv_bfloat16* __restrict__ y, int64_t yp,
                                                                   float* __restrict__ stats, int P, int H, int W) {
  __shared__ __align__(16) float wsm[27][kStemCo];  // [ci*9 + kh*3 + kw][co]
  __shared__ float red[kStemThreads / 32][kStemCo][2];
  const int n = blockIdx.z;
  const int h = blockIdx.y;
  const int64_t HW = static_cast<int64_t>(H) * W;
  for (int i = threadIdx.x; i < 27 * kStemCo; i += kStemThreads) {
    const int co = i % kStemCo, k = i / kStemCo;
    wsm[k][co] = w_oihw[co * 27 + k];
  }
  __syncthreads();
  const int w0 = (blockIdx.x * kStemThreads + threadIdx.x) * kStemPx;
  float2 acc[kStemPx][kStemCo / 2];
#pragma unroll
  for (int q = 0; q < kStemPx; ++q)
#pragma unroll
    for (int c = 0; c < kStemCo / 2; ++c) acc[q][c] = make_float2(0.f, 0.f);
  if (w0 < W) {
    const float* base = img + static_cast<int64_t>(n) * 3 * HW;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = h + kh - 1;
        const bool hin = ih >= 0 && ih < H;
        const float* rowp = base + ci * HW + static_cast<int64_t>(ih) * W;
        // the kStemPx + 2 input columns this thread needs
        float v[kStemPx + 2];
#pragma unroll
        for (int j = 0; j < kStemPx + 2; ++j) {
          const int iw = w0 + j - 1;
          v[j] = (hin && iw >= 0 && iw < W) ? __ldg(rowp + iw) : 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4* wr = reinterpret_cast<const float4*>(wsm[ci * 9 + kh * 3 + kw]);
#pragma unroll
          for (int c4 = 0; c4 < kStemCo / 4; ++c4) {
            const float4 ww = wr[c4];
            const float2 wa = make_float2(ww.x, ww.y), wb = make_float2(ww.z, ww.w);
#pragma unroll
            for (int q = 0; q < kStemPx; ++q) {
              const float2 vv = make_float2(v[q + kw], v[q + kw]);
              acc[q][2 * c4] = __ffma2_rn(vv, wa, acc[q][2 * c4]);
              acc[q][2 * c4 + 1] = __ffma2_rn(vv, wb, acc[q][2 * c4 + 1]);
            }
          }
        }
      }
#pragma unroll
    for (int q = 0; q < kStemPx; ++q) {
      if (w0 + q < W) {
        uint4* dst = reinterpret_cast<uint4*>(y + (static_cast<int64_t>(n) * HW + static_cast<int64_t>(h) * W + w0 + q) * yp);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack_bf16x2(acc[q][4 * j].x, acc[q][4 * j].y), pack_bf16x2(acc[q][4 * j + 1].x, acc[q][4 * j + 1].y),
                              pack_bf16x2(acc[q][4 * j + 2].x, acc[q][4 * j + 2].y),
                              pack_bf16x2(acc[q][4 * j + 3].x, acc[q][4 * j + 3].y));
      }
    }
  }
  if (stats) {
    // sums of the values as stored (bf16-rounded) over this block's pixels: per-lane over its pixels, then across lanes
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float f[32], g[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) f[c] = g[c] = 0.f;
#pragma unroll
    for (int q = 0; q < kStemPx; ++q) {
      const bool valid = (w0 + q) < W;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float a0 = valid ? bf16_round(acc[q][c].x) : 0.f, a1 = valid ? bf16_round(acc[q][c].y) : 0.f;
        f[2 * c] += a0;
        g[2 * c] = fmaf(a0, a0, g[2 * c]);
        f[2 * c + 1] += a1;
        g[2 * c + 1] = fmaf(a1, a1, g[2 * c + 1]);
      }
    }
    red[warp][lane][0] = warp_colsum32(f, lane);
    red[warp][lane][1] = warp_colsum32(g, lane);
    __syncthreads();
    if (threadIdx.x < kStemCo * 2) {
      const int c = threadIdx.x >> 1, k = threadIdx.x & 1;
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < kStemThreads / 32; ++wv) s += red[wv][c][k];
      const int p = blockIdx.y * gridDim.x + blockIdx.x;
      stats[((static_cast<int64_t>(n) * P + p) * kStemCo + c) * 2 + k] = s;
    }
  }
}

// dW[co][ci][kh][kw] = sum dy[n,h,w,co] * img[n,ci,h+kh-1,w+kw-1].  A warp takes 2 pixels per step: lane = (pixel
// sub-index, channel pair); the 3-row fp32 input patch of a row segment is staged in shared memory; 27 float2
// accumulators per lane updated with packed FFMA2 (one instruction = both channels of the pair).
constexpr int kSegW = 256;
__global__ void __launch_bounds__(256, 2) stem_wgrad_kernel(const float* __restrict__ img,
                                                             const __nv_bfloat16* __restrict__ dy, int64_t dyp,
                                                             float* __restrict__ partial, int N, int H, int W) {
  __shared__ float patch[3][3][kSegW + 4];
  __shared__ float red[8][kStemCo][27];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ps = lane >> 4, cp = lane & 15;  // pixel sub-index 0..1, channels 2*cp, 2*cp+1
  const int segs_w = (W + kSegW - 1) / kSegW;
  const int64_t total_segs = static_cast<int64_t>(N) * H * segs_w;
  const int64_t HW = static_cast<int64_t>(H) * W;
  float2 acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = make_float2(0.f, 0.f);
  for (int64_t seg = blockIdx.x; seg < total_segs; seg += gridDim.x) {
    const int sw = static_cast<int>(seg % segs_w);
    const int64_t r = seg / segs_w;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const int w0 = sw * kSegW;
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * (kSegW + 2); i += 256) {
      const int col = i % (kSegW + 2);
      const int rr = i / (kSegW + 2);  // ci*3 + kh
      const int ci = rr / 3, kh = rr % 3;
      const int ih = h + kh - 1, iw = w0 + col - 1;
      float v = 0.f;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = img[(static_cast<int64_t>(n) * 3 + ci) * HW + static_cast<int64_t>(ih) * W + iw];
      patch[ci][kh][col] = v;
    }
    __syncthreads();
    const int wend = min(kSegW, W - w0);
    const __nv_bfloat16* drow = dy + (static_cast<int64_t>(n) * HW + static_cast<int64_t>(h) * W + w0) * dyp + 2 * cp;
#pragma unroll 2
    for (int j0 = warp * 2; j0 < wend; j0 += 16) {
      const int j = j0 + ps;
      float2 d = make_float2(0.f, 0.f);
      if (j < wend) {
        const uint32_t raw = *reinterpret_cast<const uint32_t*>(drow + static_cast<int64_t>(j) * dyp);
        d = make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
      }
      const int jj = j < wend ? j : 0;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float p0 = patch[ci][kh][jj], p1 = patch[ci][kh][jj + 1], p2 = patch[ci][kh][jj + 2];
          const int k = ci * 9 + kh * 3;
          acc[k] = __ffma2_rn(d, make_float2(p0, p0), acc[k]);
          acc[k + 1] = __ffma2_rn(d, make_float2(p1, p1), acc[k + 1]);
          acc[k + 2] = __ffma2_rn(d, make_float2(p2, p2), acc[k + 2]);
        }
    }
  }
  // combine the two pixel sub-groups (lanes cp and cp+16), then the warps
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float vx = acc[k].x + __shfl_xor_sync(0xffffffffu, acc[k].x, 16);
    const float vy = acc[k].y + __shfl_xor_sync(0xffffffffu, acc[k].y, 16);
    if (ps == 0) {
      red[warp][2 * cp][k] = vx;
      red[warp][2 * cp + 1][k] = vy;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kStemCo * 27; i += 256) {
    const int co = i / 27, k = i % 27;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][co][k];
    partial[static_cast<int64_t>(blockIdx.x) * (kStemCo * 27) + i] = s;
  }
}

__global__ void stem_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int blocks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kStemCo * 27) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += partial[static_cast<int64_t>(b) * (kStemCo * 27) + i];
  dw[i] = static_cast<float>(s);
}

static int stem_wgrad_blocks() { return num_sms() * 4; }

}  // namespace b200

using namespace b200;

template <typename T>
static int pack_weights_impl(const float* w_oihw, void* w_fprop, void* w_dgrad, int Cout, int Cin, void* stream) {
  B200_CHECK_ARG(w_oihw && w_fprop, "pack_conv_weights: null pointer");
  const int64_t total = static_cast<int64_t>(Cout) * Cin * 9;
  pack_weights_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, static_cast<T*>(w_fprop), static_cast<T*>(w_dgrad), Cout, Cin);
  B200_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}
extern "C" int b200unet_pack_conv_weights(const float* w_oihw, void* w_fprop, void* w_dgrad, int Cout, int Cin,
                                          void* stream) {
  return pack_weights_impl<__nv_bfloat16>(w_oihw, w_fprop, w_dgrad, Cout, Cin, stream);
}
extern "C" int b200unet_pack_conv_weights_f32(const float* w_oihw, void* w_fprop, void* w_dgrad, int Cout, int Cin,
                                              void* stream) {
  return pack_weights_impl<float>(w_oihw, w_fprop, w_dgrad, Cout, Cin, stream);
}

extern "C" int b200unet_conv_fprop_simt_partials(int OH, int OW) {
  int64_t p = (static_cast<int64_t>(OH) * OW) / 1024;
  return static_cast<int>(p < 1 ? 1 : (p > 64 ? 64 : p));
}

template <typename T>
static int conv_fprop_simt_impl(const b200unet_conv_fprop_args* a, void* stream) {
  B200_CHECK_ARG(a && a->x && a->w && a->y, "conv_fprop_simt: null pointer");
  B200_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv_fprop_simt: stride %d unsupported", a->stride);
  const int s = a->stride, OH = (a->H - 1) / s + 1, OW = (a->W - 1) / s + 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(a->N) * OH * OW * a->Cout;
  conv_fprop_simt_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(
      static_cast<const T*>(a->x), a->x_pitch, static_cast<const T*>(a->w), static_cast<T*>(a->y), a->y_pitch, a->N, a->H,
      a->W, a->Cin, a->Cout, s, OH, OW);
  B200_LAUNCH_CHECK("conv_fprop_simt_kernel");
  if (a->stats) {
    const int P = b200unet_conv_fprop_simt_partials(OH, OW);
    const int64_t HW = static_cast<int64_t>(OH) * OW;
    stats_partial_kernel<T><<<dim3(P, a->N), 128, 0, st>>>(static_cast<const T*>(a->y), a->y_pitch, a->stats, P, HW,
                                                           a->Cout, ceil_div64(HW, P));
    B200_LAUNCH_CHECK("stats_partial_kernel");
  }
  return 0;
}
extern "C" int b200unet_conv_fprop_simt(const b200unet_conv_fprop_args* a, void* stream) {
  return conv_fprop_simt_impl<__nv_bfloat16>(a, stream);
}
extern "C" int b200unet_conv_fprop_f32(const b200unet_conv_fprop_args* a, void* stream) {
  return conv_fprop_simt_impl<float>(a, stream);
}

template <typename T>
static int conv_dgrad_simt_impl(const b200unet_conv_dgrad_args* a, void* stream) {
  B200_CHECK_ARG(a && a->dy && a->wt && a->dx, "conv_dgrad_simt: null pointer");
  B200_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv_dgrad_simt: stride %d unsupported", a->stride);
  const int s = a->stride, OH = (a->H - 1) / s + 1, OW = (a->W - 1) / s + 1;
  const int64_t total = static_cast<int64_t>(a->N) * a->H * a->W * a->Cin;
  conv_dgrad_simt_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(a->dy), a->dy_pitch, static_cast<const T*>(a->wt), static_cast<T*>(a->dx), a->dx_pitch, a->N,
      a->H, a->W, a->Cin, a->Cout, s, OH, OW);
  B200_LAUNCH_CHECK("conv_dgrad_simt_kernel");
  return 0;
}
extern "C" int b200unet_conv_dgrad_simt(const b200unet_conv_dgrad_args* a, void* stream) {
  return conv_dgrad_simt_impl<__nv_bfloat16>(a, stream);
}
extern "C" int b200unet_conv_dgrad_f32(const b200unet_conv_dgrad_args* a, void* stream) {
  return conv_dgrad_simt_impl<float>(a, stream);
}

template <typename T>
static int conv_wgrad_simt_impl(const b200unet_conv_wgrad_args* a, void* stream) {
  B200_CHECK_ARG(a && a->x && a->dy && a->dw, "conv_wgrad_simt: null pointer");
  B200_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv_wgrad_simt: stride %d unsupported", a->stride);
  const int s = a->stride, OH = (a->H - 1) / s + 1, OW = (a->W - 1) / s + 1;
  const int64_t warps = static_cast<int64_t>(a->Cout) * a->Cin * 9;
  conv_wgrad_simt_kernel<T><<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(a->x), a->x_pitch, static_cast<const T*>(a->dy), a->dy_pitch, a->dw, a->N, a->H, a->W, a->Cin,
      a->Cout, s, OH, OW);
  B200_LAUNCH_CHECK("conv_wgrad_simt_kernel");
  return 0;
}
extern "C" int b200unet_conv_wgrad_simt(const b200unet_conv_wgrad_args* a, void* stream) {
  return conv_wgrad_simt_impl<__nv_bfloat16>(a, stream);
}
extern "C" int b200unet_conv_wgrad_f32(const b200unet_conv_wgrad_args* a, void* stream) {
  return conv_wgrad_simt_impl<float>(a, stream);
}

static int stem_blocks_w(int W) { return ceil_div(W, kStemThreads * kStemPx); }

extern "C" int b200unet_stem_partials(int H, int W) { return H * stem_blocks_w(W); }

extern "C" int b200unet_stem_fprop(const float* img_nchw, const float* w_oihw, void* y, int64_t y_pitch, float* stats,
                                   int N, int H, int W, void* stream) {
  B200_CHECK_ARG(img_nchw && w_oihw && y, "stem_fprop: null pointer");
  B200_CHECK_ARG(y_pitch % 8 == 0 && y_pitch >= kStemCo, "stem_fprop: bad output pitch");
  const int P = b200unet_stem_partials(H, W);
  B200_CHECK_ARG(H <= 65535 && N <= 65535, "stem_fprop: H and N must fit the grid");
  stem_fprop_kernel<<<dim3(stem_blocks_w(W), H, N), kStemThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      img_nchw, w_oihw, static_cast<__nv_bfloat16*>(y), y_pitch, stats, P, H, W);
  B200_LAUNCH_CHECK("stem_fprop_kernel");
  return 0;
}

extern "C" int64_t b200unet_stem_wgrad_workspace(int N, int H, int W) {
  (void)N; (void)H; (void)W;
  return static_cast<int64_t>(stem_wgrad_blocks()) * kStemCo * 27 * 4;
}

extern "C" int b200unet_stem_wgrad(const float* img_nchw, const void* dy, int64_t dy_pitch, float* dw_oihw,
                                   float* workspace, int64_t workspace_bytes, int N, int H, int W, void* stream) {
  B200_CHECK_ARG(img_nchw && dy && dw_oihw && workspace, "stem_wgrad: null pointer");
  const int blocks = stem_wgrad_blocks();
  B200_CHECK_ARG(workspace_bytes >= static_cast<int64_t>(blocks) * kStemCo * 27 * 4, "stem_wgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  stem_wgrad_kernel<<<blocks, 256, 0, st>>>(img_nchw, static_cast<const __nv_bfloat16*>(dy), dy_pitch, workspace, N, H, W);
  B200_LAUNCH_CHECK("stem_wgrad_kernel");
  stem_wgrad_finalize_kernel<<<ceil_div(kStemCo * 27, 256), 256, 0, st>>>(workspace, dw_oihw, blocks);
  B200_LAUNCH_CHECK("stem_wgrad_finalize_kernel");
  return 0;
}
