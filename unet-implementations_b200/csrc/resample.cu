// Exact 2x bilinear upsampling (align_corners=False) written straight into the decoder concat buffer, and its
// backward in gather form.  Reference: F.interpolate in UpBlock.forward (Our_UNet/models/unet.py:219-225) followed by
// torch.cat([x, skip], dim=1) (unet.py:228) -- the cat never runs here because both producers write their channel
// slice of one NHWC buffer.  Also the NCHW<->NHWC converters used by the per-module entry points.
//
// Per axis (n = input length):  out[2i]   = .25*in[max(i-1,0)] + .75*in[i]
//                               out[2i+1] = .75*in[i] + .25*in[min(i+1,n-1)]
// Backward (gather):            din[i] = .75*(d[2i] + d[2i+1]) + .25*(d[clamp(2i-1)] + d[clamp(2i+2)])
// where an index that falls off the output is redirected to the edge sample (that is where the clamped tap went).
#include "common.cuh"
#include "ptx.cuh"
#include "vec8.cuh"

namespace b200 {

// Forward: one thread = 8 channels of one INPUT COLUMN, walking R consecutive input rows.  Per input row it loads the
// three horizontal neighbours once (3 loads per input pixel instead of 9, and the fused activation below runs 3x
// instead of 9x per element), interpolates along w into the two output columns (L, R), and each output row pair is
// .75 * this row + .25 * the row above / below, kept in registers from the previous / next step -- interpolation along
// w first, then along h (the association of ATen's upsample_bilinear2d).  The loop is fully unrolled so the loads of
// the following rows are in flight while a row pair is stored.  grid (ceil(W*C/8 / 256), ceil(H/R), N): no integer
// division by runtime values.
// Optional fused producer: with (na, nb) the input is a RAW conv output y and every loaded value becomes
// leaky_relu(na[n,c] * y + nb[n,c]) first -- the InstanceNorm/LeakyReLU/dropout apply pass of the layer that feeds the
// upsample (its only consumer), so the activated low-resolution tensor is never written or re-read.
template <typename T, int R>
__global__ void __launch_bounds__(256, 2) upsample2x_fwd_kernel(const T* __restrict__ x, int64_t xp,
                                                                 T* __restrict__ out, int64_t op, int H, int W,
                                                                 int c8n, int c8shift, const float* __restrict__ na,
                                                                 const float* __restrict__ nb, float slope) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int iw = t >> c8shift;  // c8n is a power of two on this path
  if (iw >= W) return;
  const int c0 = (t & (c8n - 1)) << 3;
  const int ih0 = blockIdx.y * R, n = blockIdx.z;
  const int wm = iw > 0 ? iw - 1 : 0, wp = iw < W - 1 ? iw + 1 : W - 1;
  const T* b = x + static_cast<int64_t>(n) * H * W * xp + c0;
  const int64_t om = static_cast<int64_t>(wm) * xp, oc = static_cast<int64_t>(iw) * xp, od = static_cast<int64_t>(wp) * xp;
  float ra[8], rb[8];
  if (na) {
    const int C = c8n << 3;
    const float4* pa = reinterpret_cast<const float4*>(na + n * C + c0);
    const float4* pb = reinterpret_cast<const float4*>(nb + n * C + c0);
    const float4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
    ra[0] = a0.x; ra[1] = a0.y; ra[2] = a0.z; ra[3] = a0.w; ra[4] = a1.x; ra[5] = a1.y; ra[6] = a1.z; ra[7] = a1.w;
    rb[0] = b0.x; rb[1] = b0.y; rb[2] = b0.z; rb[3] = b0.w; rb[4] = b1.x; rb[5] = b1.y; rb[6] = b1.z; rb[7] = b1.w;
  }
  // horizontal pass of input row `row`: the two output columns 2iw (Lo) and 2iw + 1 (Ro)
  auto hmix = [&](int row, float (&Lo)[8], float (&Ro)[8]) {
    const T* rp = b + static_cast<int64_t>(row) * W * xp;
    float a[8], c[8], d[8];
    Vec8<T>::ldg(rp + om).unpack(a);
    Vec8<T>::ldg(rp + oc).unpack(c);
    Vec8<T>::ldg(rp + od).unpack(d);
    if (na) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t0 = fmaf(ra[j], a[j], rb[j]), t1 = fmaf(ra[j], c[j], rb[j]), t2 = fmaf(ra[j], d[j], rb[j]);
        a[j] = t0 > 0.f ? t0 : t0 * slope;
        c[j] = t1 > 0.f ? t1 : t1 * slope;
        d[j] = t2 > 0.f ? t2 : t2 * slope;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      Lo[j] = 0.75f * c[j] + 0.25f * a[j];
      Ro[j] = 0.75f * c[j] + 0.25f * d[j];
    }
  };
  float L[3][8], Rr[3][8];  // rotating: row above, this row, row below
  hmix(ih0 > 0 ? ih0 - 1 : 0, L[0], Rr[0]);
  hmix(ih0, L[1], Rr[1]);
  const int OW = 2 * W;
  T* o = out + (static_cast<int64_t>(n) * 2 * H + 2 * ih0) * OW * op + static_cast<int64_t>(2 * iw) * op + c0;
  const int64_t orow = static_cast<int64_t>(OW) * op;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int ih = ih0 + r;
    if (ih >= H) break;
    float (&Lp)[8] = L[r % 3], (&Lc)[8] = L[(r + 1) % 3], (&Ln)[8] = L[(r + 2) % 3];
    float (&Rp)[8] = Rr[r % 3], (&Rc)[8] = Rr[(r + 1) % 3], (&Rn)[8] = Rr[(r + 2) % 3];
    hmix(ih < H - 1 ? ih + 1 : H - 1, Ln, Rn);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.75f * Lc[j] + 0.25f * Lp[j];
    Vec8<T>::st(o, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.75f * Rc[j] + 0.25f * Rp[j];
    Vec8<T>::st(o + op, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.75f * Lc[j] + 0.25f * Ln[j];
    Vec8<T>::st(o + orow, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.75f * Rc[j] + 0.25f * Rn[j];
    Vec8<T>::st(o + orow + op, v);
    o += 2 * orow;
  }
}

// Backward (gather), separable: din[ih, iw] = sum_a wt[a] * hsum(row 2ih-1+a),  hsum(r) = sum_b wt[b] * d[r, 2iw-1+b],
// wt = (.25, .75, .75, .25), indices clamped (a clamped index is where the reference's edge-replicated tap landed, so its
// weight stays).  One thread = 8 channels of one input COLUMN and kUpRows consecutive input rows: it walks the
// 2*kUpRows + 2 output rows once, forms each row's horizontal sum (4 loads) and adds it into the (at most two) input
// rows that use it -- 4 + 4/kUpRows loads per input pixel instead of 16.
constexpr int kUpRows = 4;

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const T* __restrict__ dout, int64_t dp,
                                                              T* __restrict__ dx, int64_t xp, int H, int W,
                                                              int c8n, int c8shift) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int iw = t >> c8shift;
  if (iw >= W) return;
  const int c0 = (t & (c8n - 1)) << 3;
  const int ih0 = blockIdx.y * kUpRows, n = blockIdx.z;
  const int OH = 2 * H, OW = 2 * W;
  const T* b = dout + static_cast<int64_t>(n) * OH * OW * dp + c0;
  int cols[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int ow = 2 * iw - 1 + a;
    cols[a] = ow < 0 ? 0 : (ow >= OW ? OW - 1 : ow);
  }
  float acc[kUpRows][8];
#pragma unroll
  for (int r = 0; r < kUpRows; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
  // output rows 2*ih0 - 1 .. 2*ih0 + 2*kUpRows (unclamped index k = 0 .. 2*kUpRows + 1); input row ih0 + r uses
  // k = 2r .. 2r + 3 with weights (.25, .75, .75, .25)
#pragma unroll
  for (int k = 0; k < 2 * kUpRows + 2; ++k) {
    const int oh_u = 2 * ih0 - 1 + k;
    const int oh = oh_u < 0 ? 0 : (oh_u >= OH ? OH - 1 : oh_u);
    const T* rp = b + static_cast<int64_t>(oh) * OW * dp;
    Vec8<T> raw[4];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) raw[bb] = Vec8<T>::ldg(rp + static_cast<int64_t>(cols[bb]) * dp);
    float d0[8], d1[8], d2[8], d3[8], hs[8];
    raw[0].unpack(d0);
    raw[1].unpack(d1);
    raw[2].unpack(d2);
    raw[3].unpack(d3);
#pragma unroll
    for (int j = 0; j < 8; ++j) hs[j] = 0.25f * (d0[j] + d3[j]) + 0.75f * (d1[j] + d2[j]);
    // rows that use output row k: r = (k - a) / 2 for a in {0..3} with k - a even and 0 <= r < kUpRows
#pragma unroll
    for (int r = 0; r < kUpRows; ++r) {
      const int a = k - 2 * r;
      if (a >= 0 && a < 4) {
        const float wgt = (a == 0 || a == 3) ? 0.25f : 0.75f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(wgt, hs[j], acc[r][j]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kUpRows; ++r) {
    const int ih = ih0 + r;
    if (ih < H) Vec8<T>::st(dx + ((static_cast<int64_t>(n) * H + ih) * W + iw) * xp + c0, acc[r]);
  }
}

// ------------------------------------------------------------------------------------------------ layout
// NCHW fp32 -> NHWC bf16 through a 32x32 shared-memory transpose (coalesced on both sides)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t dp, int C,
                                    int64_t HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const int64_t p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? src[(static_cast<int64_t>(n) * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t p = p0 + j;
    const int c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[(static_cast<int64_t>(n) * HW + p) * dp + c] = from_f32<T>(tile[threadIdx.x][j]);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int64_t sp, float* __restrict__ dst, int C,
                                    int64_t HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t p = p0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? __bfloat162float(src[(static_cast<int64_t>(n) * HW + p) * sp + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const int64_t p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[(static_cast<int64_t>(n) * C + c) * HW + p] = tile[threadIdx.x][j];
  }
}


// fp32 NCHW image (C <= 8) -> bf16 NHWC with the channels zero-padded to 32: the operand layout of the tensor-core
// weight-gradient kernel for the stem (K = pixels; rows ci >= C of the result are zero and are dropped by the caller).
// One thread per pixel: C coalesced plane reads, one 64-byte row written.
__global__ void __launch_bounds__(256) image_to_nhwc32_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                               int C, int64_t HW) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  const int n = blockIdx.y;
  const int64_t p = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= HW) return;
  float f[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) f[c] = c < C ? __ldg(src + (static_cast<int64_t>(n) * C + c) * HW + p) : 0.f;
  uint4* o = reinterpret_cast<uint4*>(dst + (static_cast<int64_t>(n) * HW + p) * 32);
  o[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  o[1] = z;
  o[2] = z;
  o[3] = z;
}

}  // namespace b200

using namespace b200;

static int ilog2_exact(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

template <typename T>
static int upsample_fwd_impl(const void* x, int64_t x_pitch, void* out, int64_t out_pitch, int N, int H, int W, int C,
                             void* stream, const float* na = nullptr, const float* nb = nullptr, float slope = 0.f) {
  B200_CHECK_ARG(x && out, "upsample2x_fwd: null pointer");
  B200_CHECK_ARG(C % 8 == 0 && x_pitch % 8 == 0 && out_pitch % 8 == 0, "upsample2x_fwd: C and pitches must be multiples of 8");
  const int c8n = C / 8, sh = ilog2_exact(c8n);
  B200_CHECK_ARG(sh >= 0, "upsample2x_fwd: C/8 = %d must be a power of two", c8n);
  B200_CHECK_ARG(H <= 65535 && N <= 65535, "upsample2x_fwd: H and N must fit the grid");
  const dim3 block(256);
#define B200_UP_LAUNCH(RR)                                                                                             \
  launch_k(upsample2x_fwd_kernel<T, RR>, dim3(ceil_div(W * c8n, 256), ceil_div(H, RR), N), block, 0,                     \
           static_cast<cudaStream_t>(stream), static_cast<const T*>(x), x_pitch, static_cast<T*>(out), out_pitch, H, W,  \
           c8n, sh, na, nb, slope)
  // 8 rows per thread measured best at 64^2 .. 256^2 inputs (4: -6 %, 16: -1 %); the small levels need the blocks
  if (H >= 64) B200_UP_LAUNCH(8);
  else B200_UP_LAUNCH(2);
#undef B200_UP_LAUNCH
  B200_LAUNCH_CHECK("upsample2x_fwd_kernel");
  return 0;
}

template <typename T>
static int upsample_bwd_impl(const void* dout, int64_t dout_pitch, void* dx, int64_t dx_pitch, int N, int H, int W,
                             int C, void* stream) {
  B200_CHECK_ARG(dout && dx, "upsample2x_bwd: null pointer");
  B200_CHECK_ARG(C % 8 == 0 && dout_pitch % 8 == 0 && dx_pitch % 8 == 0, "upsample2x_bwd: C and pitches must be multiples of 8");
  const int c8n = C / 8, sh = ilog2_exact(c8n);
  B200_CHECK_ARG(sh >= 0, "upsample2x_bwd: C/8 = %d must be a power of two", c8n);
  B200_CHECK_ARG(H <= 65535 && N <= 65535, "upsample2x_bwd: H and N must fit the grid");
  launch_k(upsample2x_bwd_kernel<T>, dim3(ceil_div(W * c8n, 256), ceil_div(H, kUpRows), N), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const T*>(dout), dout_pitch, static_cast<T*>(dx), dx_pitch, H, W, c8n, sh);
  B200_LAUNCH_CHECK("upsample2x_bwd_kernel");
  return 0;
}

extern "C" int b200unet_upsample2x_fwd(const void* x, int64_t x_pitch, void* out, int64_t out_pitch, int N, int H,
                                       int W, int C, void* stream) {
  return upsample_fwd_impl<__nv_bfloat16>(x, x_pitch, out, out_pitch, N, H, W, C, stream);
}
extern "C" int b200unet_upsample2x_fwd_f32(const void* x, int64_t x_pitch, void* out, int64_t out_pitch, int N, int H,
                                           int W, int C, void* stream) {
  return upsample_fwd_impl<float>(x, x_pitch, out, out_pitch, N, H, W, C, stream);
}
extern "C" int b200unet_upsample2x_norm_fwd(const void* y, int64_t y_pitch, const float* a, const float* b, float slope,
                                            void* out, int64_t out_pitch, int N, int H, int W, int C, void* stream) {
  B200_CHECK_ARG(a && b, "upsample2x_norm_fwd: null affine");
  return upsample_fwd_impl<__nv_bfloat16>(y, y_pitch, out, out_pitch, N, H, W, C, stream, a, b, slope);
}
extern "C" int b200unet_upsample2x_norm_fwd_f32(const void* y, int64_t y_pitch, const float* a, const float* b,
                                                float slope, void* out, int64_t out_pitch, int N, int H, int W, int C,
                                                void* stream) {
  B200_CHECK_ARG(a && b, "upsample2x_norm_fwd: null affine");
  return upsample_fwd_impl<float>(y, y_pitch, out, out_pitch, N, H, W, C, stream, a, b, slope);
}
extern "C" int b200unet_upsample2x_bwd(const void* dout, int64_t dout_pitch, void* dx, int64_t dx_pitch, int N, int H,
                                       int W, int C, void* stream) {
  return upsample_bwd_impl<__nv_bfloat16>(dout, dout_pitch, dx, dx_pitch, N, H, W, C, stream);
}
extern "C" int b200unet_upsample2x_bwd_f32(const void* dout, int64_t dout_pitch, void* dx, int64_t dx_pitch, int N,
                                           int H, int W, int C, void* stream) {
  return upsample_bwd_impl<float>(dout, dout_pitch, dx, dx_pitch, N, H, W, C, stream);
}

extern "C" int b200unet_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int64_t dst_pitch, int N, int C, int64_t HW,
                                              void* stream) {
  B200_CHECK_ARG(src && dst, "nchw_f32_to_nhwc_bf16: null pointer");
  dim3 grid((unsigned)ceil_div64(HW, 32), ceil_div(C, 32), N);
  nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), dst_pitch, C, HW);
  B200_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}
extern "C" int b200unet_nchw_f32_to_nhwc_f32(const float* src, void* dst, int64_t dst_pitch, int N, int C, int64_t HW,
                                             void* stream) {
  B200_CHECK_ARG(src && dst, "nchw_f32_to_nhwc_f32: null pointer");
  dim3 grid((unsigned)ceil_div64(HW, 32), ceil_div(C, 32), N);
  nchw_to_nhwc_kernel<float><<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<float*>(dst), dst_pitch, C, HW);
  B200_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

extern "C" int b200unet_nhwc_bf16_to_nchw_f32(const void* src, int64_t src_pitch, float* dst, int N, int C, int64_t HW,
                                              void* stream) {
  B200_CHECK_ARG(src && dst, "nhwc_bf16_to_nchw_f32: null pointer");
  dim3 grid((unsigned)ceil_div64(HW, 32), ceil_div(C, 32), N);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), src_pitch, dst, C, HW);
  B200_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}

extern "C" int b200unet_image_to_nhwc32_bf16(const float* src, void* dst, int N, int C, int64_t HW, void* stream) {
  B200_CHECK_ARG(src && dst, "image_to_nhwc32_bf16: null pointer");
  B200_CHECK_ARG(C >= 1 && C <= 8 && N <= 65535, "image_to_nhwc32_bf16: C must be in [1,8]");
  launch_k(image_to_nhwc32_kernel, dim3((unsigned)ceil_div64(HW, 256), N), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      src, static_cast<__nv_bfloat16*>(dst), C, HW);
  B200_LAUNCH_CHECK("image_to_nhwc32_kernel");
  return 0;
}
