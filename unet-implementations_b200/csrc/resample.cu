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

namespace b200 {

__device__ __forceinline__ void unpack8r(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// one thread = 8 channels of one OUTPUT pixel
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t xp,
                                                              __nv_bfloat16* __restrict__ out, int64_t op, int N, int H,
                                                              int W, int C) {
  const int c8n = C >> 3;
  const int OH = 2 * H, OW = 2 * W;
  const int64_t total = static_cast<int64_t>(N) * OH * OW * c8n;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % c8n) << 3;
  int64_t r = i / c8n;
  const int ow = static_cast<int>(r % OW);
  r /= OW;
  const int oh = static_cast<int>(r % OH);
  const int n = static_cast<int>(r / OH);
  // source rows/cols and weights
  const int ih = oh >> 1, iw = ow >> 1;
  const int h_near = ih, h_far = (oh & 1) ? min(ih + 1, H - 1) : max(ih - 1, 0);
  const int w_near = iw, w_far = (ow & 1) ? min(iw + 1, W - 1) : max(iw - 1, 0);
  const __nv_bfloat16* b = x + static_cast<int64_t>(n) * H * W * xp + c0;
  float nn[8], nf[8], fn[8], ff[8];
  unpack8r(*reinterpret_cast<const uint4*>(b + (static_cast<int64_t>(h_near) * W + w_near) * xp), nn);
  unpack8r(*reinterpret_cast<const uint4*>(b + (static_cast<int64_t>(h_near) * W + w_far) * xp), nf);
  unpack8r(*reinterpret_cast<const uint4*>(b + (static_cast<int64_t>(h_far) * W + w_near) * xp), fn);
  unpack8r(*reinterpret_cast<const uint4*>(b + (static_cast<int64_t>(h_far) * W + w_far) * xp), ff);
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // same association as ATen's upsample_bilinear2d: interpolate along w, then along h
    const float top = 0.75f * nn[j] + 0.25f * nf[j];
    const float bot = 0.75f * fn[j] + 0.25f * ff[j];
    o[j] = 0.75f * top + 0.25f * bot;
  }
  *reinterpret_cast<uint4*>(out + ((static_cast<int64_t>(n) * OH + oh) * OW + ow) * op + c0) =
      make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

// one thread = 8 channels of one INPUT pixel; gathers its 4x4 output neighbourhood
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int64_t dp,
                                                              __nv_bfloat16* __restrict__ dx, int64_t xp, int N, int H,
                                                              int W, int C) {
  const int c8n = C >> 3;
  const int OH = 2 * H, OW = 2 * W;
  const int64_t total = static_cast<int64_t>(N) * H * W * c8n;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % c8n) << 3;
  int64_t r = i / c8n;
  const int iw = static_cast<int>(r % W);
  r /= W;
  const int ih = static_cast<int>(r % H);
  const int n = static_cast<int>(r / H);
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const __nv_bfloat16* b = dout + static_cast<int64_t>(n) * OH * OW * dp + c0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int oh = 2 * ih - 1 + a;
    oh = oh < 0 ? 0 : (oh >= OH ? OH - 1 : oh);
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      int ow = 2 * iw - 1 + bb;
      ow = ow < 0 ? 0 : (ow >= OW ? OW - 1 : ow);
      float d[8];
      unpack8r(*reinterpret_cast<const uint4*>(b + (static_cast<int64_t>(oh) * OW + ow) * dp), d);
      const float wgt = wt[a] * wt[bb];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, d[j], acc[j]);
    }
  }
  *reinterpret_cast<uint4*>(dx + ((static_cast<int64_t>(n) * H + ih) * W + iw) * xp + c0) =
      make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                 pack_bf16x2(acc[6], acc[7]));
}

// ------------------------------------------------------------------------------------------------ layout
// NCHW fp32 -> NHWC bf16 through a 32x32 shared-memory transpose (coalesced on both sides)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t dp, int C,
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
    if (c < C && p < HW) dst[(static_cast<int64_t>(n) * HW + p) * dp + c] = __float2bfloat16_rn(tile[threadIdx.x][j]);
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

}  // namespace b200

using namespace b200;

extern "C" int b200unet_upsample2x_fwd(const void* x, int64_t x_pitch, void* out, int64_t out_pitch, int N, int H,
                                       int W, int C, void* stream) {
  B200_CHECK_ARG(x && out, "upsample2x_fwd: null pointer");
  B200_CHECK_ARG(C % 8 == 0 && x_pitch % 8 == 0 && out_pitch % 8 == 0, "upsample2x_fwd: C and pitches must be multiples of 8");
  const int64_t total = static_cast<int64_t>(N) * 4 * H * W * (C / 8);
  upsample2x_fwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(out), out_pitch, N, H, W, C);
  B200_LAUNCH_CHECK("upsample2x_fwd_kernel");
  return 0;
}

extern "C" int b200unet_upsample2x_bwd(const void* dout, int64_t dout_pitch, void* dx, int64_t dx_pitch, int N, int H,
                                       int W, int C, void* stream) {
  B200_CHECK_ARG(dout && dx, "upsample2x_bwd: null pointer");
  B200_CHECK_ARG(C % 8 == 0 && dout_pitch % 8 == 0 && dx_pitch % 8 == 0, "upsample2x_bwd: C and pitches must be multiples of 8");
  const int64_t total = static_cast<int64_t>(N) * H * W * (C / 8);
  upsample2x_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), dout_pitch, static_cast<__nv_bfloat16*>(dx), dx_pitch, N, H, W, C);
  B200_LAUNCH_CHECK("upsample2x_bwd_kernel");
  return 0;
}

extern "C" int b200unet_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int64_t dst_pitch, int N, int C, int64_t HW,
                                              void* stream) {
  B200_CHECK_ARG(src && dst, "nchw_f32_to_nhwc_bf16: null pointer");
  dim3 grid((unsigned)ceil_div64(HW, 32), ceil_div(C, 32), N);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), dst_pitch, C, HW);
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
