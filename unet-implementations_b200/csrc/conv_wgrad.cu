// Weight gradient of the 3x3 convolutions on tcgen05 / TMEM for the general case (any stride, Cout a multiple of 32):
//   D[128 = (tap, ci) rows, BN = co] += X_tap^T[rows, 64 pixels] . dY[64 pixels, BN]
// both operands MN-major (channels contiguous in NHWC), K = pixels, split over CTAs, fp32 partials reduced
// deterministically by wgrad_finalize_kernel into the OIHW gradient.  Stride-1 layers with Cout in {32, 64} take the
// narrow-output kernel (conv_wgrad_narrow.cu).  Replaces the weight-gradient half of aten::convolution_backward for
// nn.Conv2d (Our_UNet/models/unet.py:106-115).
// Warp roles (288 threads): warps 0..3 = TMA producers, warp 4 = TMEM allocator + MMA issuer, warps 5..8 = epilogue.
#include "common.cuh"
#include "ptx.cuh"
#include "conv_common.cuh"

namespace b200 {

struct WTap {
  int map, dh, dw;
};

struct WgradParams {
  int N, OH, OW, blocks_w, blocks_h, TWk, THk;
  int total_kb, kb_per_split;
  WTap taps[9];
  int cin, cout;
  int units_per_tap, total_units;
  int G;  // M-groups (of 128 rows) per CTA
  float* partial;  // [S][9][cin][cout]
};

struct WgradMaps {
  CUtensorMap src[4];
  CUtensorMap dy;
};

constexpr int kWgStages = 3;

// U  = channels per unit on the M side (64 -> 128B-swizzled blocks, 32 -> 64B-swizzled blocks)
// BN = output-channel tile (N side); 32 uses one 64B-swizzled block, otherwise BN/64 128B-swizzled blocks
template <int U, int BN>
__global__ void __launch_bounds__(kConvThreads, 1) wgrad_kernel(const __grid_constant__ WgradMaps maps,
                                                                 const __grid_constant__ WgradParams p) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel of the stream may become resident as CTAs retire
  constexpr int UPG = 128 / U;                      // units per M-group
  constexpr int kUnitBytes = 64 * U * 2;            // 64 pixels x U channels
  constexpr int kGroupBytes = UPG * kUnitBytes;     // 16 KB
  constexpr int kDyBlockCh = BN >= 64 ? 64 : BN;    // channels per dY block
  constexpr int kDyBlocks = BN / kDyBlockCh;
  constexpr int kDyBlockBytes = 64 * kDyBlockCh * 2;
  constexpr int kDyBytes = kDyBlocks * kDyBlockBytes;
  constexpr uint32_t kSwzA = (U == 64) ? kSwz128 : kSwz64;
  constexpr uint32_t kSwzB = (kDyBlockCh == 64) ? kSwz128 : kSwz64;
  constexpr uint32_t kRowA = U * 2, kRowB = kDyBlockCh * 2;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int G = p.G;
  const int stage_bytes = kDyBytes + G * kGroupBytes;
  const int split = blockIdx.x;
  const int group0 = blockIdx.y * G;
  const int n0 = blockIdx.z * BN;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.total_kb);
  const int num_kb = kb_end - kb_begin;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(G * BN)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], kProducerWarps);  // every producer arrives with the bytes of the loads it issues
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&tmem_base_holder, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  pdl_wait();  // barrier init / TMEM allocation above ran under the previous kernel's tail; global memory from here on

  // units of this CTA that exist (the last group of a layer may be partially filled)
  int valid_units = p.total_units - group0 * UPG;
  if (valid_units > G * UPG) valid_units = G * UPG;
  if (valid_units < 0) valid_units = 0;

  if (warp < kProducerWarps) {
    if (elect_one()) {
      // every producer walks all K-blocks and issues loads j = warp, warp+4, ... of each (dY blocks first, then units)
      const int blocks_per_img = p.blocks_w * p.blocks_h;
      for (int i = 0; i < num_kb; ++i) {
        const int kb = kb_begin + i;
        const int n_img = kb / blocks_per_img;
        const int b_in = kb - n_img * blocks_per_img;
        const int bh = b_in / p.blocks_w;
        const int bw = b_in - bh * p.blocks_w;
        const int h0 = bh * p.THk, w0 = bw * p.TWk;
        const int s = i % kWgStages;
        const uint32_t ph = (i / kWgStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint32_t my_bytes = 0;
        for (int j = warp; j < kDyBlocks + valid_units; j += kProducerWarps)
          my_bytes += (j < kDyBlocks) ? kDyBlockBytes : kUnitBytes;
        mbar_expect_tx(&full_bar[s], my_bytes);
        uint8_t* sb = smem + s * stage_bytes;
        for (int j = warp; j < kDyBlocks + valid_units; j += kProducerWarps) {
          if (j < kDyBlocks) {
            tma_load_4d(sb + j * kDyBlockBytes, &maps.dy, &full_bar[s], n0 + j * kDyBlockCh, w0, h0, n_img);
          } else {
            const int u = j - kDyBlocks;
            const int unit = group0 * UPG + u;
            const int tap = unit / p.units_per_tap;
            const int c0 = (unit - tap * p.units_per_tap) * U;
            const WTap t = p.taps[tap];
            tma_load_4d(sb + kDyBytes + u * kUnitBytes, &maps.src[t.map], &full_bar[s], c0, w0 + t.dw, h0 + t.dh,
                        n_img);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (elect_one()) {
      // lean issue loop (see ptx.cuh): 32-bit descriptor halves, the K advance is an add on the low word
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
      constexpr uint32_t a_hi = umma_desc_hi(8 * kRowA, kSwzA), b_hi = umma_desc_hi(8 * kRowB, kSwzB);
      constexpr uint32_t kStepA = (16 * kRowA) >> 4, kStepB = (16 * kRowB) >> 4;  // 16 pixels of K, in 16-byte units
      const int groups = (valid_units + UPG - 1) / UPG;
      const uint32_t lo0 = umma_desc_lo(smem_u32(smem), 0);
      constexpr uint32_t a_lbo = ((kUnitBytes >> 4) & 0x3FFFu) << 16, b_lbo = ((kDyBlockBytes >> 4) & 0x3FFFu) << 16;
      const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
      uint32_t s = 0, ph = 0, acc = 0;
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t b_lo = lo0 + s * stage16 + b_lbo;
        uint32_t a_lo = lo0 + s * stage16 + (kDyBytes >> 4) + a_lbo;
        for (int g = 0; g < groups; ++g, a_lo += (kGroupBytes >> 4)) {
          // MN-major canonical layout: LBO = byte stride between channel blocks, SBO = 8 pixel rows
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lean(tmem_base + g * BN, a_lo + k * kStepA, a_hi, b_lo + k * kStepB, b_hi, idesc, acc | k);
        }
        acc = 1;
        umma_commit(&empty_bar[s]);
        if (++s == kWgStages) { s = 0; ph ^= 1; }
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (num_kb > 0) {
      mbar_wait(&tmem_full_bar, 0);
      tc_fence_after();
    }
    const int groups = (valid_units + UPG - 1) / UPG;
    for (int g = 0; g < groups; ++g) {
      const int u = g * UPG + row / U;
      const int unit = group0 * UPG + u;
      const bool uvalid = u < valid_units;
      int tap = 0, ci = 0;
      if (uvalid) {
        tap = unit / p.units_per_tap;
        ci = (unit - tap * p.units_per_tap) * U + (row % U);
      }
      float* dst = p.partial + ((static_cast<size_t>(split) * 9 + tap) * p.cin + ci) * p.cout + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        if (num_kb > 0) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * BN + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0;
        }
        if (uvalid) {
          float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dw[co][ci][tap] = sum_s partial[s][tap][ci][co]
__global__ void wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int S, int cin,
                                      int cout) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  const int64_t total = static_cast<int64_t>(9) * cin * cout;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % cout);
  const int64_t r = i / cout;
  const int ci = static_cast<int>(r % cin);
  const int tap = static_cast<int>(r / cin);
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += partial[static_cast<int64_t>(s) * total + i];
  dw[(static_cast<int64_t>(co) * cin + ci) * 9 + tap] = acc;
}

// Same reduction through a shared-memory transpose: a block owns 32 output channels x 8 input channels; the partials
// are read along co (128-byte rows), the OIHW result is written as 288-byte runs (8 ci x 9 taps per output channel).
// The plain kernel above scatters 4-byte writes 36 bytes apart: 0.6 ms per step over the 22 layers (ncu).
__global__ void __launch_bounds__(256) wgrad_finalize_tiled_kernel(const float* __restrict__ partial,
                                                                    float* __restrict__ dw, int S, int cin, int cout) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  __shared__ float tile[32][73];
  const int co_l = threadIdx.x & 31, ci_l = threadIdx.x >> 5;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 8;
  const int64_t total = static_cast<int64_t>(9) * cin * cout;
  // the sum over s keeps its fixed order, but the loads of 8 splits are issued together: one dependent load per
  // addition made this kernel latency-bound (9 S serial L2 round trips: 20 us for 38 MB)
#pragma unroll 3
  for (int tap = 0; tap < 9; ++tap) {
    const float* src = partial + (static_cast<int64_t>(tap) * cin + ci0 + ci_l) * cout + co0 + co_l;
    float acc = 0.f;
    int s = 0;
    for (; s + 8 <= S; s += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcs(src + static_cast<int64_t>(s + j) * total);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    if (s + 4 <= S) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldcs(src + static_cast<int64_t>(s + j) * total);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += v[j];
      s += 4;
    }
    for (; s < S; ++s) acc += __ldcs(src + static_cast<int64_t>(s) * total);
    tile[co_l][ci_l * 9 + tap] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * 72; idx += 256) {
    const int c = idx / 72, r = idx - c * 72;
    dw[(static_cast<int64_t>(co0 + c) * cin + ci0) * 9 + r] = tile[c][r];
  }
}

// Many partials, small tensor (the narrow layers: S = 148 splits of a 9 x 32 x 32 .. 9 x 192 x 64 gradient): the sum
// over S is the long axis, so 8 thread groups of a block share it (fixed-order combine); a block owns 32 consecutive
// output channels of one (tap, ci).
__global__ void __launch_bounds__(256) wgrad_finalize_splits_kernel(const float* __restrict__ partial,
                                                                     float* __restrict__ dw, int S, int cin, int cout) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel may get resident; wait for the previous one
  pdl_wait();
  __shared__ float red[8][32];
  const int co_l = threadIdx.x & 31, z = threadIdx.x >> 5;
  const int64_t total = static_cast<int64_t>(9) * cin * cout;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 32 + co_l;  // [tap][ci][co] index
  const int s_per = (S + 7) / 8;
  const int s_lo = z * s_per, s_hi = min(S, s_lo + s_per);
  float a0 = 0.f, a1 = 0.f;
  int s = s_lo;
  for (; s + 1 < s_hi; s += 2) {
    a0 += partial[static_cast<int64_t>(s) * total + i];
    a1 += partial[static_cast<int64_t>(s + 1) * total + i];
  }
  if (s < s_hi) a0 += partial[static_cast<int64_t>(s) * total + i];
  red[z][co_l] = a0 + a1;
  __syncthreads();
  if (z == 0) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][co_l];
    const int co = static_cast<int>(i % cout);
    const int64_t r = i / cout;
    const int ci = static_cast<int>(r % cin);
    const int tap = static_cast<int>(r / cin);
    dw[(static_cast<int64_t>(co) * cin + ci) * 9 + tap] = v;
  }
}

int launch_wgrad_finalize(const float* partial, float* dw, int S, int cin, int cout, cudaStream_t st) {
  if (cout % 32 == 0 && S >= 16) {
    launch_k(wgrad_finalize_splits_kernel, dim3((unsigned)(static_cast<int64_t>(9) * cin * cout / 32)), dim3(256), 0, st, partial, dw, S, cin,
                                                                                                        cout);
    B200_LAUNCH_CHECK("wgrad_finalize_splits_kernel");
    return 0;
  }
  if (cin % 8 == 0 && cout % 32 == 0) {
    launch_k(wgrad_finalize_tiled_kernel, dim3(cout / 32, cin / 8), dim3(256), 0, st, partial, dw, S, cin, cout);
    B200_LAUNCH_CHECK("wgrad_finalize_tiled_kernel");
    return 0;
  }
  const int64_t total = static_cast<int64_t>(9) * cin * cout;
  launch_k(wgrad_finalize_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, st, partial, dw, S, cin, cout);
  B200_LAUNCH_CHECK("wgrad_finalize_kernel");
  return 0;
}

struct WgradPlan {
  int U, BN, G, S, gy, gz, total_kb, kb_per_split, TWk, THk, blocks_w, blocks_h, OH, OW;
  int64_t smem_bytes, partial_floats;
};

static int plan_wgrad(int N, int H, int W, int Cin, int Cout, int stride, WgradPlan* pl) {
  pl->U = pick_bk(Cin);
  pl->BN = pick_bn(Cout);
  if (!pl->U || !pl->BN) return set_error(kErrUnsupported, "conv_wgrad: Cin=%d Cout=%d outside the envelope", Cin, Cout);
  pl->OH = (H - 1) / stride + 1;
  pl->OW = (W - 1) / stride + 1;
  pl->TWk = 16;
  pl->THk = 4;
  pl->blocks_w = ceil_div(pl->OW, pl->TWk);
  pl->blocks_h = ceil_div(pl->OH, pl->THk);
  pl->total_kb = N * pl->blocks_w * pl->blocks_h;
  const int upg = 128 / pl->U;
  const int total_units = 9 * (Cin / pl->U);
  const int groups_total = ceil_div(total_units, upg);
  int G = 512 / pl->BN;
  const int dy_bytes = 64 * pl->BN * 2;
  const int gsm = (70 * 1024 - dy_bytes) / 16384;  // keep a stage under ~70 KB so three stages fit
  if (G > gsm) G = gsm;
  if (G > groups_total) G = groups_total;
  if (G < 1) G = 1;
  pl->G = G;
  pl->gy = ceil_div(groups_total, G);
  pl->gz = Cout / pl->BN;
  const int base_ctas = pl->gy * pl->gz;
  // split K so that the whole grid is ONE wave (<= one CTA per SM): rounding the split count up instead
  // (ceil(148 / base)) left a second, almost empty wave on nearly every layer of the model (e.g. 27 x 6 = 162 CTAs)
  int S = num_sms() / base_ctas;
  if (S > pl->total_kb) S = pl->total_kb;
  if (S < 1) S = 1;
  pl->kb_per_split = ceil_div(pl->total_kb, S);
  pl->S = ceil_div(pl->total_kb, pl->kb_per_split);
  pl->smem_bytes = static_cast<int64_t>(kWgStages) * (dy_bytes + G * 16384) + 1024;
  pl->partial_floats = static_cast<int64_t>(pl->S) * 9 * Cin * Cout;
  return 0;
}

template <int U, int BN>
static int launch_wgrad(const WgradMaps& maps, const WgradParams& p, const WgradPlan& pl, cudaStream_t st) {
  auto kern = wgrad_kernel<U, BN>;
  static int attr_bytes = 0;
  if (attr_bytes < pl.smem_bytes) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    attr_bytes = (int)pl.smem_bytes;
  }
  dim3 grid(pl.S, pl.gy, pl.gz);
  launch_k(kern, dim3(grid), dim3(kConvThreads), pl.smem_bytes, st, maps, p);
  B200_LAUNCH_CHECK("wgrad_kernel");
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200unet_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int stride) {
  if (wgradn_supported(Cin, Cout, stride)) return wgradn_workspace_bytes(N, H, W, Cin, Cout);
  WgradPlan pl;
  if (plan_wgrad(N, H, W, Cin, Cout, stride, &pl)) return -1;
  return pl.partial_floats * 4;
}

extern "C" int b200unet_conv_wgrad(const b200unet_conv_wgrad_args* a, void* stream) {
  B200_CHECK_ARG(a && a->x && a->dy && a->dw && a->workspace, "conv_wgrad: null pointer");
  B200_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv_wgrad: stride %d unsupported", a->stride);
  B200_CHECK_ARG(a->x_pitch % 8 == 0 && a->dy_pitch % 8 == 0, "conv_wgrad: pitches must be multiples of 8 elements");
  if (wgradn_supported(a->Cin, a->Cout, a->stride)) return wgradn_launch(a, static_cast<cudaStream_t>(stream));
  WgradPlan pl;
  int rc;
  if ((rc = plan_wgrad(a->N, a->H, a->W, a->Cin, a->Cout, a->stride, &pl))) return rc;
  B200_CHECK_ARG(a->workspace_bytes >= pl.partial_floats * 4, "conv_wgrad: workspace too small (%lld < %lld)",
                 (long long)a->workspace_bytes, (long long)(pl.partial_floats * 4));
  WgradParams p{};
  WgradMaps maps;
  p.N = a->N;
  p.OH = pl.OH;
  p.OW = pl.OW;
  p.blocks_w = pl.blocks_w;
  p.blocks_h = pl.blocks_h;
  p.TWk = pl.TWk;
  p.THk = pl.THk;
  p.total_kb = pl.total_kb;
  p.kb_per_split = pl.kb_per_split;
  p.cin = a->Cin;
  p.cout = a->Cout;
  p.units_per_tap = a->Cin / pl.U;
  p.total_units = 9 * p.units_per_tap;
  p.G = pl.G;
  p.partial = a->workspace;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a->x);
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(a->dy);
  if (a->stride == 1) {
    if ((rc = make_act_map(&maps.src[0], x, a->x_pitch, a->N, a->H, a->W, a->Cin, 1, 1, 0, 0, pl.U, pl.TWk, pl.THk)))
      return rc;
    maps.src[1] = maps.src[2] = maps.src[3] = maps.src[0];
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) p.taps[kh * 3 + kw] = WTap{0, kh - 1, kw - 1};
  } else {
    B200_CHECK_ARG(a->H >= 2 && a->W >= 2, "conv_wgrad: stride-2 input must be at least 2x2");
    for (int hp = 0; hp < 2; ++hp)
      for (int wp = 0; wp < 2; ++wp)
        if ((rc = make_act_map(&maps.src[hp * 2 + wp], x, a->x_pitch, a->N, a->H, a->W, a->Cin, 2, 2, hp, wp, pl.U,
                               pl.TWk, pl.THk)))
          return rc;
    const int par[3] = {1, 0, 1}, sh[3] = {-1, 0, 0};
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) p.taps[kh * 3 + kw] = WTap{par[kh] * 2 + par[kw], sh[kh], sh[kw]};
  }
  const int dyc = pl.BN >= 64 ? 64 : pl.BN;
  if ((rc = make_act_map(&maps.dy, dy, a->dy_pitch, a->N, pl.OH, pl.OW, a->Cout, 1, 1, 0, 0, dyc, pl.TWk, pl.THk)))
    return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define WG(u, bn) \
  if (pl.U == u && pl.BN == bn) rc = launch_wgrad<u, bn>(maps, p, pl, st); else
  WG(64, 256) WG(64, 128) WG(64, 64) WG(64, 32) WG(32, 256) WG(32, 128) WG(32, 64) WG(32, 32)
  rc = set_error(kErrUnsupported, "no wgrad instantiation for U=%d BN=%d", pl.U, pl.BN);
#undef WG
  if (rc) return rc;
  return launch_wgrad_finalize(a->workspace, a->dw, pl.S, a->Cin, a->Cout, st);
}
