// 3x3 / stride-1 convolutions 32 -> 32 channels on dense tensors (the 512^2 level of the UNet: e0c2, d4c2, forward and
// data gradient) with PIXEL PAIRS as operand rows.  Same role as conv_narrow.cu (nn.Conv2d forward and the data-gradient
// half of aten::convolution_backward, Our_UNet/models/unet.py:106-115); dispatched from nconv_launch.
//
// Why.  With 32 K-side channels an operand row is 64 bytes; K-major operands in the 64 B swizzle are fetched at about
// half the rate of 128-byte rows (a 128 x 96 x 16 MMA takes ~145 cycles in nconv<32,32> instead of 64: role counters,
// DESIGN.md 3), and zero-padding the row to 128 bytes halves the real bytes in flight instead.  A dense NHWC tensor
// with 32 channels IS a tensor with 64 channels and W/2 pixels: two horizontally adjacent pixels (x0, x1) form one
// 128-byte row, K = (pixel parity, ci) = 64, all of it real data.  The column taps become four products per pair P:
//     c0[P] = x0 W[kw=1] + x1 W[kw=2]    -> y0[P]        (y0 = output at pixel 2P,  y1 = output at pixel 2P+1)
//     c1[P] = x0 W[kw=0] + x1 W[kw=1]    -> y1[P]
//     c2[P] =              x1 W[kw=0]    -> y0[P+1]
//     c3[P] = x0 W[kw=2]                 -> y1[P-1]
// stacked on N = 4 * 32 = 128 (a full-rate MMA; 6 of its 8 weight blocks are non-zero, the same 75 % as N = 96), and
//     y0[P] = c0[P] + c2[P-1],   y1[P] = c1[P] + c3[P+1]
// is one warp shuffle per output pixel and channel instead of two.  The row taps (kh) are window offsets into one
// 6-row patch as in conv_narrow.cu.  A tile is 4 image rows x 32 pairs (lanes 0 and 31 are halo: 4 x 60 outputs), the
// output row of a pair is again 128 bytes, stored through the same (W/2, 64-channel) view.  The block-structured
// weight tiles [3 kh][(co, j) x (parity, ci)] are built in shared memory by the epilogue warps at kernel start from
// the ordinary packed weights (18 KB from L2 per CTA), so the C ABI and the packing are unchanged; the data gradient
// only reads them flipped (kh -> 2 - kh, kw -> 2 - kw).
// Roles (512 threads): warps 0..1 = TMA producers (one 24 KB patch per tile, 6 in flight), warp 2 = TMEM owner + the
// MMA-issuing thread (12 MMAs per tile), warps 4..11 = two epilogue groups alternating tiles (tcgen05.ld, shuffle,
// bf16 staging, TMA store), warps 12..15 = InstanceNorm partial sums of the staged tiles (forward only), handed over
// through named barriers.
#include "common.cuh"
#include "ptx.cuh"
#include "conv_common.cuh"
#include <stdlib.h>

namespace b200 {

constexpr int kPcTH = 4, kPcTW = 32, kPcValidW = 30;  // tile rows / pair columns (lanes) / valid pair columns
constexpr int kPcPatchRows = kPcTH + 2;
constexpr int kPcProducers = 2, kPcMmaWarp = 2, kPcEpiWarp0 = 4;
constexpr int kPcNG = 2;       // epilogue groups = accumulators
constexpr int kPcStatWarp0 = kPcEpiWarp0 + 4 * kPcNG;  // four statistics warps after the epilogue groups
constexpr int kPcC = 32;       // channels on both sides
constexpr int kPcSlots = 6;    // patches in flight

struct PConvParams {
  int N, H, Wp, tiles_w, tiles_h;  // Wp = W / 2 pair columns
  int tiles_per_cta;
  int stat_slots;
  float* stats;
  const __nv_bfloat16* w;  // packed [32 N-side][3][3][32 K-side]
  BwdSums bs;              // data gradient only: bs.y != NULL turns the statistics warps into the norm-backward reduction
};

struct PConvMaps {
  CUtensorMap src;  // (64, W/2, H, N) view, box (64, 32, 6, 1)
  CUtensorMap out;  // (64, W/2, H, N) view, box (64, 30, 4, 1)
};

struct PConvCfg {
  static constexpr int kThreads = 32 * (kPcStatWarp0 + 4);
  static constexpr int kASlotBytes = kPcPatchRows * kPcTW * 128;  // 24 KB
  static constexpr int kWin16 = (kPcTW * 128) >> 4;                // one image row of the patch, 16-byte units
  static constexpr int kBTileBytes = 128 * 128;                    // [(co, j) x 64] of one kh
  static constexpr int kBBytes = 3 * kBTileBytes;
  static constexpr int kStageBufBytes = 16 * 1024;                 // 4 x 30 pairs x 128 B = 15 KB
  static constexpr int kSmemBytes = kPcSlots * kASlotBytes + kBBytes + kPcNG * kStageBufBytes + 1024;
  static constexpr uint32_t kTmemCols = 256;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert(kPcSlots % kPcProducers == 0, "each producer owns a fixed subset of slots");
};

// which column tap (of the effective, i.e. already flipped for the data gradient, kernel) product j takes from the pixel
// of parity pi; -1 = zero block
__device__ __forceinline__ int pc_tap(int j, int pi) {
  // j:        0        1        2        3
  // pi = 0:   1        0       -1        2
  // pi = 1:   2        1        0       -1
  const int t0 = (j == 0) ? 1 : (j == 1) ? 0 : (j == 2) ? -1 : 2;
  const int t1 = (j == 0) ? 2 : (j == 1) ? 1 : (j == 2) ? 0 : -1;
  return pi ? t1 : t0;
}

template <bool REV>
__global__ void __launch_bounds__(PConvCfg::kThreads, 1) pconv_kernel(const __grid_constant__ PConvMaps maps,
                                                                       const __grid_constant__ PConvParams p) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel of the stream may become resident as CTAs retire
  using Cfg = PConvCfg;
  constexpr int NG = kPcNG, A_SLOTS = kPcSlots;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[A_SLOTS], a_empty[A_SLOTS];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t tmem_full_bar[NG], tmem_empty_bar[NG];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + A_SLOTS * Cfg::kASlotBytes;
  uint8_t* staging = smem_b + Cfg::kBBytes;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int total_tiles = tiles_per_img * p.N;
  const int tile_lo = min(static_cast<int>(blockIdx.x) * p.tiles_per_cta, total_tiles);
  const int tile_hi = min(tile_lo + p.tiles_per_cta, total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_SLOTS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(&b_full, NG * 128);  // every epilogue thread arrives once after writing its share of the weight tiles
    for (int b = 0; b < NG; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == kPcMmaWarp) {
    tmem_alloc(&tmem_base_holder, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  pdl_wait();  // barrier init / TMEM allocation above ran under the previous kernel's tail; global memory from here on

  if (warp < kPcProducers) {
    if (elect_one()) {
      tma_prefetch_desc(&maps.src);
      int it = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        if ((it % kPcProducers) != warp) continue;
        const int n_img = tile / tiles_per_img;
        const int t_in = tile - n_img * tiles_per_img;
        // column-major tile order inside an image: the next tile is the one below, its 2 halo rows are in L2
        const int tw_ = t_in / p.tiles_h;
        const int h0 = (t_in - tw_ * p.tiles_h) * kPcTH, w0 = tw_ * kPcValidW;
        const int slot = it % A_SLOTS;
        const uint32_t ph = static_cast<uint32_t>((it / A_SLOTS) & 1);
        mbar_wait(&a_empty[slot], ph ^ 1);
        mbar_expect_tx(&a_full[slot], Cfg::kASlotBytes);
        tma_load_4d(smem + slot * Cfg::kASlotBytes, &maps.src, &a_full[slot], 0, w0 - 1, h0 - 1, n_img);
      }
    }
  } else if (warp == kPcMmaWarp) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t hi = umma_desc_hi(8 * 128, kSwz128);
      constexpr uint32_t kASlot16 = Cfg::kASlotBytes >> 4, kBTile16 = Cfg::kBTileBytes >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem), 16);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b), 16);
      const uint32_t a_full0 = smem_u32(&a_full[0]), a_empty0 = smem_u32(&a_empty[0]);
      mbar_wait(&b_full, 0);
      tc_fence_after();
      uint32_t aslot = 0, aph = 0;
      int it = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        const int buf = it % NG;
        mbar_wait(&tmem_empty_bar[buf], static_cast<uint32_t>(((it / NG) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 128;
        mbar_wait_u32(a_full0 + aslot * 8, aph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + aslot * kASlot16;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lean(d_tmem, a_lo + kh * Cfg::kWin16 + 2 * k, hi, b_lo0 + kh * kBTile16 + 2 * k, hi, idesc,
                           (kh | k) ? 1u : 0u);
        }
        umma_commit_u32(a_empty0 + aslot * 8);
        if (++aslot == A_SLOTS) { aslot = 0; aph ^= 1; }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else if (warp >= kPcEpiWarp0 && warp < kPcStatWarp0) {
    const int g = (warp - kPcEpiWarp0) >> 2;
    const int q = warp & 3;
    const int et = threadIdx.x - 32 * (kPcEpiWarp0 + 4 * g);  // 0..127 inside the group
    // ---------------------------------------------------------------- weight tiles: [kh][n = 4*co + j][k = 32*pi + ci]
    {
      const int ft = threadIdx.x - 32 * kPcEpiWarp0;  // 0 .. NG*128 - 1
      for (int idx = ft; idx < 3 * 128 * 8; idx += NG * 128) {
        const int kh = idx >> 10;
        const int r = idx & 1023;
        const int n = r >> 3, c = r & 7;  // row, 16-byte chunk (8 K-side channels)
        const int co = n >> 2, j = n & 3, pi = c >> 2, ci0 = (c & 3) << 3;
        const int tap = pc_tap(j, pi);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (tap >= 0) {
          const int khs = REV ? 2 - kh : kh, kws = REV ? 2 - tap : tap;
          v = __ldg(reinterpret_cast<const uint4*>(p.w + ((co * 3 + khs) * 3 + kws) * kPcC + ci0));
        }
        *reinterpret_cast<uint4*>(smem_b + kh * Cfg::kBTileBytes + n * 128 + (((c ^ (n & 7)) & 7) << 4)) = v;
      }
      fence_proxy_async_smem();
      mbar_arrive(&b_full);
    }
    const int bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    const bool do_stats = p.stats != nullptr;
    const bool col_ok = lane >= 1 && lane <= kPcValidW;
    const int srow = q * kPcValidW + lane - 1;  // staging row of this thread's output pair
    uint8_t* stg = staging + g * Cfg::kStageBufBytes;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 128;
    int tile = tile_lo + g;
    int n_img = tile / tiles_per_img;
    int tw = (tile - n_img * tiles_per_img) / p.tiles_h;
    int th = tile - n_img * tiles_per_img - tw * p.tiles_h;
    uint32_t ph = 0;
    for (; tile < tile_hi; tile += NG, ph ^= 1) {
      const int h0 = th * kPcTH, w0 = tw * kPcValidW;
      mbar_wait(&tmem_full_bar[g], ph);
      tc_fence_after();
      if (et == 0) tma_store_wait_read_all();  // this group's previous store has read the staging buffer
      named_bar_sync(bar_a, 128);
      if (do_stats && tile != tile_lo + g) named_bar_sync(7 + g, 256);  // ... and so have the statistics warps
      const uint32_t base = smem_u32(stg) + srow * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < kPcC; c0 += 8) {
        // 8 output channels = 32 accumulator columns (co, j) starting at 4 * c0
        uint32_t v[32];
        tmem_ld_32x16(t_addr + 4 * c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld_32x16(t_addr + 4 * c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        tmem_ld_wait();
        if (c0 + 8 >= kPcC) {
          // all TMEM reads of this warp are done: hand the accumulator back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[g]);
        }
        float y0[8], y1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          y0[c] = __uint_as_float(v[4 * c]) + __shfl_up_sync(0xffffffffu, __uint_as_float(v[4 * c + 2]), 1);
          y1[c] = __uint_as_float(v[4 * c + 1]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[4 * c + 3]), 1);
        }
        if (col_ok) {
          // chunk cj of the pair's 128-byte row: channels [8 cj, 8 cj + 8) of pixel 0 (cj < 4) or pixel 1
          const int cj0 = c0 >> 3, cj1 = 4 + (c0 >> 3);
          const uint32_t a0 = base + (((cj0 ^ (srow & 7)) & 7) << 4), a1 = base + (((cj1 ^ (srow & 7)) & 7) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pack_bf16x2(y0[0], y0[1])),
                       "r"(pack_bf16x2(y0[2], y0[3])), "r"(pack_bf16x2(y0[4], y0[5])), "r"(pack_bf16x2(y0[6], y0[7]))
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(pack_bf16x2(y1[0], y1[1])),
                       "r"(pack_bf16x2(y1[2], y1[3])), "r"(pack_bf16x2(y1[4], y1[5])), "r"(pack_bf16x2(y1[6], y1[7]))
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(bar_b, 128);
      if (et == 0) {
        tma_store_4d(&maps.out, stg, 0, w0, h0, n_img);
        tma_store_commit();
      }
      if (do_stats) named_bar_arrive(5 + g, 256);  // staged tile complete: hand it to the statistics warps
      th += NG;
      while (th >= p.tiles_h) {
        th -= p.tiles_h;
        if (++tw == p.tiles_w) {
          tw = 0;
          ++n_img;
        }
      }
    }
    if (et == 0) tma_store_wait_all();
  } else if (warp >= kPcStatWarp0 && p.stats != nullptr) {
    // ------------------------------------------------ statistics warps: the InstanceNorm partial sums of every staged
    // tile, read from the staging buffers beside the TMA store.  In the epilogue groups this pass cost 50 us of 270
    // (its shared-memory reads sat on each group's serial path); here it only has to keep up with the tile rate.
    const int q = warp & 3;
    const int n_tiles = tile_hi - tile_lo;
    // InstanceNorm partial sums of the STORED bf16 values: lane = (row group lane >> 3, 16-byte chunk lane & 7) of the
    // staged 128-byte rows, 8 channels per lane; summed over all tiles of an image in registers.
    // Forward: (sum y, sum y^2).  Data gradient with p.bs.y (producer-side norm-backward sums): the staged values are
    // dz, the matching 16 bytes of the consuming unit's raw output y come straight from global memory (issued BEFORE the
    // wait for the staged tile, 8 loads in flight per lane), and the sums are (sum gm, sum gm * y), gm = dz * lrelu'.
    const bool bwd = REV && p.bs.y != nullptr;
    float s1[8], s2[8], pa[8], pb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = pa[i] = pb[i] = 0.f;
    int acc_img = -1;
    auto flush = [&](int img) {
      const int first_tile = img * tiles_per_img;
      const int b0 = first_tile / p.tiles_per_cta;
      const int slot = (static_cast<int>(blockIdx.x) - b0) * NG * 4 + q;
      float* dst = p.stats + (static_cast<size_t>(img) * p.stat_slots + slot) * kPcC * 2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // over the four row groups (lane bits 3, 4), then over the two pixel parities (chunk bit 2 = lane bit 2)
        float a = s1[i], b = s2[i];
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        b += __shfl_xor_sync(0xffffffffu, b, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 16);
        b += __shfl_xor_sync(0xffffffffu, b, 16);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        b += __shfl_xor_sync(0xffffffffu, b, 4);
        s1[i] = a;
        s2[i] = b;
      }
      if (lane < 4) {  // channels [8 lane, 8 lane + 8)
        float4* d4 = reinterpret_cast<float4*>(dst + 16 * lane);
        d4[0] = make_float4(s1[0], s2[0], s1[1], s2[1]);
        d4[1] = make_float4(s1[2], s2[2], s1[3], s2[3]);
        d4[2] = make_float4(s1[4], s2[4], s1[5], s2[5]);
        d4[3] = make_float4(s1[6], s2[6], s1[7], s2[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    };
    int n_img = tile_lo / tiles_per_img;
    int tw = (tile_lo - n_img * tiles_per_img) / p.tiles_h;
    int th = tile_lo - n_img * tiles_per_img - tw * p.tiles_h;
    const int rsub = lane >> 3, ch = lane & 7;
    constexpr int kSteps = (kPcValidW + 3) / 4;
    // y words of tile (img, th_, tw_): this lane's 16 bytes of rows [30q, 30q + 30) four rows per step
    auto load_y = [&](int img, int th_, int tw_, uint4 (&dst)[kSteps]) {
      const int hh = th_ * kPcTH + q, ww0 = tw_ * kPcValidW;
      const __nv_bfloat16* yrow = p.bs.y + (static_cast<int64_t>(img) * p.H + hh) * p.Wp * 64 + ch * 8;
#pragma unroll
      for (int s = 0; s < kSteps; ++s) {
        const int col = 4 * s + rsub;
        dst[s] = make_uint4(0u, 0u, 0u, 0u);
        if (hh < p.H && col < kPcValidW && ww0 + col < p.Wp)
          dst[s] = __ldg(reinterpret_cast<const uint4*>(yrow + static_cast<int64_t>(ww0 + col) * 64));
      }
    };
    // one tile ahead: the loads of tile it + 1 are in flight while tile it is consumed (a whole tile period, ~1.6k
    // cycles, against ~1k cycles of HBM latency; issued and consumed inside the same tile they sat on the hand-over path
    // to the epilogue groups and the data gradient ran 2x slower)
    uint4 ya[kSteps], yb[kSteps];
    if (bwd && n_tiles > 0) load_y(n_img, th, tw, ya);
    // one tile: consume `cur` (the y words loaded one tile ago), load the next tile's into `nxt`.  The two buffers swap
    // roles by calling this twice per loop trip -- a register copy nxt -> cur at the end of the tile would wait for the
    // loads it is supposed to overlap.
    auto one_tile = [&](int it, uint4 (&cur)[kSteps], uint4 (&nxt)[kSteps]) {
      const int g = it % NG;
      const uint8_t* stg = staging + g * Cfg::kStageBufBytes;
      const int h0 = th * kPcTH, w0 = tw * kPcValidW;
      if (n_img != acc_img) {
        if (acc_img >= 0) flush(acc_img);
        acc_img = n_img;
        if (bwd) {  // folded affine of this lane's 8 channels (chunk ch = pixel parity ch >> 2, channel octet ch & 3)
          const float4* a4 = reinterpret_cast<const float4*>(p.bs.a + n_img * kPcC + ((ch & 3) << 3));
          const float4* b4 = reinterpret_cast<const float4*>(p.bs.b + n_img * kPcC + ((ch & 3) << 3));
          const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
          pa[0] = a0.x; pa[1] = a0.y; pa[2] = a0.z; pa[3] = a0.w; pa[4] = a1.x; pa[5] = a1.y; pa[6] = a1.z; pa[7] = a1.w;
          pb[0] = b0.x; pb[1] = b0.y; pb[2] = b0.z; pb[3] = b0.w; pb[4] = b1.x; pb[5] = b1.y; pb[6] = b1.z; pb[7] = b1.w;
        }
      }
      const bool row_in = (h0 + q) < p.H;
      // coordinates of the next tile
      int th_n = th + 1, tw_n = tw, img_n = n_img;
      if (th_n == p.tiles_h) {
        th_n = 0;
        if (++tw_n == p.tiles_w) {
          tw_n = 0;
          ++img_n;
        }
      }
      if (bwd && it + 1 < n_tiles) load_y(img_n, th_n, tw_n, nxt);
      named_bar_sync(5 + g, 256);
      {
        // warp q sums staging rows [30q, 30q + 30) (its own image row), four rows per step: one 16-byte chunk
        // (8 channels of one pixel of the pair) per lane; only pairs inside the image count (W is even: a pair is
        // inside or outside as a whole)
#pragma unroll
        for (int s = 0; s < kSteps; ++s) {
          const int col = 4 * s + rsub;
          const int r = q * kPcValidW + col;
          uint4 w = make_uint4(0u, 0u, 0u, 0u);
          if (row_in && col < kPcValidW && w0 + col < p.Wp)
            w = lds_v4(smem_u32(stg) + r * 128 + (((ch ^ (r & 7)) & 7) << 4));
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
          if (bwd) {
            const uint32_t yy[4] = {cur[s].x, cur[s].y, cur[s].z, cur[s].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float x0 = __uint_as_float(ww[i] << 16), x1 = __uint_as_float(ww[i] & 0xffff0000u);
              const float y0 = __uint_as_float(yy[i] << 16), y1 = __uint_as_float(yy[i] & 0xffff0000u);
              const float g0 = x0 * (fmaf(pa[2 * i], y0, pb[2 * i]) > 0.f ? 1.f : p.bs.slope);
              const float g1 = x1 * (fmaf(pa[2 * i + 1], y1, pb[2 * i + 1]) > 0.f ? 1.f : p.bs.slope);
              s1[2 * i] += g0;
              s2[2 * i] = fmaf(g0, y0, s2[2 * i]);
              s1[2 * i + 1] += g1;
              s2[2 * i + 1] = fmaf(g1, y1, s2[2 * i + 1]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float x0 = __uint_as_float(ww[i] << 16), x1 = __uint_as_float(ww[i] & 0xffff0000u);
              s1[2 * i] += x0;
              s2[2 * i] = fmaf(x0, x0, s2[2 * i]);
              s1[2 * i + 1] += x1;
              s2[2 * i + 1] = fmaf(x1, x1, s2[2 * i + 1]);
            }
          }
        }
      }
      if (it + NG < n_tiles) named_bar_arrive(7 + g, 256);  // the group may overwrite its staging buffer
      th = th_n;
      tw = tw_n;
      n_img = img_n;
    };
    for (int it = 0; it < n_tiles; it += 2) {
      one_tile(it, ya, yb);
      if (it + 1 < n_tiles) one_tile(it + 1, yb, ya);
    }
    if (acc_img >= 0) flush(acc_img);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kPcMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// -------------------------------------------------------------------------------------------------- host side
struct PConvGrid {
  int grid, tiles_per_cta, stat_slots, tiles_w, tiles_h;
};

static PConvGrid pconv_grid(int N, int H, int W) {
  PConvGrid g;
  g.tiles_w = ceil_div(W / 2, kPcValidW);
  g.tiles_h = ceil_div(H, kPcTH);
  const long long per_img = static_cast<long long>(g.tiles_w) * g.tiles_h;
  const long long total = per_img * N;
  g.tiles_per_cta = static_cast<int>(ceil_div64(total, num_sms()));
  if (g.tiles_per_cta < 1) g.tiles_per_cta = 1;
  g.grid = static_cast<int>(ceil_div64(total, g.tiles_per_cta));
  g.stat_slots = 4 * kPcNG * (static_cast<int>(ceil_div64(per_img, g.tiles_per_cta)) + 1);  // (CTA, group, lane quarter)
  return g;
}

bool pconv_supported(int k_channels, int n_channels, int stride, int W, int64_t src_pitch, int64_t out_pitch) {
  static const bool off = [] { const char* e = getenv("B200UNET_NO_PCONV"); return e && e[0] == '1'; }();  // A/B knob
  return !off && stride == 1 && k_channels == kPcC && n_channels == kPcC && W % 2 == 0 && W >= 128 &&
         src_pitch == kPcC && out_pitch == kPcC;  // dense tensors only: the (W/2, 64) view needs pitch == channels
}

int pconv_stat_slots(int N, int H, int W) { return (W % 2 == 0 && W >= 128) ? pconv_grid(N, H, W).stat_slots : 0; }

// src: [N,H,W,32] dense (x for fprop, dy for dgrad); wpack: [32 N-side][3][3][32 K-side]; out: [N,H,W,32] dense
int pconv_launch(const void* src, const void* wpack, void* out, float* stats, int N, int H, int W, int rev, int stat_slots,
                 cudaStream_t st, const BwdSums* bs) {
  const PConvGrid g = pconv_grid(N, H, W);
  PConvParams p{};
  PConvMaps maps;
  p.N = N;
  p.H = H;
  p.Wp = W / 2;
  p.tiles_w = g.tiles_w;
  p.tiles_h = g.tiles_h;
  p.tiles_per_cta = g.tiles_per_cta;
  p.stat_slots = stat_slots > g.stat_slots ? stat_slots : g.stat_slots;
  p.stats = stats;
  p.w = static_cast<const __nv_bfloat16*>(wpack);
  if (bs) {
    if (!rev || !stats || bs->y_pitch != kPcC)
      return set_error(kErrInvalid, "pconv: the norm-backward sums need the data gradient, a partial buffer and a dense y");
    p.bs = *bs;
  }
  int rc;
  if ((rc = make_act_map(&maps.src, static_cast<const __nv_bfloat16*>(src), 64, N, H, W / 2, 64, 1, 1, 0, 0, 64, kPcTW,
                         kPcPatchRows)))
    return rc;
  if ((rc = make_act_map(&maps.out, static_cast<const __nv_bfloat16*>(out), 64, N, H, W / 2, 64, 1, 1, 0, 0, 64, kPcValidW,
                         kPcTH)))
    return rc;
  auto kern = rev ? pconv_kernel<true> : pconv_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[rev ? 1 : 0]) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PConvCfg::kSmemBytes));
    attr_set[rev ? 1 : 0] = true;
  }
  if (p.stats)
    B200_CUDA(cudaMemsetAsync(p.stats, 0, static_cast<size_t>(p.N) * p.stat_slots * kPcC * 2 * sizeof(float), st));
  launch_k(kern, dim3(g.grid), dim3(PConvCfg::kThreads), PConvCfg::kSmemBytes, st, maps, p);
  B200_LAUNCH_CHECK("pconv_kernel");
  return 0;
}

}  // namespace b200
