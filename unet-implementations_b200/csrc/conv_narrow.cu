// 3x3 / stride-1 convolutions with a NARROW output (N = Cout for fprop, N = Cin for dgrad, N in {32, 64}) on
// tcgen05 / TMEM: the 512^2 and 256^2 levels of the UNet.  Same role as conv_fprop_dgrad.cu (nn.Conv2d forward and the
// data-gradient half of aten::convolution_backward, Our_UNet/models/unet.py:106-115); dispatched from there.
//
// Why a second kernel.  On B200 a tcgen05.mma with M = 128 costs max(~45, N/2) cycles (tools/micro/umma_rate.cu): the
// A operand (128 x 16 bf16 = 4 KB) is re-read from shared memory for every MMA, so N = 32 runs at 36 % and N = 64 at
// 67 % of the tensor pipe, and the single issuing thread spends one MMA + one barrier round trip per tap.  Here the
// three COLUMN taps (kw) are stacked on N instead:
//     E[pixel, (co, kw)] = sum_{kh, ci} X[h + kh - 1, w', ci] * W[co, ci, kh, kw]          (N = 3 * Cout = 96 / 192)
//     out[h, w, co]      = E[(h, w-1), (co, 0)] + E[(h, w), (co, 1)] + E[(h, w+1), (co, 2)]
//   * one A patch of 6 image rows x 32 pixels per channel chunk serves all 9 taps (the kh shift is a 32-row window
//     offset inside the patch, swizzle-aligned), 3x fewer MMAs, each 3x wider;
//   * a tile is 4 image rows x 32 pixels = one warp per row, so the +-1 pixel shift of the epilogue is a warp
//     shuffle; lanes 0 and 31 are halo: a tile produces 4 x 30 outputs and tiles step by 30 columns (94 % of M);
//   * the weights need no second packing: a 4-D tensor map over the packed [Cout][3][3][Cin] matrix delivers the
//     [(co, kw) x BK] tile of one kh directly; the whole slab stays resident in shared memory.
// Persistent CTAs (384 / 512 threads): warps 0..1 = TMA producers, warp 2 = TMEM owner + lean MMA-issuing thread,
// warps 4.. = NG epilogue groups of four warps (3 for Cout = 32, 2 for Cout = 64: TMEM holds NG accumulators).  The epilogue (tcgen05.ld, 2 shuffles per channel, pack, staging, statistics)
// is ~800 instructions per tile and warp with one warp per scheduler, i.e. several times the 6..12 MMAs of a tile
// (ncu: tensor pipe 12 % busy with a single group), so consecutive tiles alternate between the groups: group g owns
// TMEM accumulator buffer g, staging buffer g and its own named barriers / bulk-store groups.
#include "common.cuh"
#include "ptx.cuh"
#include "conv_common.cuh"
#include <stdlib.h>

namespace b200 {

constexpr int kNcTH = 4, kNcTW = 32, kNcValidW = 30;  // tile rows / columns (lanes) / valid output columns
constexpr int kNcPatchRows = kNcTH + 2;
constexpr int kNcProducers = 2, kNcMmaWarp = 2, kNcEpiWarp0 = 4;
constexpr int kNcMaxGroups = 3;

struct NConvParams {
  int N, H, W, tiles_w, tiles_h;
  int cin;   // K-side channels (multiple of BK)
    int tiles_per_cta;
  int stat_slots;
  float* stats;
  long long* debug;  // optional [gridDim.x][8] cycle counters (developer instrumentation; NULL in production)
  BwdSums bs;        // data gradient only: bs.y != NULL makes the statistics pass the norm-backward reduction
};

struct NConvMaps {
  CUtensorMap src;  // box (BK, 32, 6, 1)
  CUtensorMap w;    // 4-D (K-side channel, kw, kh, N-side channel), box (BK, 3, 1, CO)
  CUtensorMap out;  // box (CO, 30, 4, 1)
};

template <int BK, int CO, int A_SLOTS, int NG = 2>
struct NConvCfg {
  static constexpr int kThreads = 32 * (kNcEpiWarp0 + 4 * NG);
  static constexpr int kRowBytes = BK * 2;
  static constexpr int kASlotBytes = kNcPatchRows * kNcTW * kRowBytes;  // 24 KB (BK = 64) / 12 KB
  static constexpr int kWin16 = (kNcTW * kRowBytes) >> 4;               // one image row of the patch, 16-byte units
  static constexpr int kNeff = 3 * CO;
  static constexpr int kBTileBytes = kNeff * kRowBytes;                 // [(co, kw) x BK] of one kh: 6 / 24 KB
  static constexpr int kBResBytes = 80 * 1024;
  static constexpr int kStageBufBytes = ((kNcTH * kNcValidW * CO * 2 + 1023) / 1024) * 1024;
  static constexpr int kSmemBytes = A_SLOTS * kASlotBytes + kBResBytes + NG * kStageBufBytes + 1024;
  static constexpr uint32_t kSwz = (BK == 64) ? kSwz128 : kSwz64;
  static constexpr uint32_t kSbo = 8 * kRowBytes;
  static constexpr uint32_t kTmemCols = (NG * kNeff <= 256) ? 256 : 512;
  static_assert(NG * kNeff <= 512 && NG <= kNcMaxGroups, "accumulators must fit in TMEM");
  static_assert(kBTileBytes % 1024 == 0, "weight tiles must keep the swizzle alignment");
  static_assert(A_SLOTS % kNcProducers == 0, "each producer owns a fixed subset of slots");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// REV = false (fprop): patch window kh pairs with weight row kh, column tap kw of channel co at accumulator column
// 3*co + kw.  REV = true (dgrad): both reversed (window kh pairs with weight row 2 - kh, the taps swap sides).
template <int BK, int CO, int A_SLOTS, bool REV, int NG>
__global__ void __launch_bounds__(32 * (kNcEpiWarp0 + 4 * NG), 1) nconv_kernel(const __grid_constant__ NConvMaps maps,
                                                                                 const __grid_constant__ NConvParams p) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel of the stream may become resident as CTAs retire
  using Cfg = NConvCfg<BK, CO, A_SLOTS, NG>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[A_SLOTS], a_empty[A_SLOTS];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t tmem_full_bar[NG], tmem_empty_bar[NG];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + A_SLOTS * Cfg::kASlotBytes;
  uint8_t* staging = smem_b + Cfg::kBResBytes;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int total_tiles = tiles_per_img * p.N;
  const int chunks = (p.cin + BK - 1) / BK;  // the last chunk may be partly out of range: TMA zero-fills it, the MMAs skip it
  const int tile_lo = min(static_cast<int>(blockIdx.x) * p.tiles_per_cta, total_tiles);
  const int tile_hi = min(tile_lo + p.tiles_per_cta, total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_SLOTS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(&b_full, 1);
    for (int b = 0; b < NG; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == kNcMmaWarp) {
    tmem_alloc(&tmem_base_holder, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  pdl_wait();  // barrier init / TMEM allocation above ran under the previous kernel's tail; global memory from here on

  if (warp < kNcProducers) {
    if (elect_one()) {
      if (warp == 0) {
        // resident weights: tile (c, kh) = [(n, kw) x BK] of weight row kh and K chunk c
        tma_prefetch_desc(&maps.w);
        mbar_expect_tx(&b_full, static_cast<uint32_t>(chunks) * 3 * Cfg::kBTileBytes);
        for (int c = 0; c < chunks; ++c)
          for (int kh = 0; kh < 3; ++kh)
            tma_load_4d(smem_b + (c * 3 + kh) * Cfg::kBTileBytes, &maps.w, &b_full, c * BK, 0, kh, 0);
      }
      tma_prefetch_desc(&maps.src);
      long long dbg_wait = 0, dbg_issue = 0;
      int it = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        const int n_img = tile / tiles_per_img;
        const int t_in = tile - n_img * tiles_per_img;
        // column-major tile order inside an image: the next tile is the one below, its 2 halo rows are in L2
        const int tw_ = t_in / p.tiles_h;
        const int h0 = (t_in - tw_ * p.tiles_h) * kNcTH, w0 = tw_ * kNcValidW;
        const long long g0 = static_cast<long long>(it) * chunks;
        int c = (warp - static_cast<int>(g0 % kNcProducers) + kNcProducers) % kNcProducers;
        for (; c < chunks; c += kNcProducers) {
          const long long g = g0 + c;
          const int slot = static_cast<int>(g % A_SLOTS);
          const uint32_t ph = static_cast<uint32_t>((g / A_SLOTS) & 1);
          const long long t0 = (kInstr && p.debug) ? clock64() : 0;
          mbar_wait(&a_empty[slot], ph ^ 1);
          const long long t1 = (kInstr && p.debug) ? clock64() : 0;
          mbar_expect_tx(&a_full[slot], Cfg::kASlotBytes);
          tma_load_4d(smem + slot * Cfg::kASlotBytes, &maps.src, &a_full[slot], c * BK, w0 - 1, h0 - 1, n_img);
          if (kInstr && p.debug) {
            dbg_wait += t1 - t0;
            dbg_issue += clock64() - t1;
          }
        }
      }
      if (kInstr && p.debug && warp == 0) {
        p.debug[blockIdx.x * 8 + 0] = dbg_wait;
        p.debug[blockIdx.x * 8 + 1] = dbg_issue;
      }
    }
  } else if (warp == kNcMmaWarp) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, Cfg::kNeff, 0, 0);
      constexpr uint32_t hi = umma_desc_hi(Cfg::kSbo, Cfg::kSwz);
      constexpr uint32_t kASlot16 = Cfg::kASlotBytes >> 4, kBTile16 = Cfg::kBTileBytes >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem), 16);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b), 16);
      const uint32_t a_full0 = smem_u32(&a_full[0]), a_empty0 = smem_u32(&a_empty[0]);
      mbar_wait(&b_full, 0);
      tc_fence_after();
      uint32_t aslot = 0, aph = 0;
      long long dbg_te = 0, dbg_af = 0;
      const long long dbg_m0 = kInstr ? clock64() : 0;
      int it = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        const int buf = it % NG;  // accumulator (= epilogue group) of this tile
        const long long t0 = (kInstr && p.debug) ? clock64() : 0;
        mbar_wait(&tmem_empty_bar[buf], static_cast<uint32_t>(((it / NG) & 1) ^ 1));
        if (kInstr && p.debug) dbg_te += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * Cfg::kNeff;
        uint32_t acc = 0;
        uint32_t b_c = b_lo0;
        for (int c = 0; c < chunks; ++c, b_c += 3 * kBTile16) {
          const long long t1 = (kInstr && p.debug) ? clock64() : 0;
          mbar_wait_u32(a_full0 + aslot * 8, aph);
          if (kInstr && p.debug) dbg_af += clock64() - t1;
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + aslot * kASlot16;
          const int ksteps = min(BK, p.cin - c * BK) >> 4;  // K = 16 steps of this chunk that hold real channels
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            // window kh of the patch pairs with weight row kh (fprop) or 2 - kh (dgrad)
            const uint32_t b_lo = b_c + (REV ? (2 - kh) : kh) * kBTile16;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (k < ksteps) {
                umma_bf16_lean(d_tmem, a_lo + kh * Cfg::kWin16 + 2 * k, hi, b_lo + 2 * k, hi, idesc, acc);
                acc = 1;
              }
            }
          }
          umma_commit_u32(a_empty0 + aslot * 8);
          if (++aslot == A_SLOTS) { aslot = 0; aph ^= 1; }
        }
        umma_commit(&tmem_full_bar[buf]);
      }
      if (kInstr && p.debug) {
        p.debug[blockIdx.x * 8 + 2] = dbg_te;
        p.debug[blockIdx.x * 8 + 3] = dbg_af;
        p.debug[blockIdx.x * 8 + 4] = clock64() - dbg_m0;
      }
    }
  } else if (warp >= kNcEpiWarp0) {
    // ------------------------------------------------ epilogue group g (warps 4..7 / 8..11): tiles it = g, g+2, ...;
    // warp q of the group = TMEM lane quarter q = tile row q
    const int g = (warp - kNcEpiWarp0) >> 2;
    const int q = warp & 3;
    const int et = threadIdx.x - 32 * (kNcEpiWarp0 + 4 * g);  // 0..127 inside the group
    const int bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    const bool do_stats = p.stats != nullptr;
    const bool col_ok = lane >= 1 && lane <= kNcValidW;
    const int srow = q * kNcValidW + lane - 1;  // staging row of this thread's output pixel
    uint8_t* stg = staging + g * Cfg::kStageBufBytes;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * Cfg::kNeff;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};       // [s1 c0, s2 c0, s1 c1, s2 c1] of this lane's channel pair
    // data gradient with p.bs.y: the statistics pass becomes the norm-backward reduction of the consuming unit -- the
    // staged values are dz, the matching words of that unit's raw output y are loaded from global memory at the START of
    // the tile (kYSteps loads in flight per lane, consumed after the TMA store has been issued) and the sums are
    // (sum gm, sum gm * y) with gm = dz * lrelu'(a*y + b)
    const bool bwd = REV && p.bs.y != nullptr;
    constexpr int kYSteps = (CO == 64) ? kNcValidW : kNcValidW / 2;
    const uint32_t ycp = (CO == 64) ? static_cast<uint32_t>(lane) : (static_cast<uint32_t>(lane) & 15);  // channel pair
    const uint32_t ypar = (CO == 64) ? 0u : (static_cast<uint32_t>(lane) >> 4);
    float pa0 = 0.f, pa1 = 0.f, pb0 = 0.f, pb1 = 0.f;
    int acc_img = -1;
    auto flush = [&](int img) {
      const int first_tile = img * tiles_per_img;
      const int b0 = first_tile / p.tiles_per_cta;
      const int slot = ((static_cast<int>(blockIdx.x) - b0) * NG + g) * 4 + q;
      float* dst = p.stats + (static_cast<size_t>(img) * p.stat_slots + slot) * CO * 2;
      if (CO == 64) {
        *reinterpret_cast<float4*>(dst + (2 * lane) * 2) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else {
        const float a0 = acc[0] + __shfl_down_sync(0xffffffffu, acc[0], 16);
        const float a1 = acc[1] + __shfl_down_sync(0xffffffffu, acc[1], 16);
        const float a2 = acc[2] + __shfl_down_sync(0xffffffffu, acc[2], 16);
        const float a3 = acc[3] + __shfl_down_sync(0xffffffffu, acc[3], 16);
        if (lane < 16) *reinterpret_cast<float4*>(dst + (2 * lane) * 2) = make_float4(a0, a1, a2, a3);
      }
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    };
    // tile coordinates advance incrementally (NG tiles per step): no divisions in the loop
    int tile = tile_lo + g;
    int n_img = tile / tiles_per_img;
    int tw = (tile - n_img * tiles_per_img) / p.tiles_h;
    int th = tile - n_img * tiles_per_img - tw * p.tiles_h;
    uint32_t ph = 0;  // parity of this group's accumulator barrier
    long long dbg_tf = 0, dbg_sw = 0;
    const long long dbg_e0 = kInstr ? clock64() : 0;
    for (; tile < tile_hi; tile += NG, ph ^= 1) {
      const int h0 = th * kNcTH, w0 = tw * kNcValidW;
      if (do_stats && n_img != acc_img) {
        if (acc_img >= 0) flush(acc_img);
        acc_img = n_img;
        if (bwd) {
          const float2 a2 = __ldg(reinterpret_cast<const float2*>(p.bs.a + n_img * CO) + ycp);
          const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bs.b + n_img * CO) + ycp);
          pa0 = a2.x; pa1 = a2.y; pb0 = b2.x; pb1 = b2.y;
        }
      }
      uint32_t yw[REV ? kYSteps : 1];
      if (REV && bwd) {
        const bool row_ok = (h0 + q) < p.H;
        const __nv_bfloat16* yrow = p.bs.y + (static_cast<int64_t>(n_img) * p.H + (h0 + q)) * p.W * p.bs.y_pitch + 2 * ycp;
#pragma unroll
        for (int i = 0; i < kYSteps; ++i) {
          const int col = (CO == 64) ? i : (2 * i + static_cast<int>(ypar));
          yw[REV ? i : 0] = 0u;
          if (row_ok && w0 + col < p.W)
            yw[REV ? i : 0] = __ldg(reinterpret_cast<const uint32_t*>(yrow + static_cast<int64_t>(w0 + col) * p.bs.y_pitch));
        }
      }
      const long long t0 = (kInstr && p.debug) ? clock64() : 0;
      mbar_wait(&tmem_full_bar[g], ph);
      const long long t1 = (kInstr && p.debug) ? clock64() : 0;
      tc_fence_after();
      if (et == 0) tma_store_wait_read_all();  // this group's previous store has read the staging buffer
      named_bar_sync(bar_a, 128);
      if (kInstr && p.debug) {
        dbg_tf += t1 - t0;
        dbg_sw += clock64() - t1;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < CO; c0 += 16) {
        // 16 output channels = 48 accumulator columns (co, kw) starting at 3 * c0 (48 live registers: the kernel has to
        // fit 128 registers per thread with three epilogue groups)
        uint32_t v[48];
        tmem_ld_32x16(t_addr + 3 * c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld_32x16(t_addr + 3 * c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        tmem_ld_32x16(t_addr + 3 * c0 + 32, *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
        tmem_ld_wait();
        if (c0 + 16 >= CO) {
          // all TMEM reads of this warp are done: hand the accumulator back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[g]);
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float o[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * j + e;
            // the left neighbour's tap for this pixel is its kw = 0 column (fprop order), the right neighbour's kw = 2
            const float from_left = __shfl_up_sync(0xffffffffu, __uint_as_float(REV ? v[3 * c + 2] : v[3 * c]), 1);
            const float from_right = __shfl_down_sync(0xffffffffu, __uint_as_float(REV ? v[3 * c] : v[3 * c + 2]), 1);
            o[e] = __uint_as_float(v[3 * c + 1]) + from_left + from_right;
          }
          pk[j] = pack_bf16x2(o[0], o[1]);
        }
        if (col_ok) {
          // two 16-byte chunks (8 channels each) of this pixel's staging row, swizzled like the TMA store box
          const uint32_t base = smem_u32(stg) + srow * (CO * 2);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int cj = (c0 >> 3) + i;
            const uint32_t addr = (CO == 64) ? base + (((cj ^ (srow & 7)) & 7) << 4)
                                             : base + (((cj ^ ((srow >> 1) & 3)) & 3) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                         "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                         : "memory");
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(bar_b, 128);
      if (et == 0) {
        tma_store_4d(&maps.out, stg, 0, w0, h0, n_img);
        tma_store_commit();
      }
      if (do_stats) {
        // warp q sums staging rows [30q, 30q + 30) (its own image row): lane = channel pair (CO = 64) or
        // (row parity, channel pair) (CO = 32); only pixels inside the image count
        constexpr int kSteps = (CO == 64) ? kNcValidW : kNcValidW / 2;
        const uint32_t cp = (CO == 64) ? static_cast<uint32_t>(lane) : (static_cast<uint32_t>(lane) & 15);
        const uint32_t par = (CO == 64) ? 0u : (static_cast<uint32_t>(lane) >> 4);
        const uint32_t cw = cp >> 2, wi = (cp & 3) << 2;
        const bool row_in = (h0 + q) < p.H;
        float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
        for (int i0 = 0; i0 < kSteps; i0 += 5) {
          uint32_t w[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const int col = (CO == 64) ? (i0 + i) : (2 * (i0 + i) + static_cast<int>(par));
            const int r = q * kNcValidW + col;
            const uint32_t off = (CO == 64) ? (r * 128 + (((cw ^ (r & 7)) & 7) << 4) + wi)
                                            : (r * 64 + (((cw ^ ((r >> 1) & 3)) & 3) << 4) + wi);
            w[i] = lds_u32(smem_u32(stg) + off);
            if (!row_in || w0 + col >= p.W) w[i] = 0;
          }
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float x0 = __uint_as_float(w[i] << 16), x1 = __uint_as_float(w[i] & 0xffff0000u);
            if (REV && bwd) {
              const uint32_t yy = yw[REV ? (i0 + i) : 0];
              const float y0 = __uint_as_float(yy << 16), y1 = __uint_as_float(yy & 0xffff0000u);
              const float g0 = fmaf(pa0, y0, pb0) > 0.f ? x0 : x0 * p.bs.slope;
              const float g1 = fmaf(pa1, y1, pb1) > 0.f ? x1 : x1 * p.bs.slope;
              s1a += g0;
              s2a = fmaf(g0, y0, s2a);
              s1b += g1;
              s2b = fmaf(g1, y1, s2b);
            } else {
              s1a += x0;
              s2a = fmaf(x0, x0, s2a);
              s1b += x1;
              s2b = fmaf(x1, x1, s2b);
            }
          }
        }
        acc[0] += s1a;
        acc[1] += s2a;
        acc[2] += s1b;
        acc[3] += s2b;
      }
      // next tile of this group
      th += NG;
      while (th >= p.tiles_h) {
        th -= p.tiles_h;
        if (++tw == p.tiles_w) {
          tw = 0;
          ++n_img;
        }
      }
    }
    if (do_stats && acc_img >= 0) flush(acc_img);
    if (et == 0) tma_store_wait_all();
    if (kInstr && p.debug && et == 0 && g == 0) {
      p.debug[blockIdx.x * 8 + 5] = dbg_tf;
      p.debug[blockIdx.x * 8 + 6] = clock64() - dbg_e0;
      p.debug[blockIdx.x * 8 + 7] = dbg_sw;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kNcMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// -------------------------------------------------------------------------------------------------- host side
struct NConvGrid {
  int grid, tiles_per_cta, stat_slots, tiles_w, tiles_h;
};

static NConvGrid nconv_grid(int N, int H, int W) {
  NConvGrid g;
  g.tiles_w = ceil_div(W, kNcValidW);
  g.tiles_h = ceil_div(H, kNcTH);
  const long long per_img = static_cast<long long>(g.tiles_w) * g.tiles_h;
  const long long total = per_img * N;
  g.tiles_per_cta = static_cast<int>(ceil_div64(total, num_sms()));
  if (g.tiles_per_cta < 1) g.tiles_per_cta = 1;
  g.grid = static_cast<int>(ceil_div64(total, g.tiles_per_cta));
  g.stat_slots = 4 * kNcMaxGroups * (static_cast<int>(ceil_div64(per_img, g.tiles_per_cta)) + 1);  // (CTA, group, lane quarter)
  return g;
}

// K chunk.  64 channels = 128-byte operand rows run the MMAs at full rate; K-major operands with 64-byte rows (64 B
// swizzle) were measured at ~120 cycles per N = 96 MMA instead of 64.  For 96 channels the 32-channel tail therefore
// also uses a 64-channel box: the TMA zero-fills the channels past k_channels and the MMA loop skips their K = 16
// steps (d4c1 fprop 632 -> 596 us).  For 32 channels the doubled shared-memory fill costs more than the faster MMAs
// save (313 -> 345 us), so those layers keep 64-byte rows.
static int nconv_bk(int k_channels) { return k_channels > 32 ? 64 : 32; }

bool nconv_supported(int k_channels, int n_channels, int stride, int W) {
  if (stride != 1 || (n_channels != 32 && n_channels != 64) || k_channels % 32 != 0 || k_channels <= 0) return false;
  if (W < 64) return false;  // the 30-of-32 column tiling only pays on wide images
  {  // developer knob: B200UNET_NO_NCONV=<n_channels> sends that output width to gconv_kernel instead (A/B)
    const char* e = getenv("B200UNET_NO_NCONV");
    if (e && atoi(e) == n_channels) return false;
  }
  const int BK = nconv_bk(k_channels);
  return static_cast<long long>(ceil_div(k_channels, BK)) * 3 * (3 * n_channels * BK * 2) <= 80 * 1024;  // resident weights
}

int nconv_stat_slots(int N, int H, int W) {
  const int a = nconv_grid(N, H, W).stat_slots, b = pconv_stat_slots(N, H, W);  // either kernel may run for Cout = 32
  return a > b ? a : b;
}

static long long* nconv_debug_buffer() {
  static long long* buf = nullptr;
  static int state = -1;
  if (state < 0) {
    const char* e = getenv("B200UNET_GCONV_DEBUG");
    state = (e && e[0] == '1') ? 1 : 0;
    if (state) cudaMalloc(&buf, 148 * 8 * sizeof(long long));
  }
  return state ? buf : nullptr;
}

template <int BK, int CO, int A_SLOTS, bool REV, int NG>
static int launch_nconv(const NConvMaps& maps, NConvParams& p, const NConvGrid& g, cudaStream_t st) {
  // p.stat_slots = P of the caller's buffer (>= this kernel's own slot count); unused slots stay zero
  using Cfg = NConvCfg<BK, CO, A_SLOTS, NG>;
  auto kern = nconv_kernel<BK, CO, A_SLOTS, REV, NG>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  if (p.stats)
    B200_CUDA(cudaMemsetAsync(p.stats, 0, static_cast<size_t>(p.N) * p.stat_slots * CO * 2 * sizeof(float), st));
  p.debug = nconv_debug_buffer();
  if (kInstr && p.debug) cudaMemsetAsync(p.debug, 0, 148 * 8 * sizeof(long long), st);
  launch_k(kern, dim3(g.grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, maps, p);
  B200_LAUNCH_CHECK("nconv_kernel");
  if (kInstr && p.debug) {  // developer instrumentation (B200UNET_GCONV_DEBUG=1): per-tile cycle counts of CTA 0
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.debug, sizeof(h), cudaMemcpyDeviceToHost);
    const long long tiles = g.tiles_per_cta, gt = (tiles + NG - 1) / NG;
    fprintf(stderr,
            "nconv<%d,%d,%d,%d,%d> K %d tiles/CTA %lld | per tile: producer0 wait %lld issue %lld | mma total %lld "
            "wait_tmem_empty %lld wait_a_full %lld | epilogue group 0 per ITS tile: total %lld wait_tmem_full %lld store+bar %lld\n",
            BK, CO, A_SLOTS, (int)REV, NG, p.cin, tiles, h[0] / tiles, h[1] / tiles, h[4] / tiles, h[2] / tiles, h[3] / tiles,
            h[6] / gt, h[5] / gt, h[7] / gt);
  }
  return 0;
}

// src: [N,H,W,K-side channels] (x for fprop, dy for dgrad); wpack: [N-side channels][3][3][K-side channels];
// out: [N,H,W,n_channels] (y / dx); rev = 0 fprop, 1 dgrad
int nconv_launch(const void* src, int64_t src_pitch, const void* wpack, void* out, int64_t out_pitch, float* stats, int N,
                 int H, int W, int k_channels, int n_channels, int rev, int stat_slots, cudaStream_t st, const BwdSums* bs) {
  if (pconv_supported(k_channels, n_channels, 1, W, src_pitch, out_pitch) && (!bs || bs->y_pitch == n_channels))
    return pconv_launch(src, wpack, out, stats, N, H, W, rev, stat_slots, st, bs);
  const int BK = nconv_bk(k_channels);
  const NConvGrid g = nconv_grid(N, H, W);
  NConvParams p{};
  NConvMaps maps;
  p.N = N;
  p.H = H;
  p.W = W;
  p.tiles_w = g.tiles_w;
  p.tiles_h = g.tiles_h;
  p.cin = k_channels;
  p.tiles_per_cta = g.tiles_per_cta;
  p.stat_slots = stat_slots > g.stat_slots ? stat_slots : g.stat_slots;
  p.stats = stats;
  if (bs) {
    if (!rev || !stats || bs->y_pitch % 2 != 0)
      return set_error(kErrInvalid, "nconv: the norm-backward sums need the data gradient and a partial buffer");
    p.bs = *bs;
  }
  int rc;
  if ((rc = make_act_map(&maps.src, static_cast<const __nv_bfloat16*>(src), src_pitch, N, H, W, k_channels, 1, 1, 0, 0, BK,
                         kNcTW, kNcPatchRows)))
    return rc;
  if ((rc = make_act_map(&maps.out, static_cast<const __nv_bfloat16*>(out), out_pitch, N, H, W, n_channels, 1, 1, 0, 0,
                         n_channels, kNcValidW, kNcTH)))
    return rc;
  {
    // weights as a 4-D tensor (K-side channel, kw, kh, N-side channel) over the packed [n][kh][kw][k] matrix
    const uint64_t kc = static_cast<uint64_t>(k_channels);
    uint64_t dims[4] = {kc, 3, 3, static_cast<uint64_t>(n_channels)};
    uint64_t strides[3] = {kc * 2, 3 * kc * 2, 9 * kc * 2};
    uint32_t box[4] = {static_cast<uint32_t>(BK), 3, 1, static_cast<uint32_t>(n_channels)};
    if ((rc = make_tmap_bf16(&maps.w, wpack, 4, dims, strides, box,
                             BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
  }
  // epilogue groups: two everywhere.  Three (possible for Cout = 32: three accumulators of 96 columns, 128 registers)
  // were measured equal to two (301 / 265 us fprop / dgrad at 512^2 x 32): with two groups the epilogue is no longer
  // the bottleneck; the kernel then runs at ~16 B/cycle/SM of L2<->SM traffic (3.9 TB/s of DRAM) for both widths
#define NC(bk, co, as, ng) \
  if (BK == bk && n_channels == co) \
    return rev ? launch_nconv<bk, co, as, true, ng>(maps, p, g, st) : launch_nconv<bk, co, as, false, ng>(maps, p, g, st);
  NC(64, 64, 4, 2) NC(64, 32, 4, 2) NC(32, 64, 8, 2) NC(32, 32, 8, 2)
#undef NC
  return set_error(kErrUnsupported, "no nconv instantiation for BK=%d N=%d", BK, n_channels);
}

}  // namespace b200
