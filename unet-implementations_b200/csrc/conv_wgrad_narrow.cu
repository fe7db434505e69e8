// Weight gradient of the 3x3 / stride-1 convolutions with a NARROW output (Cout = 32 or 64): the 512^2 and 256^2
// levels of the UNet (encoder 0/1, decoder 3/4), where K = N*H*W is in the millions and Cout is too small to fill
// the tensor-core array (a tcgen05.mma costs max(64, N/2) cycles on B200: N = 32 runs at 25 % -- DESIGN.md 3).
// Replaces the weight-gradient half of aten::convolution_backward for nn.Conv2d (Our_UNet/models/unet.py:106-115).
//
// Formulation.  dW[co, kh, kw, ci] = sum_{n,oh,w'} X[n, oh+kh-1, w', ci] * dY[n, oh, w'-kw+1, co]   (zero outside).
//   The kh shift is put on X and the kw shift on dY, so that each operand is loaded ONCE per pixel block instead of
//   once per tap, and both the M and the N side of the MMA are widened by a factor 3:
//     A (MN-major, straight from NHWC): one X patch of (R+3) image rows x 16 pixels per channel chunk.  The M = 128
//        rows of an MMA are 128/CH "kh slots" of CH channels; slot j is the same patch advanced by j image rows, so
//        the slots are addressed by the descriptor's leading-dimension stride -- no extra loads.
//     B (MN-major): three dY tiles of R rows x 16 pixels, shifted by kw-1 pixels (three cheap TMA loads of the narrow
//        tensor), laid out back to back so that N = 3*Cout = (kw, co).
//     D[(kh, ci), (kw, co)] += A^T B accumulates in TMEM over the CTA's whole pixel range (split-K over the grid);
//        one K = 16 MMA per image row of the block.  fp32 partials per split are reduced in fixed order by
//        wgrad_finalize_kernel (deterministic).
// L2->SM traffic per 128-pixel block drops from 9 X loads + 1 dY load to 1.4 X loads + 3 dY loads, and the MMA count
// from 9*Cin/32 to Cin/32 (CH = 32) per image row.
#include "common.cuh"
#include "ptx.cuh"
#include "conv_common.cuh"

namespace b200 {

struct WgradNParams {
  int N, OH, OW, blocks_w, blocks_h;
  int total_kb, kb_per_split;
  int cin, cout;
  int chunks_per_cta;  // CPB: channel chunks (of CH) whose accumulators live in this CTA's TMEM
  int stages;
  int col_major;       // block order inside an image: 1 = down the columns (the 3 halo rows of the next block are in L2)
  float* partial;      // [S][9][cin][cout]
};

struct WgradNMaps {
  CUtensorMap x;   // box (CH, 16, R+3, 1)
  CUtensorMap dy;  // box (CO, 16, R, 1)
};

constexpr int kWnR = 8;         // image rows per pixel block (K = 16*R = 128 pixels per stage)
// MODE >= 1: pixel blocks of kWnR1 rows x kWnW1 pixels.  Measured with 4 x 32 (4 KB instead of 2 KB contiguous per TMA
// row, 7/4 instead of 11/8 halo rows): 216-236 against 226 us (32 -> 32 pairs), 165 against 167 (64 -> 64), 421 against
// 429 us (192 -> 64): no difference, the block shape is not what bounds the loads (DRAM reads at 5.3 TB/s are).
constexpr int kWnR1 = 8, kWnW1 = 16;
constexpr int kWnMaxStages = 6;  // the loop is bound by load latency x bytes in flight (ncu: 47 % of samples on the
                                 // full barrier, DRAM 45 %, tensor 31 % with 4 stages of 35 KB): use all of shared memory

template <int CH, int CO>
struct WnCfg {
  static constexpr int kSlots = 128 / CH;              // kh slots per MMA (4 or 2)
  static constexpr int kGroups = (3 + kSlots - 1) / kSlots;  // MMAs per (row, chunk): 1 (CH=32) or 2 (CH=64)
  static constexpr int kRowA = CH * 2, kRowB = CO * 2;  // bytes per pixel row in smem
  static constexpr int kXBytes = (kWnR + 3) * 16 * kRowA;  // one chunk's patch
  static constexpr int kDyBytes = kWnR * 16 * kRowB;        // one shifted dY tile
  static constexpr int kNcols = 3 * CO;                     // accumulator columns per MMA group
  // ONE: a single (16 + 2)-pixel wide dY tile; the three kw shifts are three N units ONE 128-byte ROW apart (LBO = one
  // row) -- the swizzle is a function of the absolute shared-memory address, so a start address that is a multiple of
  // 128 (or, MODE 2, of 16) but not of 1024 bytes reads what TMA wrote: results bit-identical to the three-tile form,
  // measured; with the descriptor's base-offset field set to (address >> 7) & 7 they are wrong, so it stays 0.
  // 18 KB instead of 48 KB per stage for Cout = 64.
  static constexpr int kXBytes1 = (kWnR1 + 3) * kWnW1 * kRowA;
  static constexpr int kDyTx1 = kWnR1 * (kWnW1 + 2) * kRowB;
  static constexpr int kDyBytes1 = ((kDyTx1 + 1023) / 1024) * 1024;
  static constexpr uint32_t kSwzA = (CH == 64) ? kSwz128 : kSwz64;
  static constexpr uint32_t kSwzB = (CO == 64) ? kSwz128 : kSwz64;
};

// MODE 0: three shifted dY tiles.  MODE 1 (ONE): one wide dY tile, three N units one row apart.  MODE 2 (PAIR, on the
// pixel-pair views of 32-channel tensors, see wgradn_launch_pairs): the wide tile again, but N = 128 contiguous
// elements starting HALF a row in -- [odd pixel of pair q-1 | pair q | even pixel of pair q+1] = the four pixels
// 2q-1 .. 2q+2 that the two pixels of X pair q meet under the three kw shifts: 6 of the 8 (px, pixel) blocks are real
// products instead of 6 of 12, the MMA costs 64 instead of 96 cycles and reads 8 instead of 10 KB of shared memory.
template <int CH, int CO, int MODE = 0>
__global__ void __launch_bounds__(kConvThreads, 1) wgradn_kernel(const __grid_constant__ WgradNMaps maps,
                                                                  const __grid_constant__ WgradNParams p) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel of the stream may become resident as CTAs retire
  using Cfg = WnCfg<CH, CO>;
  constexpr bool ONE = MODE >= 1, PAIR = MODE == 2;
  static_assert(!PAIR || (CH == 64 && CO == 64), "the pair form runs on the 64-channel pair views");
  constexpr int kNcols = PAIR ? 128 : Cfg::kNcols;  // accumulator columns per MMA group
  constexpr int R = ONE ? kWnR1 : kWnR, BW = ONE ? kWnW1 : 16;  // pixel block: rows x pixels
  constexpr int kXBytes = ONE ? Cfg::kXBytes1 : Cfg::kXBytes;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWnMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWnMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int CPB = p.chunks_per_cta;
  const int STAGES = p.stages;
  constexpr int kNB = ONE ? 1 : 3;                                          // dY loads per stage
  constexpr int kDyTotal = ONE ? Cfg::kDyBytes1 : 3 * Cfg::kDyBytes;
  const int stage_bytes = ONE ? ((CPB * kXBytes + kDyTotal + 1023) / 1024) * 1024 : CPB * kXBytes + kDyTotal;
  const int split = blockIdx.x;
  const int chunk0 = blockIdx.y * CPB;  // first channel chunk of this CTA
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.total_kb);
  const int num_kb = kb_end - kb_begin;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(CPB * Cfg::kGroups * kNcols)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
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

  if (warp < kProducerWarps) {
    if (elect_one()) {
      // every producer walks all pixel blocks and issues loads j = warp, warp+4, ... of each (3 dY tiles, CPB X patches)
      tma_prefetch_desc(&maps.x);
      tma_prefetch_desc(&maps.dy);
      const int blocks_per_img = p.blocks_w * p.blocks_h;
      for (int i = 0; i < num_kb; ++i) {
        const int kb = kb_begin + i;
        const int n_img = kb / blocks_per_img;
        const int b_in = kb - n_img * blocks_per_img;
        // block order: down the columns, so that the 3 halo rows of the next block are still in L2.  Round 1 measured
        // the opposite (381 vs 345 us at 512^2 x 32) when the loads were bound by 64-byte TMA rows; now that these kernels
        // run at the DRAM rate the saved re-reads win: 538 -> 519 us (96 -> 32), 218 -> 216 (32 -> 32 pairs)
        int bh, bw;
        if (p.col_major) {
          bw = b_in / p.blocks_h;
          bh = b_in - bw * p.blocks_h;
        } else {
          bh = b_in / p.blocks_w;
          bw = b_in - bh * p.blocks_w;
        }
        const int h0 = bh * R, w0 = bw * BW;
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint32_t my_bytes = 0;
        for (int j = warp; j < kNB + CPB; j += kProducerWarps)
          my_bytes += (j < kNB) ? (ONE ? Cfg::kDyTx1 : Cfg::kDyBytes) : kXBytes;
        mbar_expect_tx(&full_bar[s], my_bytes);
        uint8_t* sb = smem + s * stage_bytes;
        for (int j = warp; j < kNB + CPB; j += kProducerWarps) {
          if (j < kNB) {
            // B: dY shifted by kw - 1 pixels: B_kw[oh, w'] = dY[oh, w' - kw + 1]   (kw = j); ONE: columns w0-1 .. w0+16
            if (ONE) tma_load_4d(sb, &maps.dy, &full_bar[s], 0, w0 - 1, h0, n_img);
            else tma_load_4d(sb + j * Cfg::kDyBytes, &maps.dy, &full_bar[s], 0, w0 + 1 - j, h0, n_img);
          } else {
            // A: X patch rows [h0 - 1, h0 + R + 2) of channel chunk j - kNB
            tma_load_4d(sb + kDyTotal + (j - kNB) * kXBytes, &maps.x, &full_bar[s],
                        (chunk0 + j - kNB) * CH, w0, h0 - 1, n_img);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (elect_one()) {
      // lean issue loop (see ptx.cuh): 32-bit descriptor halves, image-row advance = an add on the low word
      constexpr uint32_t idesc = umma_idesc_bf16(128, kNcols, 1, 1);
      constexpr uint32_t a_hi = umma_desc_hi(8 * Cfg::kRowA, Cfg::kSwzA), b_hi = umma_desc_hi(8 * Cfg::kRowB, Cfg::kSwzB);
      constexpr uint32_t kRowStepA = (BW * Cfg::kRowA) >> 4;                                      // one image row
      constexpr uint32_t kRowStepB = ((ONE ? BW + 2 : 16) * Cfg::kRowB) >> 4;
      constexpr uint32_t kHalfA = (16 * Cfg::kRowA) >> 4, kHalfB = (16 * Cfg::kRowB) >> 4;  // 16 pixels of K
      // A: leading-dimension stride = one image row (kh slots);  B: leading-dimension stride = one shifted dY tile
      // (ONE: one pixel row of the wide tile; N unit u starts at column u, i.e. holds the shift kw = 2 - u)
      constexpr uint32_t a_lbo = (((BW * Cfg::kRowA) >> 4) & 0x3FFFu) << 16;
      constexpr uint32_t b_lbo = (((ONE ? Cfg::kRowB : Cfg::kDyBytes) >> 4) & 0x3FFFu) << 16;
      const uint32_t lo0 = umma_desc_lo(smem_u32(smem), 0);
      const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
      uint32_t s = 0, ph = 0, acc = 0;
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t b0 = lo0 + s * stage16 + b_lbo + (PAIR ? 4u : 0u);  // PAIR: the window starts 64 bytes into row 0
        const uint32_t x0 = lo0 + s * stage16 + (kDyTotal >> 4) + a_lbo;
#pragma unroll 2
        for (int rh = 0; rh < R * (BW / 16); ++rh) {  // one K = 16 step: 16 pixels of one image row of the block
          const int r = rh / (BW / 16), h = rh % (BW / 16);
          const uint32_t b_lo = b0 + r * kRowStepB + h * kHalfB;
          uint32_t a_lo = x0 + r * kRowStepA + h * kHalfA;
          for (int j = 0; j < CPB; ++j, a_lo += (kXBytes >> 4)) {
#pragma unroll
            for (int g = 0; g < Cfg::kGroups; ++g)
              umma_bf16_lean(tmem_base + (j * Cfg::kGroups + g) * kNcols, a_lo + g * Cfg::kSlots * kRowStepA, a_hi,
                             b_lo, b_hi, idesc, acc | rh);
          }
        }
        acc = 1;
        umma_commit(&empty_bar[s]);
        if (++s == static_cast<uint32_t>(STAGES)) { s = 0; ph ^= 1; }
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: TMEM -> fp32 partials
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int slot = row / CH;
    const int ci_l = row - slot * CH;
    if (num_kb > 0) {
      mbar_wait(&tmem_full_bar, 0);
      tc_fence_after();
    }
    for (int j = 0; j < CPB; ++j) {
      const int ci = (chunk0 + j) * CH + ci_l;
#pragma unroll 1
      for (int g = 0; g < Cfg::kGroups; ++g) {
        const int kh = g * Cfg::kSlots + slot;
#pragma unroll 1
        for (int kw = 0; kw < (PAIR ? 2 : 3); ++kw) {  // PAIR: two 64-column halves of the 128-column window
#pragma unroll 1
          for (int c0 = 0; c0 < CO; c0 += 32) {
            uint32_t v[32];
            if (num_kb > 0) {
              tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (j * Cfg::kGroups + g) * kNcols +
                                kw * CO + c0,
                            v);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int t = 0; t < 32; ++t) v[t] = 0;
            }
            if (kh < 3) {
              // column block kw of a group holds the shift kw (ONE: 2 - kw); the partials are always [kh][shift]
              const int sh = ONE ? 2 - kw : kw;
              // PAIR: partials are [split][kh][(px, ci)][128 window columns = 4 pixels x 32 co]
              float* dst = PAIR ? p.partial + ((static_cast<size_t>(split) * 3 + kh) * 64 + ci) * 128 + kw * 64 + c0
                                : p.partial + ((static_cast<size_t>(split) * 9 + kh * 3 + sh) * p.cin + ci) * p.cout + c0;
              float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
              for (int t = 0; t < 8; ++t)
                d4[t] = make_float4(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1]),
                                    __uint_as_float(v[4 * t + 2]), __uint_as_float(v[4 * t + 3]));
            }
          }
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

// ---------------------------------------------------------------------------------------------- 32 -> 32 on PIXEL PAIRS
// wgradn_kernel<32,32> is bound by its loads: 64-byte operand rows (32 channels) arrive at about half the rate of
// 128-byte rows (ncu: 47 % of the samples on the full barrier, tensor pipe 31 %, DRAM 45 %; 25 B/cycle/SM landing).  A
// DENSE 32-channel NHWC tensor is also a 64-channel tensor with W/2 "pair" pixels (the view pconv_kernel uses), and on
// those views the problem is exactly wgradn_kernel<64,64>: M = (kh slot, px, ci), N = (j, py, co), K = pairs, with
//   D_j[kh][(px,ci),(py,co)] = sum_q Xp[h+kh-1, q][(px,ci)] * dYp[h, q-j+1][(py,co)]     (pixel 2q+px against 2(q-j+1)+py)
// so the product belongs to kw = 2j - 1 + px - py.  Six of the twelve (j, px, py) blocks land on a valid kw, two per kw;
// the finalize kernel below adds them (the MMAs execute 2x the useful FLOPs -- the array is idle anyway: the kernel is
// within reach of the HBM floor instead of bound by half-rate loads).
__global__ void __launch_bounds__(256) wgrad_finalize_pairs_kernel(const float* __restrict__ partial,
                                                                    float* __restrict__ dw, int S) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][32];
  const int co = threadIdx.x & 31, z = threadIdx.x >> 5;
  const int ci = blockIdx.x & 31, tap = blockIdx.x >> 5;  // tap = kh*3 + kw of the OUTPUT
  const int kh = tap / 3, kw = tap - kh * 3;
  // the two (j, px, py) blocks of this kw
  const int j0 = kw == 0 ? 0 : 1, px0 = kw == 0 ? 1 : (kw == 1 ? 0 : 1), py0 = 0;
  const int j1 = kw == 2 ? 2 : 1, px1 = kw == 1 ? 1 : 0, py1 = 1;
  const int64_t total = static_cast<int64_t>(9) * 64 * 64;
  const int64_t i0 = (static_cast<int64_t>(kh * 3 + j0) * 64 + px0 * 32 + ci) * 64 + py0 * 32 + co;
  const int64_t i1 = (static_cast<int64_t>(kh * 3 + j1) * 64 + px1 * 32 + ci) * 64 + py1 * 32 + co;
  const int s_per = (S + 7) / 8;
  const int s_lo = z * s_per, s_hi = min(S, s_lo + s_per);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
  for (int s = s_lo; s < s_hi; ++s) {
    a0 += __ldcs(partial + static_cast<int64_t>(s) * total + i0);
    a1 += __ldcs(partial + static_cast<int64_t>(s) * total + i1);
  }
  red[z][co] = a0 + a1;
  __syncthreads();
  if (z == 0) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][co];
    dw[(static_cast<int64_t>(co) * 32 + ci) * 9 + tap] = v;
  }
}

// MODE 2 partials: [split][kh][(px, ci)][window pixel b = 0..3][co]; pixel b of the window is 2q - 1 + b, X pixel is 2q + px,
// so the block holds kw = px + 2 - b:  dW[kh][kw] = D[px = 0][b = 2 - kw] + D[px = 1][b = 3 - kw].
__global__ void __launch_bounds__(256) wgrad_finalize_pairs_window_kernel(const float* __restrict__ partial,
                                                                           float* __restrict__ dw, int S) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][32];
  const int co = threadIdx.x & 31, z = threadIdx.x >> 5;
  const int ci = blockIdx.x & 31, tap = blockIdx.x >> 5;
  const int kh = tap / 3, kw = tap - kh * 3;
  const int64_t total = static_cast<int64_t>(3) * 64 * 128;
  const int64_t i0 = (static_cast<int64_t>(kh) * 64 + ci) * 128 + (2 - kw) * 32 + co;
  const int64_t i1 = (static_cast<int64_t>(kh) * 64 + 32 + ci) * 128 + (3 - kw) * 32 + co;
  const int s_per = (S + 7) / 8;
  const int s_lo = z * s_per, s_hi = min(S, s_lo + s_per);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
  for (int s = s_lo; s < s_hi; ++s) {
    a0 += __ldcs(partial + static_cast<int64_t>(s) * total + i0);
    a1 += __ldcs(partial + static_cast<int64_t>(s) * total + i1);
  }
  red[z][co] = a0 + a1;
  __syncthreads();
  if (z == 0) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][co];
    dw[(static_cast<int64_t>(co) * 32 + ci) * 9 + tap] = v;
  }
}

// B200UNET_WGRAD_PAIRS: 0 = wgradn<32,32>; 1 = pair views with three N units (MODE 0 / 1); default 2 = MODE 2
static int wgradn_pairs_mode() {
  const char* e = getenv("B200UNET_WGRAD_PAIRS");
  return e ? (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2)) : 2;
}

struct WgradNPlan {
  int CH, CO, CPB, gy, S, stages, blocks_w, blocks_h, total_kb, kb_per_split;
  int64_t smem_bytes, partial_floats;
};

static bool wgradn_one_dy(int Cin, int Cout) {
  const char* e = getenv("B200UNET_WGRAD_ONEDY");
  return Cout == 64 && Cin % 64 == 0 && !(e && e[0] == '0');
}

static void plan_wgradn(int N, int H, int W, int Cin, int Cout, WgradNPlan* pl, int force_one = -1) {
  pl->CO = Cout;
  pl->CH = (Cin % 64 == 0) ? 64 : 32;
  const int slots = 128 / pl->CH;
  const int groups = (3 + slots - 1) / slots;
  const int chunks = Cin / pl->CH;
  const int ncols = 3 * Cout;
  int cpb = 512 / (groups * ncols);  // accumulators that fit in TMEM
  if (cpb < 1) cpb = 1;
  if (cpb > chunks) cpb = chunks;
  // shared memory: at least 2 stages must fit
  const bool one = force_one >= 0 ? force_one != 0 : wgradn_one_dy(Cin, Cout);
  const int R = one ? kWnR1 : kWnR, BW = one ? kWnW1 : 16;
  const int xbytes = (R + 3) * BW * pl->CH * 2;
  const int dybytes = one ? ((R * (BW + 2) * Cout * 2 + 1023) / 1024) * 1024 : 3 * kWnR * 16 * Cout * 2;
  while (cpb > 1 && 2 * (cpb * xbytes + dybytes + 1023) > 216 * 1024) --cpb;
  while (chunks % cpb != 0) --cpb;  // every CTA gets the same number of chunks
  pl->CPB = cpb;
  pl->gy = chunks / cpb;
  const int stage = one ? ((cpb * xbytes + dybytes + 1023) / 1024) * 1024 : cpb * xbytes + dybytes;
  int stages = (216 * 1024) / stage;
  if (stages > kWnMaxStages) stages = kWnMaxStages;
  if (stages < 2) stages = 2;
  pl->stages = stages;
  pl->smem_bytes = static_cast<int64_t>(stages) * stage + 1024;
  pl->blocks_w = ceil_div(W, BW);
  pl->blocks_h = ceil_div(H, R);
  pl->total_kb = N * pl->blocks_w * pl->blocks_h;
  int S = num_sms() / pl->gy;
  if (S < 1) S = 1;
  if (S > pl->total_kb) S = pl->total_kb;
  pl->kb_per_split = ceil_div(pl->total_kb, S);
  pl->S = ceil_div(pl->total_kb, pl->kb_per_split);
  pl->partial_floats = static_cast<int64_t>(pl->S) * 9 * Cin * Cout;
}

bool wgradn_supported(int Cin, int Cout, int stride) {
  return stride == 1 && (Cout == 32 || Cout == 64) && Cin % 32 == 0 && Cin > 0;
}

int64_t wgradn_workspace_bytes(int N, int H, int W, int Cin, int Cout) {
  WgradNPlan pl;
  plan_wgradn(N, H, W, Cin, Cout, &pl);
  int64_t bytes = pl.partial_floats * 4;
  if (Cin == 32 && Cout == 32 && W % 2 == 0) {  // the pixel-pair form (chosen at launch when both tensors are dense)
    plan_wgradn(N, H, W / 2, 64, 64, &pl);
    if (pl.partial_floats * 4 > bytes) bytes = pl.partial_floats * 4;
  }
  return bytes;
}

template <int CH, int CO, int MODE = 0>
static int launch_wgradn(const WgradNMaps& maps, const WgradNParams& p, const WgradNPlan& pl, cudaStream_t st) {
  auto kern = wgradn_kernel<CH, CO, MODE>;
  static int attr_bytes = 0;
  if (attr_bytes < pl.smem_bytes) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    attr_bytes = (int)pl.smem_bytes;
  }
  launch_k(kern, dim3(pl.S, pl.gy), dim3(kConvThreads), pl.smem_bytes, st, maps, p);
  B200_LAUNCH_CHECK("wgradn_kernel");
  return 0;
}

// 32 -> 32 with both tensors dense: wgradn_kernel<64,64> on the (W/2, 64-channel) views + the pair finalize
static int wgradn_launch_pairs(const b200unet_conv_wgrad_args* a, cudaStream_t st) {
  WgradNPlan pl;
  const int Wp = a->W / 2;
  const bool window = wgradn_pairs_mode() == 2;
  const bool one = window || wgradn_one_dy(64, 64);
  plan_wgradn(a->N, a->H, Wp, 64, 64, &pl, one);
  B200_CHECK_ARG(a->workspace_bytes >= pl.partial_floats * 4, "conv_wgrad: workspace too small (%lld < %lld)",
                 (long long)a->workspace_bytes, (long long)(pl.partial_floats * 4));
  WgradNParams p{};
  WgradNMaps maps;
  p.N = a->N;
  p.OH = a->H;
  p.OW = Wp;
  p.blocks_w = pl.blocks_w;
  p.blocks_h = pl.blocks_h;
  p.total_kb = pl.total_kb;
  p.kb_per_split = pl.kb_per_split;
  p.cin = 64;
  p.cout = 64;
  p.chunks_per_cta = pl.CPB;
  p.stages = pl.stages;
  { const char* e = getenv("B200UNET_WGRADN_COLMAJOR"); p.col_major = (e && e[0] == '0') ? 0 : 1; }
  p.partial = a->workspace;
  int rc;
  if ((rc = make_act_map(&maps.x, static_cast<const __nv_bfloat16*>(a->x), 64, a->N, a->H, Wp, 64, 1, 1, 0, 0, 64,
                         one ? kWnW1 : 16, (one ? kWnR1 : kWnR) + 3)))
    return rc;
  if ((rc = make_act_map(&maps.dy, static_cast<const __nv_bfloat16*>(a->dy), 64, a->N, a->H, Wp, 64, 1, 1, 0, 0, 64,
                         one ? kWnW1 + 2 : 16, one ? kWnR1 : kWnR)))
    return rc;
  if ((rc = window ? launch_wgradn<64, 64, 2>(maps, p, pl, st)
                   : one ? launch_wgradn<64, 64, 1>(maps, p, pl, st) : launch_wgradn<64, 64>(maps, p, pl, st)))
    return rc;
  if (window)
    launch_k(wgrad_finalize_pairs_window_kernel, dim3(9 * 32), dim3(256), 0, st, static_cast<const float*>(a->workspace),
             a->dw, pl.S);
  else
    launch_k(wgrad_finalize_pairs_kernel, dim3(9 * 32), dim3(256), 0, st, static_cast<const float*>(a->workspace), a->dw,
             pl.S);
  B200_LAUNCH_CHECK("wgrad_finalize_pairs_kernel");
  return 0;
}

int wgradn_launch(const b200unet_conv_wgrad_args* a, cudaStream_t st) {
  if (a->Cin == 32 && a->Cout == 32 && a->x_pitch == 32 && a->dy_pitch == 32 && a->W % 2 == 0 && wgradn_pairs_mode() != 0)
    return wgradn_launch_pairs(a, st);
  WgradNPlan pl;
  plan_wgradn(a->N, a->H, a->W, a->Cin, a->Cout, &pl);
  B200_CHECK_ARG(a->workspace_bytes >= pl.partial_floats * 4, "conv_wgrad: workspace too small (%lld < %lld)",
                 (long long)a->workspace_bytes, (long long)(pl.partial_floats * 4));
  WgradNParams p{};
  WgradNMaps maps;
  p.N = a->N;
  p.OH = a->H;
  p.OW = a->W;
  p.blocks_w = pl.blocks_w;
  p.blocks_h = pl.blocks_h;
  p.total_kb = pl.total_kb;
  p.kb_per_split = pl.kb_per_split;
  p.cin = a->Cin;
  p.cout = a->Cout;
  p.chunks_per_cta = pl.CPB;
  p.stages = pl.stages;
  { const char* e = getenv("B200UNET_WGRADN_COLMAJOR"); p.col_major = (e && e[0] == '0') ? 0 : 1; }
  p.partial = a->workspace;
  int rc;
  const bool one = wgradn_one_dy(a->Cin, a->Cout);
  if ((rc = make_act_map(&maps.x, static_cast<const __nv_bfloat16*>(a->x), a->x_pitch, a->N, a->H, a->W, a->Cin, 1, 1, 0,
                         0, pl.CH, one ? kWnW1 : 16, (one ? kWnR1 : kWnR) + 3)))
    return rc;
  if ((rc = make_act_map(&maps.dy, static_cast<const __nv_bfloat16*>(a->dy), a->dy_pitch, a->N, a->H, a->W, a->Cout, 1,
                         1, 0, 0, pl.CO, one ? kWnW1 + 2 : 16, one ? kWnR1 : kWnR)))
    return rc;
  if (pl.CH == 32 && pl.CO == 32) rc = launch_wgradn<32, 32>(maps, p, pl, st);
  else if (pl.CH == 32 && pl.CO == 64) rc = launch_wgradn<32, 64>(maps, p, pl, st);
  else if (pl.CH == 64 && pl.CO == 32) rc = launch_wgradn<64, 32>(maps, p, pl, st);
  else if (one) rc = launch_wgradn<64, 64, 1>(maps, p, pl, st);
  else rc = launch_wgradn<64, 64>(maps, p, pl, st);
  if (rc) return rc;
  return launch_wgrad_finalize(a->workspace, a->dw, pl.S, a->Cin, a->Cout, st);
}

}  // namespace b200
