// Implicit-GEMM 3x3 convolutions on tcgen05 / TMEM, operands staged by TMA (sm_100a only): forward and data gradient.
//
// Replaces nn.Conv2d forward and the data-gradient half of aten::convolution_backward for the ConvBlock convs of the
// reference (Our_UNet/models/unet.py:106-115).  Activations are NHWC bf16, accumulation fp32 in tensor memory.
//
//   D[128 pixels, BN] = sum over taps (kh,kw) and channel chunks c of
//        A_tap[128 pixels, BK channels]  x  W[BN, tap*C + c*BK .. +BK]^T
//   A 128-pixel tile is a 16-row x 8-pixel patch of one image of the output lattice.
//
// Operand staging ("patch loads").  The K-major A operand of an MMA is 16 core-matrix groups of 8 rows (128-byte rows,
// 64-byte for BK = 32) at a uniform stride (SBO).  With an 8-pixel wide tile every group is ONE image row of the tile,
// so a tap-shifted A tile is a WINDOW of a (16 + 2) x (8 + 2) input patch: SBO = the patch's row pitch (10 pixels),
// start = the tap's (row, column) offset inside the patch.  The shared-memory swizzle is a function of the absolute
// address, so a window may start at any pixel (measured: bit-identical results for starts that are not multiples of
// the 1024-byte swizzle period, conv_wgrad_narrow.cu).  ONE patch load per channel chunk therefore serves all 9 taps:
// 1.4x the tile's own bytes from L2 instead of 3.75x with one 10 x 16 patch per column shift (round 1) or 9x with one
// tile per tap, and 1 instead of 3 / 9 TMA instructions (a 4-D cp.async.bulk.tensor costs its issuing thread ~420
// cycles on B200, tools/micro/tma_rate.cu).  The zero padding is TMA out-of-bounds fill.
// Stride 2: fprop reads the four parity sub-lattices of the input (4 patch loads per chunk); dgrad writes the four
// parity sub-lattices of dx through four launches (gather form: no scatter, no atomics).
// Weights (B): K-major [BN x BK] tiles of the packed weight matrix; when the whole slab of a layer (taps x chunks)
// fits in ~72 KB it is loaded ONCE per CTA and stays resident (all Cout <= 96 layers), otherwise it streams through
// its own ring.
//
// Persistent CTAs (one per SM) with three pipelines: A ring and B ring (TMA <-> MMA) running across tiles, and two
// TMEM accumulator buffers (MMA <-> epilogue) so that the epilogue of tile i overlaps the mainloop of tile i+1.
// Warp roles (288 threads): warps 0..3 = TMA producers (one elected lane each, issuing concurrently; each owns a
// fixed subset of ring slots so that it observes every phase of its barriers in order), warp 4 = TMEM allocator +
// single-thread MMA issuer, warps 5..8 = epilogue: tcgen05.ld -> bf16 -> swizzled staging -> TMA store, plus the
// per-(n,c) sum / sum-of-squares partials of the stored values that InstanceNorm needs (deterministic, no atomics).
//
// CTA pairs (template argument CG = 2; the streamed-weight layers).  Two CTAs of a cluster on the two SMs of a TPC work
// on two M tiles of the same image and the same N tile.  Only the even CTA (the leader) issues MMAs, with
// tcgen05.mma.cta_group::2 and M = 256: the hardware reads each CTA's own A window and HALF of the weight tile (BN / 2
// rows) from each CTA's shared memory at the same offsets and writes each CTA's 128 accumulator rows into its own
// TMEM.  Protocol: TMA loads of both CTAs complete on the leader's full barriers (which expect the bytes of both);
// tcgen05.commit multicasts every completion (slot free, accumulator ready) to the barriers at the same offset in both
// CTAs; the peer's epilogue warps release the accumulator buffer with a remote arrive on the leader's barrier (count
// 8); a cluster barrier after the barrier initialisation and before the TMEM release keeps either CTA from running
// ahead of, or leaving before, its peer.  Everything else -- producers, rings, epilogue, statistics -- is per CTA and
// unchanged.
#include "common.cuh"
#include "ptx.cuh"
#include "conv_common.cuh"
#include <stdlib.h>

namespace b200 {

constexpr int kTH = 16, kTW = 8;  // output patch of a tile (rows x pixels)
constexpr int kMaxLoads = 4;      // one patch per source lattice (stride 1: one; stride-2 fprop: the four parity classes)
constexpr int kMaxTaps = 9;

struct ALoad {
  int dh, dw;              // patch origin relative to the tile origin (in the load's source lattice)
  int rows, cols;          // patch size in image rows / pixels: tile + the span of the taps it serves
  int ntaps;               // taps served by this patch
  uint32_t a_hi;           // high word of the A descriptors of this patch (SBO = cols pixels)
  uint32_t mt_step16;      // window offset of the next M tile of a CTA tile (kTH image rows), 16-byte units
  uint32_t winoff16[kMaxTaps];  // start of each tap's window inside the patch, 16-byte units
  int koff[kMaxTaps];      // K offset (elements) of each tap in the packed weight matrix
};

struct GConvParams {
  int N, OH, OW, tiles_w, tiles_h;
  int nloads, ntaps_total;
  int s1;  // 1 when ONE patch serves 9 taps (every stride-1 conv): the issue loop is fully unrolled
  ALoad loads[kMaxLoads];
  int cin;   // channels per tap on the K side (multiple of BK)
  int cout;  // total N of the GEMM
  int tiles_per_cta;
  int stat_slots;  // P: partial-sum slots per image in `stats` ([N][P][cout][2])
  int mt;          // M tiles (of 128 pixels = 8 image rows) per CTA tile: 1, or 2 when the kernel shares each weight tile
  int out_split;   // 0: one output map.  > 0: output channel block j*out_split.. goes to output map j (the parity
                   // sub-lattices of the stride-2 data gradient stacked on N), channel coordinate relative to the block
  float* stats;
  long long* debug;  // optional [gridDim.x][8] cycle counters (developer instrumentation; NULL in production)
};

struct GConvMaps {
  CUtensorMap src[kMaxLoads];  // one per patch load (the box height is part of the descriptor)
  CUtensorMap w;
  CUtensorMap out[8];  // one per output channel block when the output is split (parity sub-lattices / two tensors)
};

// MT > 1: the CTA tile is MT vertically adjacent 16 x 8 patches (one (16 MT + 2)-row TMA patch, MT accumulators);
// every streamed weight tile feeds all of them, which divides the dominant L2->SM stream of the N <= 128 layers by MT
// (their MMA thread waited on weight tiles 35 % of the time).
// CG = 2: a CTA PAIR (cluster of two, one TPC) runs every MMA with cta_group::2, M = 256: each CTA keeps its own
// 128-pixel tile(s), A patches, accumulators and epilogue, but only HALF of every streamed weight tile (BN / 2 rows of
// B) -- the pair's tensor cores exchange the halves.  That halves the L2 -> SM weight stream AND the shared-memory
// operand reads of B per SM, the two things that hold the streamed-weight layers at 65-84 % tensor pipe.
template <int BK, int BN, int A_SLOTS, int B_SLOTS, bool B_RES, int OC_ = 0, int MT = 1, int CG = 1>
struct GConvCfg {
  static_assert(CG == 1 || (CG == 2 && !B_RES && BN % 32 == 0), "CTA pairs stream their weights");
  static constexpr int kRowBytes = BK * 2;
  static constexpr int kTHc = kTH * MT;                                         // image rows of a CTA tile
  static constexpr int kASlotBytes = (((kTHc + 2) * (kTW + 2) * kRowBytes + 1023) / 1024) * 1024;  // largest patch
  static constexpr int kBRows = BN / CG;                                        // weight rows this CTA loads per tile
  static constexpr int kBTileBytes = ((kBRows * BK * 2 + 1023) / 1024) * 1024; // one [BN / CG x BK] weight tile
  static constexpr int kBTileTx = kBRows * BK * 2;
  static constexpr int kBResBytes = 72 * 1024;                                 // resident weight slab budget
  static constexpr int kBBytes = B_RES ? kBResBytes : B_SLOTS * kBTileBytes;
  static constexpr int kOC = OC_ ? OC_ : ((BN % 64 == 0) ? 64 : 32);           // channels per TMA-store box
  static constexpr int kStageBufBytes = 128 * kOC * 2;                         // one staging buffer (two are used)
  static constexpr int kSmemBytes = A_SLOTS * kASlotBytes + kBBytes + 2 * kStageBufBytes + 1024;
  static constexpr uint32_t kSwz = (BK == 64) ? kSwz128 : kSwz64;
  static constexpr uint32_t kSbo = 8 * kRowBytes;
  static constexpr uint32_t kTmemCols = (2 * MT * BN <= 32) ? 32 : (2 * MT * BN <= 64) ? 64 : (2 * MT * BN <= 128) ? 128
                                        : (2 * MT * BN <= 256) ? 256 : 512;  // two buffers of MT accumulators
  static_assert(2 * MT * BN <= 512, "accumulators must fit in TMEM");
  // producers: A loads go to warps [0, NA), B tiles (streaming) to warps [4 - NB, 4)
  static constexpr int NA = B_RES ? ((A_SLOTS % 4 == 0) ? 4 : (A_SLOTS % 2 == 0) ? 2 : 1)
                                  : ((A_SLOTS % 2 == 0) ? 2 : 1);
  static constexpr int NB = B_RES ? 0 : ((B_SLOTS % 2 == 0) ? 2 : 1);
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int BK, int BN, int A_SLOTS, int B_SLOTS, bool B_RES, int OC_ = 0, int MT = 1, int CG = 1>
__global__ void __launch_bounds__(kConvThreads, 1) gconv_kernel(const __grid_constant__ GConvMaps maps,
                                                                 const __grid_constant__ GConvParams p) {
  pdl_launch_dependents();  // PDL (common.cuh): the next kernel of the stream may become resident as CTAs retire
  using Cfg = GConvCfg<BK, BN, A_SLOTS, B_SLOTS, B_RES, OC_, MT, CG>;
  constexpr int kBBar = B_RES ? 1 : B_SLOTS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[A_SLOTS], a_empty[A_SLOTS];
  __shared__ __align__(8) uint64_t b_full[kBBar], b_empty[kBBar];
  __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + A_SLOTS * Cfg::kASlotBytes;
  uint8_t* staging = smem_b + Cfg::kBBytes;

  // CTA pairs: the work unit is a PAIR tile = two consecutive M tiles (rank 0 / rank 1) of the same image x one N
  // tile; the host dispatches pairs only when an image has an even number of M tiles
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cta = (CG == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int n_tiles = p.cout / BN;
  const int total_tiles = (tiles_per_img / CG) * p.N * n_tiles;
  auto m_tile_of = [&](int tile) { return (tile / n_tiles) * CG + static_cast<int>(rank); };
  const int chunks = p.cin / BK;
  const int loads_per_tile = chunks * p.nloads;
  const int btiles_per_tile = chunks * p.ntaps_total;
  // contiguous tile range of this CTA (tile = m_tile * n_tiles + n_tile): consecutive patches of one image, which
  // keeps the halo rows in L2 and lets the epilogue accumulate the InstanceNorm partial sums across tiles
  const int tile_lo = min(cta * p.tiles_per_cta, total_tiles);
  const int tile_hi = min(tile_lo + p.tiles_per_cta, total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_SLOTS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kBBar; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4 * CG);  // one arrival per epilogue warp (of both CTAs of a pair: the leader's barrier)
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (CG == 2) {
      tmem_alloc_pair(&tmem_base_holder, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&tmem_base_holder, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  pdl_wait();  // barrier init / TMEM allocation above ran under the previous kernel's tail; global memory from here on

  if (warp < kProducerWarps) {
    if (elect_one()) {
      if (B_RES && warp == 0) {
        // ---------------------------------------------------------- resident weights: the whole (taps x chunks) slab
        // of the layer, loaded once per CTA (resident mode is dispatched only when cout == BN, i.e. n0 == 0)
        tma_prefetch_desc(&maps.w);
        mbar_expect_tx(&b_full[0], static_cast<uint32_t>(btiles_per_tile) * Cfg::kBTileTx);
        int u = 0;
        for (int c = 0; c < chunks; ++c)
          for (int l = 0; l < p.nloads; ++l)
            for (int t = 0; t < p.loads[l].ntaps; ++t, ++u)
              tma_load_2d(smem_b + u * Cfg::kBTileBytes, &maps.w, &b_full[0], p.loads[l].koff[t] + c * BK, 0);
      }
      if (warp < Cfg::NA) {
        // ------------------------------------------------------------ A producer `warp` of NA: patch loads
        tma_prefetch_desc(&maps.src[0]);
        const uint32_t a_full_leader = (CG == 2) ? mapa_u32(smem_u32(&a_full[0]), 0) : 0u;
        long long dbg_wait = 0, dbg_issue = 0;
        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
          const int m_tile = m_tile_of(tile);
          const int n_img = m_tile / tiles_per_img;
          const int t_in = m_tile - n_img * tiles_per_img;
          // column-major tile order inside an image: the next tile is the one BELOW, so the 2 halo rows of its
          // patches are still in L2 (row-major order re-fetched them from HBM: +25 % DRAM reads at 256^2 / 512^2)
          const int tw = t_in / p.tiles_h;
          const int h0 = (t_in - tw * p.tiles_h) * Cfg::kTHc, w0 = tw * kTW;
          const long long g0 = static_cast<long long>(it) * loads_per_tile;  // global load index of this tile's first
          int i = (warp - static_cast<int>(g0 % Cfg::NA) + Cfg::NA) % Cfg::NA;
          for (; i < loads_per_tile; i += Cfg::NA) {
            const long long g = g0 + i;
            const int slot = static_cast<int>(g % A_SLOTS);
            const uint32_t ph = static_cast<uint32_t>((g / A_SLOTS) & 1);
            const int c = i / p.nloads;
            const int l = i - c * p.nloads;
            const ALoad& L = p.loads[l];
            const long long t0 = (kInstr && p.debug) ? clock64() : 0;
            mbar_wait(&a_empty[slot], ph ^ 1);
            const long long t1 = (kInstr && p.debug) ? clock64() : 0;
            if (CG == 2) {
              // both CTAs' patches complete on the LEADER's barrier, which expects the bytes of both
              if (rank == 0) mbar_expect_tx(&a_full[slot], 2 * L.rows * L.cols * Cfg::kRowBytes);
              tma_load_4d_pair(smem + slot * Cfg::kASlotBytes, &maps.src[l], a_full_leader + slot * 8, c * BK, w0 + L.dw, h0 + L.dh, n_img);
            } else {
              mbar_expect_tx(&a_full[slot], L.rows * L.cols * Cfg::kRowBytes);
              tma_load_4d(smem + slot * Cfg::kASlotBytes, &maps.src[l], &a_full[slot], c * BK, w0 + L.dw, h0 + L.dh, n_img);
            }
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
      } else if (!B_RES && warp >= kProducerWarps - Cfg::NB) {
        // ------------------------------------------------------------ B producer: one [BN x BK] tile per (tap, chunk)
        const int me = warp - (kProducerWarps - Cfg::NB);
        tma_prefetch_desc(&maps.w);
        const uint32_t b_full_leader = (CG == 2) ? mapa_u32(smem_u32(&b_full[0]), 0) : 0u;
        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
          const int n0 = (tile % n_tiles) * BN + static_cast<int>(rank) * Cfg::kBRows;  // this CTA's half of the rows
          const long long g0 = static_cast<long long>(it) * btiles_per_tile;
          int u = 0;
          for (int c = 0; c < chunks; ++c)
            for (int l = 0; l < p.nloads; ++l)
              for (int t = 0; t < p.loads[l].ntaps; ++t, ++u) {
                const long long g = g0 + u;
                if (static_cast<int>(g % Cfg::NB) != me) continue;
                const int slot = static_cast<int>(g % B_SLOTS);
                const uint32_t ph = static_cast<uint32_t>((g / B_SLOTS) & 1);
                mbar_wait(&b_empty[slot], ph ^ 1);
                if (CG == 2) {
                  if (rank == 0) mbar_expect_tx(&b_full[slot], 2 * Cfg::kBTileTx);
                  tma_load_2d_pair(smem_b + slot * Cfg::kBTileBytes, &maps.w, b_full_leader + slot * 8, p.loads[l].koff[t] + c * BK, n0);
                } else {
                  mbar_expect_tx(&b_full[slot], Cfg::kBTileTx);
                  tma_load_2d(smem_b + slot * Cfg::kBTileBytes, &maps.w, &b_full[slot], p.loads[l].koff[t] + c * BK, n0);
                }
              }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (rank == 0 && elect_one()) {  // pairs: the leader issues for both CTAs
      // ONE thread issues every MMA of the CTA: its scalar instruction stream is the critical path, so everything
      // that can be is hoisted -- descriptors are 32-bit adds on precomputed halves, barrier addresses are
      // precomputed, the stride-1 tap pattern (one patch, 9 windows) is fully unrolled.
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CG, BN, 0, 0);
      auto commit = [](uint32_t bar) {
        if (CG == 2) umma_commit_pair_u32(bar); else umma_commit_u32(bar);
      };
      constexpr uint32_t hi = umma_desc_hi(Cfg::kSbo, Cfg::kSwz);  // B: dense [BN x BK] tiles
      constexpr uint32_t kASlot16 = Cfg::kASlotBytes >> 4, kBTile16 = Cfg::kBTileBytes >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem), 16);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b), 16);
      const uint32_t a_full0 = smem_u32(&a_full[0]), a_empty0 = smem_u32(&a_empty[0]);
      const uint32_t b_full0 = smem_u32(&b_full[0]), b_empty0 = smem_u32(&b_empty[0]);
      if (B_RES) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      const bool s1 = p.s1 != 0;
      uint32_t aslot = 0, aph = 0, bslot = 0, bph = 0;  // ring positions and phase parities (run across tiles)
      long long dbg_te = 0, dbg_af = 0, dbg_total0 = kInstr ? clock64() : 0;
      int it = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        const int buf = it & 1;
        const long long t0 = (kInstr && p.debug) ? clock64() : 0;
        mbar_wait(&tmem_empty_bar[buf], static_cast<uint32_t>(((it >> 1) & 1) ^ 1));  // epilogue drained this buffer
        if (kInstr && p.debug) dbg_te += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (MT * BN);
        uint32_t acc = 0;
        uint32_t b_res = b_lo0;  // next resident weight tile
        for (int c = 0; c < chunks; ++c) {
          const int nl = s1 ? 1 : p.nloads;
          for (int l = 0; l < nl; ++l) {
            const long long t1 = (kInstr && p.debug) ? clock64() : 0;
            mbar_wait_u32(a_full0 + aslot * 8, aph);
            if (kInstr && p.debug) dbg_af += clock64() - t1;
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + aslot * kASlot16;
            const ALoad& L = p.loads[s1 ? 0 : l];
            const uint32_t a_hi = L.a_hi, mt_step = L.mt_step16;
            auto tap = [&](int t) {
              uint32_t b_lo;
              if (B_RES) {
                b_lo = b_res;
                b_res += kBTile16;
              } else {
                mbar_wait_u32(b_full0 + bslot * 8, bph);
                tc_fence_after();
                b_lo = b_lo0 + bslot * kBTile16;
              }
              const uint32_t at_lo = a_lo + L.winoff16[t];
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {  // the same weight tile for every M tile: window kTH image rows further down
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  if (CG == 2)
                    umma_bf16_lean_pair(d_tmem + mt * BN, at_lo + mt * mt_step + 2 * k, a_hi, b_lo + 2 * k, hi, idesc, acc | k);
                  else
                    umma_bf16_lean(d_tmem + mt * BN, at_lo + mt * mt_step + 2 * k, a_hi, b_lo + 2 * k, hi, idesc, acc | k);
              }
              acc = 1;
              if (!B_RES) {
                commit(b_empty0 + bslot * 8);
                if (++bslot == B_SLOTS) { bslot = 0; bph ^= 1; }
              }
            };
            if (s1) {
#pragma unroll
              for (int t = 0; t < kMaxTaps; ++t) tap(t);
            } else {
              const int nt = L.ntaps;
              for (int t = 0; t < nt; ++t) tap(t);
            }
            commit(a_empty0 + aslot * 8);
            if (++aslot == A_SLOTS) { aslot = 0; aph ^= 1; }
          }
        }
        commit(smem_u32(&tmem_full_bar[buf]));
      }
      if (kInstr && p.debug) {
        p.debug[blockIdx.x * 8 + 2] = dbg_te;
        p.debug[blockIdx.x * 8 + 3] = dbg_af;
        p.debug[blockIdx.x * 8 + 4] = clock64() - dbg_total0;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 5..8)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 32 * kEpiWarp0;  // 0..127
    const bool do_stats = p.stats != nullptr;
    constexpr int OC = Cfg::kOC;
    // InstanceNorm partial sums: after a staging chunk is complete, warp q re-reads rows [32q, 32q+32) of it from
    // shared memory with lane = channel pair (OC = 64) or (row parity, channel pair) (OC = 32) -- one conflict-free
    // row per step, no shuffles -- and accumulates sum / sum of squares of the STORED bf16 values in registers across
    // all tiles of the same image; the sums are flushed (one partial slot per (image, CTA, q)) when the image changes.
    constexpr int kAccChunks = (OC == 64) ? 8 : 4;  // cout <= 512 (OC = 64) / cout <= 96 (OC = 32: BN in {32, 96})
    float acc[kAccChunks][4];             // [chunk of the layer's channels][s1 c0, s2 c0, s1 c1, s2 c1]
#pragma unroll
    for (int i = 0; i < kAccChunks; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const int cout_chunks = p.cout / OC;
    int acc_img = -1;
    auto flush = [&](int img) {
      // slot of this (CTA, q) among the CTAs that cover image `img`
      const int first_tile = img * (tiles_per_img / CG) * n_tiles;
      const int b0 = first_tile / p.tiles_per_cta;
      const int slot = ((cta - b0) * CG + static_cast<int>(rank)) * 4 + q;
      float* dst = p.stats + (static_cast<size_t>(img) * p.stat_slots + slot) * p.cout * 2;
#pragma unroll
      for (int i = 0; i < kAccChunks; ++i) {
        if (i < cout_chunks) {
          if (OC == 64) {
            // lane = channel pair: channels 2*lane, 2*lane+1 of chunk i
            *reinterpret_cast<float4*>(dst + (i * 64 + 2 * lane) * 2) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          } else {
            // lanes l and l+16 hold the even / odd rows of the same channel pair: combine, lane < 16 writes
            const float a0 = acc[i][0] + __shfl_down_sync(0xffffffffu, acc[i][0], 16);
            const float a1 = acc[i][1] + __shfl_down_sync(0xffffffffu, acc[i][1], 16);
            const float a2 = acc[i][2] + __shfl_down_sync(0xffffffffu, acc[i][2], 16);
            const float a3 = acc[i][3] + __shfl_down_sync(0xffffffffu, acc[i][3], 16);
            if (lane < 16) *reinterpret_cast<float4*>(dst + (i * 32 + 2 * lane) * 2) = make_float4(a0, a1, a2, a3);
          }
        }
        acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      }
    };
    int it = 0;
    long long dbg_tf = 0, dbg_e0 = kInstr ? clock64() : 0;
    uint32_t sbuf = 0;  // staging buffer toggle (runs across tiles)
    for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
      const int m_tile = m_tile_of(tile);
      const int nt_idx = tile % n_tiles;
      const int n0 = nt_idx * BN;
      const int n_img = m_tile / tiles_per_img;
      const int t_in = m_tile - n_img * tiles_per_img;
      const int tw = t_in / p.tiles_h;
      const int h0_tile = (t_in - tw * p.tiles_h) * Cfg::kTHc, w0 = tw * kTW;
      if (do_stats && n_img != acc_img) {
        if (acc_img >= 0) flush(acc_img);
        acc_img = n_img;
      }
      // rows of this warp's quarter that lie inside the image (rows below / right of it hold TMA zero-fill garbage
      // of the accumulator and must not enter the statistics)
      const int buf = it & 1;
      const long long t0 = (kInstr && p.debug) ? clock64() : 0;
      mbar_wait(&tmem_full_bar[buf], static_cast<uint32_t>((it >> 1) & 1));
      if (kInstr && p.debug) dbg_tf += clock64() - t0;
      tc_fence_after();
#pragma unroll 1
      for (int mtj = 0; mtj < MT * (BN / OC); ++mtj, sbuf ^= 1) {
        const int mt = mtj / (BN / OC), jb = mtj - mt * (BN / OC);  // M tile of the CTA tile, channel block
        const int h0 = h0_tile + mt * kTH;
        // the store issued two groups ago read this staging buffer
        if (et == 0) tma_store_wait_read_1();
        named_bar_sync(1, 128);
        uint8_t* stg = staging + sbuf * Cfg::kStageBufBytes;
#pragma unroll 1
        for (int c0 = jb * OC; c0 < (jb + 1) * OC; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (MT * BN) + mt * BN + c0, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          // staging store, swizzled like the TMA store box (conflict-free for these writes and for the TMA read)
          if (OC == 64) {
            const uint32_t base = smem_u32(stg) + row * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int cj = ((c0 & 63) >> 3) + i;
              const uint32_t addr = base + (((cj ^ (row & 7)) & 7) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                           "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                           : "memory");
            }
          } else {
            const uint32_t base = smem_u32(stg) + row * 64;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t addr = base + (((i ^ ((row >> 1) & 3)) & 3) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                           "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                           : "memory");
            }
          }
        }
        if (mtj == MT * (BN / OC) - 1) {
          // this warp's TMEM reads of the tile are complete: hand the accumulator buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(&tmem_empty_bar[buf], 0); else mbar_arrive(&tmem_empty_bar[buf]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (et == 0) {
          const int ch = n0 + jb * OC;
          if (p.out_split > 0) {
            const int sel = ch / p.out_split;
            tma_store_4d(&maps.out[sel], stg, ch - sel * p.out_split, w0, h0, n_img);
          } else {
            tma_store_4d(&maps.out[0], stg, ch, w0, h0, n_img);
          }
          tma_store_commit();
        }
        if (do_stats) {
          // rows of this warp's quarter: 32 (OC = 64: one per step) or 2 x 16 (OC = 32: lanes >= 16 take odd rows)
          constexpr int kSteps = (OC == 64) ? 32 : 16;
          const uint32_t cp = (OC == 64) ? static_cast<uint32_t>(lane) : (static_cast<uint32_t>(lane) & 15);
          const uint32_t par = (OC == 64) ? 0u : (static_cast<uint32_t>(lane) >> 4);
          const uint32_t cw = cp >> 2, wi = (cp & 3) << 2;
          const uint8_t* sbase = stg;
          const bool full = (h0 + kTH <= p.OH) && (w0 + kTW <= p.OW);
          float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f, t1a = 0.f, t2a = 0.f, t1b = 0.f, t2b = 0.f;
#pragma unroll
          for (int i0 = 0; i0 < kSteps; i0 += 8) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // issue the loads of 8 rows first, then consume
              const int r = q * 32 + ((OC == 64) ? (i0 + i) : (2 * (i0 + i) + static_cast<int>(par)));
              const uint32_t off = (OC == 64) ? (r * 128 + (((cw ^ (r & 7)) & 7) << 4) + wi)
                                              : (r * 64 + (((cw ^ ((r >> 1) & 3)) & 3) << 4) + wi);
              w[i] = lds_u32(smem_u32(sbase) + off);
              if (!full && !((h0 + (r >> 3) < p.OH) && (w0 + (r & 7) < p.OW))) w[i] = 0;  // outside the image
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float x0 = __uint_as_float(w[i] << 16), x1 = __uint_as_float(w[i] & 0xffff0000u);
              const float y0 = __uint_as_float(w[i + 1] << 16), y1 = __uint_as_float(w[i + 1] & 0xffff0000u);
              s1a += x0;
              s2a = fmaf(x0, x0, s2a);
              s1b += x1;
              s2b = fmaf(x1, x1, s2b);
              t1a += y0;
              t2a = fmaf(y0, y0, t2a);
              t1b += y1;
              t2b = fmaf(y1, y1, t2b);
            }
          }
          // static indexing of the accumulator array (a dynamic index would push it to local memory)
          const int ci = nt_idx * (BN / OC) + jb;
#pragma unroll
          for (int i = 0; i < kAccChunks; ++i)
            if (i == ci) {
              acc[i][0] += s1a + t1a;
              acc[i][1] += s2a + t2a;
              acc[i][2] += s1b + t1b;
              acc[i][3] += s2b + t2b;
            }
        }
      }
    }
    if (do_stats && acc_img >= 0) flush(acc_img);
    if (et == 0) tma_store_wait_all();
    if (kInstr && p.debug && et == 0) {
      p.debug[blockIdx.x * 8 + 5] = dbg_tf;
      p.debug[blockIdx.x * 8 + 6] = clock64() - dbg_e0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while its peer may still arrive on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// -------------------------------------------------------------------------------------------------- host side
struct TapSpec {
  int map;       // source lattice (parity class) index 0..3
  int dh, dw;    // shift in that lattice
  int koff;      // K offset of the tap in the packed weights
};

struct SrcLattice {
  const __nv_bfloat16* base;
  int64_t pitch;
  int N, H, W, C, hstep, wstep, hoff, woff;
};

// M tiles per CTA tile (GConvCfg MT): 2 for the streamed-weight N = 128 configuration with 128-byte rows.
static bool gconv_weights_resident(int BK, int BN, int cout_total, int k_channels, int ntaps) {
  // one N tile and the whole (taps x chunks) slab inside the 72 KB budget
  const int btile = ((BN * BK * 2 + 1023) / 1024) * 1024;
  return (cout_total == BN) && (BN <= 96) && (static_cast<long long>(ntaps) * (k_channels / BK) * btile <= 72 * 1024);
}
static int pick_mt(int BK, int BN, int cout_total, int k_channels, int ntaps) {
  if (BK != 64 || gconv_weights_resident(BK, BN, cout_total, k_channels, ntaps)) return 1;
  return BN == 128 ? 2 : (BN == 64 ? 4 : 1);  // as many accumulators as TMEM holds twice (double buffering)
}

// Group the taps of each source lattice into ONE patch load and build its tensor map.  p->mt must be set.
static int build_loads(const TapSpec* taps, int ntaps, const SrcLattice* lat, int BK, GConvParams* p, GConvMaps* maps) {
  p->nloads = 0;
  p->ntaps_total = ntaps;
  if (ntaps > kMaxTaps) return set_error(kErrInvalid, "gconv: %d taps", ntaps);
  const int row_bytes = BK * 2;
  const uint32_t swz = (BK == 64) ? kSwz128 : kSwz64;
  bool used[kMaxTaps] = {false};
  for (int i = 0; i < ntaps; ++i) {
    if (used[i]) continue;
    if (p->nloads == kMaxLoads) return set_error(kErrInvalid, "gconv: more than %d source lattices", kMaxLoads);
    int idx[kMaxTaps], n = 0;
    for (int j = i; j < ntaps; ++j)
      if (!used[j] && taps[j].map == taps[i].map) idx[n++] = j;
    int dh_min = taps[idx[0]].dh, dh_max = dh_min, dw_min = taps[idx[0]].dw, dw_max = dw_min;
    for (int k = 1; k < n; ++k) {
      dh_min = taps[idx[k]].dh < dh_min ? taps[idx[k]].dh : dh_min;
      dh_max = taps[idx[k]].dh > dh_max ? taps[idx[k]].dh : dh_max;
      dw_min = taps[idx[k]].dw < dw_min ? taps[idx[k]].dw : dw_min;
      dw_max = taps[idx[k]].dw > dw_max ? taps[idx[k]].dw : dw_max;
    }
    if (dh_max - dh_min > 2 || dw_max - dw_min > 2)
      return set_error(kErrInvalid, "gconv: tap span %d x %d too large", dh_max - dh_min, dw_max - dw_min);
    ALoad& L = p->loads[p->nloads];
    L.dh = dh_min;
    L.dw = dw_min;
    L.rows = kTH * p->mt + (dh_max - dh_min);
    L.cols = kTW + (dw_max - dw_min);
    L.ntaps = n;
    L.a_hi = umma_desc_hi(static_cast<uint32_t>(L.cols * row_bytes), swz);
    L.mt_step16 = static_cast<uint32_t>(kTH * L.cols * row_bytes) >> 4;
    for (int k = 0; k < n; ++k) {
      used[idx[k]] = true;
      L.winoff16[k] = static_cast<uint32_t>(((taps[idx[k]].dh - dh_min) * L.cols + (taps[idx[k]].dw - dw_min)) * row_bytes) >> 4;
      L.koff[k] = taps[idx[k]].koff;
    }
    const SrcLattice& S = lat[taps[i].map];
    int rc = make_act_map(&maps->src[p->nloads], S.base, S.pitch, S.N, S.H, S.W, S.C, S.hstep, S.wstep, S.hoff, S.woff,
                          BK, L.cols, L.rows);
    if (rc) return rc;
    ++p->nloads;
  }
  p->s1 = (p->nloads == 1 && p->loads[0].ntaps == kMaxTaps);
  return 0;
}

struct GConvGrid {
  int grid, tiles_per_cta, stat_slots;
};
// Persistent grid: contiguous tile ranges of equal length; P = partial-sum slots per image = 4 (lane quarters) x the
// largest number of CTAs whose range can intersect one image.
// cg = 2 (CTA pairs): the unit of work is a pair tile (two M tiles x one N tile), the unit of the grid a cluster of two
// CTAs; `units` = clusters that can be resident (at most one per TPC), grid = 2 x the clusters used, tiles_per_cta =
// pair tiles per cluster, and each pair owns 8 partial-sum slots per image it touches.
static GConvGrid gconv_grid(int N, int OH, int OW, int cout, int BN, int mt = 1, int cg = 1, int units = 0) {
  GConvGrid g;
  const long long per_img = static_cast<long long>(ceil_div(OW, kTW)) * ceil_div(OH, kTH * mt) / cg * (cout / BN);
  const long long total = per_img * N;
  if (units <= 0) units = num_sms() / cg;
  g.tiles_per_cta = static_cast<int>(ceil_div64(total, units));
  if (g.tiles_per_cta < 1) g.tiles_per_cta = 1;
  g.grid = cg * static_cast<int>(ceil_div64(total, g.tiles_per_cta));
  g.stat_slots = 4 * cg * (static_cast<int>(ceil_div64(per_img, g.tiles_per_cta)) + 1);
  return g;
}

// CTA pairs (GConvCfg CG = 2) for the streamed-weight configurations.  B200UNET_CG2=0 turns them off (developer knob).
static bool cg2_enabled() {
  static int state = -1;
  if (state < 0) {
    const char* e = getenv("B200UNET_CG2");
    state = (e && e[0] == '0') ? 0 : 1;
  }
  return state != 0;
}
// Measured per layer at batch 32 (tools/conv_bench.py, B200UNET_CG2=0 against 1): -4..-14 % on every streamed-weight
// layer with at least two channel chunks and more pair tiles than clusters; +8 % on the one layer with a single chunk
// (192 <- 64 data gradient: 9 weight tiles per tile, the pair's hand-overs are not amortised) and +4 % at 16^2, where
// a cluster gets a single tile (nothing to pipeline; the cluster launch costs more than it saves) -- those stay single.
static bool gconv_pairs(const GConvParams& p, int BK, int BN, bool resident) {
  if (!cg2_enabled() || BK != 64 || resident || p.out_split != 0) return false;
  const long long per_img = static_cast<long long>(p.tiles_w) * p.tiles_h;
  if (per_img % 2 != 0) return false;  // a pair never straddles two images
  if (p.cin < 2 * BK) return false;
  if (per_img / 2 * p.N * (p.cout / BN) <= num_sms() / 2) return false;
  return (BN == 256 && p.mt == 1) || (BN == 192 && p.mt == 1) || (BN == 128 && p.mt == 2) || (BN == 64 && p.mt == 4);
}

int conv_stat_slots(int N, int OH, int OW, int Cout) {
  // the largest slot count over the kernels / N tiles that may run for this Cout (the choice depends on Cin)
  const int BN64 = pick_bn_gconv(Cout, 64), BN32 = pick_bn_gconv(Cout, 32);
  if (!BN64 || !BN32) return -1;
  int slots = gconv_grid(N, OH, OW, Cout, BN64).stat_slots;
  const int s32 = gconv_grid(N, OH, OW, Cout, BN32).stat_slots;
  if (s32 > slots) slots = s32;
  const int mt64 = BN64 == 64 ? 4 : (BN64 == 128 ? 2 : 1);
  const int s2 = gconv_grid(N, OH, OW, Cout, BN64, mt64 == 1 ? 2 : mt64).stat_slots;  // the multi-M-tile variants
  if (s2 > slots) slots = s2;
  const int s3 = gconv_grid(N, OH, OW, Cout, BN64, mt64, 2).stat_slots;  // ... as CTA pairs (an upper bound: fewer
  if (s3 > slots) slots = s3;                                            // resident clusters mean fewer slots)
  if (Cout == 32 || Cout == 64) {
    const int sn = nconv_stat_slots(N, OH, OW);
    if (sn > slots) slots = sn;
  }
  return slots;
}

static int device_pairs_fallback() { return num_sms() / 2; }

// Developer instrumentation: B200UNET_GCONV_DEBUG=1 makes every launch synchronise and print per-role wait cycles.
static long long* debug_buffer() {
  static long long* buf = nullptr;
  static int state = -1;
  if (state < 0) {
    const char* e = getenv("B200UNET_GCONV_DEBUG");
    state = (e && e[0] == '1') ? 1 : 0;
    if (state) cudaMalloc(&buf, 148 * 8 * sizeof(long long));
  }
  return state ? buf : nullptr;
}

// A cluster of two CTAs per TPC; `units` of them can be resident at once (queried once per kernel).
template <typename Kern>
static int pair_units(Kern kern, int smem_bytes) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 74);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = device_pairs_fallback();
  }
  return n;
}

template <int BK, int BN, int A_SLOTS, int B_SLOTS, int MT>
static int launch_gconv_pairs(const GConvMaps& maps, const GConvParams& p_in, cudaStream_t st) {
  GConvParams p = p_in;
  p.debug = nullptr;
  using Cfg = GConvCfg<BK, BN, A_SLOTS, B_SLOTS, false, 0, MT, 2>;
  auto kern = gconv_kernel<BK, BN, A_SLOTS, B_SLOTS, false, 0, MT, 2>;
  if (p.mt != MT) return set_error(kErrInvalid, "gconv: tile height multiplier %d does not match the kernel (%d)", p.mt, MT);
  static int units = 0;
  if (units == 0) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    units = pair_units(kern, Cfg::kSmemBytes);
  }
  int u = num_sms() / 2;  // honours the SMs reserved for NCCL
  if (units < u) u = units;
  GConvGrid gg = gconv_grid(p.N, p.OH, p.OW, p.cout, BN, MT, 2, u);
  p.tiles_per_cta = gg.tiles_per_cta;
  p.stat_slots = conv_stat_slots(p.N, p.OH, p.OW, p.cout);  // P of the caller's buffer (>= gg.stat_slots)
  if (gg.stat_slots > p.stat_slots) return set_error(kErrInvalid, "gconv: %d partial-sum slots needed, %d provided", gg.stat_slots, p.stat_slots);
  if (p.stats) B200_CUDA(cudaMemsetAsync(p.stats, 0, static_cast<size_t>(p.N) * p.stat_slots * p.cout * 2 * sizeof(float), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gg.grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, maps, p);
  B200_LAUNCH_CHECK("gconv_kernel (CTA pairs)");
  return 0;
}

template <int BK, int BN, int A_SLOTS, int B_SLOTS, bool B_RES, int OC_ = 0, int MT = 1>
static int launch_gconv(const GConvMaps& maps, const GConvParams& p_in, cudaStream_t st) {
  GConvParams p = p_in;
  p.debug = debug_buffer();
  using Cfg = GConvCfg<BK, BN, A_SLOTS, B_SLOTS, B_RES, OC_, MT>;
  auto kern = gconv_kernel<BK, BN, A_SLOTS, B_SLOTS, B_RES, OC_, MT>;
  if (p.mt != MT) return set_error(kErrInvalid, "gconv: tile height multiplier %d does not match the kernel (%d)", p.mt, MT);
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const long long total = static_cast<long long>(p.tiles_w) * p.tiles_h * p.N * (p.cout / BN);
  GConvGrid gg = gconv_grid(p.N, p.OH, p.OW, p.cout, BN, MT);
  const int grid = gg.grid;
  p.tiles_per_cta = gg.tiles_per_cta;
  p.stat_slots = conv_stat_slots(p.N, p.OH, p.OW, p.cout);  // P of the caller's buffer (>= gg.stat_slots)
  if (p.stats) B200_CUDA(cudaMemsetAsync(p.stats, 0, static_cast<size_t>(p.N) * p.stat_slots * p.cout * 2 * sizeof(float), st));
  if (kInstr && p.debug) cudaMemsetAsync(p.debug, 0, 148 * 8 * sizeof(long long), st);
  launch_k(kern, dim3(grid), dim3(kConvThreads), Cfg::kSmemBytes, st, maps, p);
  B200_LAUNCH_CHECK("gconv_kernel");
  if (kInstr && p.debug) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.debug, sizeof(h), cudaMemcpyDeviceToHost);
    const long long tiles = total / grid;
    fprintf(stderr,
            "gconv<%d,%d,%d,%d,%d> cin %d cout %d loads %d taps %d tiles/CTA %lld | per tile: producer0 wait %lld issue %lld | "
            "mma total %lld wait_tmem_empty %lld wait_a_full %lld | epi total %lld wait_tmem_full %lld\n",
            BK, BN, A_SLOTS, B_SLOTS, (int)B_RES, p.cin, p.cout, p.nloads, p.ntaps_total, tiles, h[0] / tiles,
            h[1] / tiles, h[4] / tiles, h[2] / tiles, h[3] / tiles, h[6] / tiles, h[5] / tiles);
  }
  return 0;
}

// The packed weight matrix of a launch ([rows][K] bf16, K-major): its TMA box depends on the kernel chosen here.
struct WSpec {
  const void* w;
  int rows, K;
};

static int dispatch_gconv(GConvMaps& maps, const GConvParams& p, const WSpec& ws, int BK, int BN, cudaStream_t st) {
  int rc;
  if (p.out_split == 32 && BN == 128 && (BK == 64 || BK == 32)) {
    if ((rc = make_weight_map(&maps.w, ws.w, ws.rows, ws.K, BK, BN))) return rc;
    if (BK == 64) return launch_gconv<64, 128, 4, 4, false, 32>(maps, p, st);
    return launch_gconv<32, 128, 8, 4, false, 32>(maps, p, st);
  }
  // resident weights: one N tile and the whole (taps x chunks) slab inside the 72 KB budget
  const bool res = gconv_weights_resident(BK, BN, p.cout, p.cin, p.ntaps_total);
  if (gconv_pairs(p, BK, BN, res)) {
    // each CTA holds HALF of every weight tile: the ring is twice as deep in the same shared memory
    if ((rc = make_weight_map(&maps.w, ws.w, ws.rows, ws.K, BK, BN / 2))) return rc;
    if (BN == 256) return launch_gconv_pairs<64, 256, 3, 6, 1>(maps, p, st);
    if (BN == 192) return launch_gconv_pairs<64, 192, 3, 8, 1>(maps, p, st);
    if (BN == 128) return launch_gconv_pairs<64, 128, 2, 8, 2>(maps, p, st);
    return launch_gconv_pairs<64, 64, 2, 6, 4>(maps, p, st);
  }
  if ((rc = make_weight_map(&maps.w, ws.w, ws.rows, ws.K, BK, BN))) return rc;
#define GC(bk, bn, as, bs, r) \
  if (BK == bk && BN == bn && res == r) return launch_gconv<bk, bn, as, bs, r>(maps, p, st);
  // A slots now hold a whole channel chunk (all taps) of a CTA tile: two or three are a deep enough ring
  GC(64, 256, 2, 4, false)
  GC(64, 192, 3, 4, false)
  if (BK == 64 && BN == 128 && !res && p.mt == 2) return launch_gconv<64, 128, 2, 6, false, 0, 2>(maps, p, st);
  if (BK == 64 && BN == 64 && !res && p.mt == 4) return launch_gconv<64, 64, 2, 3, false, 0, 4>(maps, p, st);
  GC(64, 128, 4, 4, false)
  GC(64, 64, 6, 4, false)
  GC(64, 64, 4, 1, true)
  GC(64, 32, 8, 4, false)
  GC(64, 32, 4, 1, true)
  GC(32, 256, 4, 4, false)
  GC(32, 128, 8, 4, false)
  GC(32, 96, 8, 4, false)
  GC(32, 96, 8, 1, true)
  GC(32, 64, 8, 4, false)
  GC(32, 64, 8, 1, true)
  GC(32, 32, 8, 4, false)
  GC(32, 32, 8, 1, true)
#undef GC
  return set_error(kErrUnsupported, "no gconv instantiation for BK=%d BN=%d", BK, BN);
}

}  // namespace b200

using namespace b200;

extern "C" int b200unet_conv_fprop_partials(int N, int OH, int OW, int Cout) { return conv_stat_slots(N, OH, OW, Cout); }

extern "C" int b200unet_conv_fprop(const b200unet_conv_fprop_args* a, void* stream) {
  B200_CHECK_ARG(a && a->x && a->w && a->y, "conv_fprop: null pointer");
  B200_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv_fprop: stride %d unsupported", a->stride);
  const int BK = pick_bk(a->Cin), BN = pick_bn_gconv(a->Cout, BK);
  if (!BK || !BN)
    return set_error(kErrUnsupported, "conv_fprop: Cin=%d Cout=%d outside the tensor-core envelope (multiples of 32)",
                     a->Cin, a->Cout);
  B200_CHECK_ARG(a->x_pitch % 8 == 0 && a->y_pitch % 8 == 0, "conv_fprop: pitches must be multiples of 8 elements");
  if (nconv_supported(a->Cin, a->Cout, a->stride, a->W))
    return nconv_launch(a->x, a->x_pitch, a->w, a->y, a->y_pitch, a->stats, a->N, a->H, a->W, a->Cin, a->Cout, 0,
                        conv_stat_slots(a->N, a->H, a->W, a->Cout), static_cast<cudaStream_t>(stream));
  const int s = a->stride;
  const int OH = (a->H - 1) / s + 1, OW = (a->W - 1) / s + 1;
  GConvParams p{};
  GConvMaps maps;
  p.N = a->N;
  p.OH = OH;
  p.OW = OW;
  p.mt = pick_mt(BK, BN, a->Cout, a->Cin, 9);
  p.tiles_w = ceil_div(OW, kTW);
  p.tiles_h = ceil_div(OH, kTH * p.mt);
  p.cin = a->Cin;
  p.cout = a->Cout;
  p.stats = a->stats;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a->x);
  TapSpec taps[9];
  SrcLattice lat[4];
  int rc;
  if (s == 1) {
    lat[0] = SrcLattice{x, a->x_pitch, a->N, a->H, a->W, a->Cin, 1, 1, 0, 0};
    // order: column shift outermost so that the three row taps of a patch are adjacent
    int n = 0;
    for (int kw = 0; kw < 3; ++kw)
      for (int kh = 0; kh < 3; ++kh) taps[n++] = TapSpec{0, kh - 1, kw - 1, (kh * 3 + kw) * a->Cin};
  } else {
    B200_CHECK_ARG(a->H >= 2 && a->W >= 2, "conv_fprop: stride-2 input must be at least 2x2");
    for (int hp = 0; hp < 2; ++hp)
      for (int wp = 0; wp < 2; ++wp)
        lat[hp * 2 + wp] = SrcLattice{x, a->x_pitch, a->N, a->H, a->W, a->Cin, 2, 2, hp, wp};
    // input row 2*oh + kh - 1:  kh=0 -> parity 1, index oh-1;  kh=1 -> parity 0, index oh;  kh=2 -> parity 1, index oh
    const int par[3] = {1, 0, 1}, sh[3] = {-1, 0, 0};
    int n = 0;
    for (int kw = 0; kw < 3; ++kw)
      for (int kh = 0; kh < 3; ++kh)
        taps[n++] = TapSpec{par[kh] * 2 + par[kw], sh[kh], sh[kw], (kh * 3 + kw) * a->Cin};
  }
  if ((rc = build_loads(taps, 9, lat, BK, &p, &maps))) return rc;
  const WSpec ws{a->w, a->Cout, 9 * a->Cin};
  const int OC = (BN % 64 == 0) ? 64 : 32;
  if ((rc = make_act_map(&maps.out[0], static_cast<const __nv_bfloat16*>(a->y), a->y_pitch, a->N, OH, OW, a->Cout, 1, 1, 0,
                         0, OC, kTW, kTH)))
    return rc;
  return dispatch_gconv(maps, p, ws, BK, BN, static_cast<cudaStream_t>(stream));
}

// Partial-sum slots per image of the producer-side norm-backward sums, 0 when the kernel that runs this data gradient
// does not compute them (today: the narrow-output stride-1 kernels of the 512^2 and 256^2 levels, where the reduction
// passes they replace are largest).
extern "C" int b200unet_conv_dgrad_bwd_slots(int N, int H, int W, int Cin, int Cout, int stride) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return nconv_supported(Cout, Cin, stride, W) ? nconv_stat_slots(N, H, W) : 0;
}

extern "C" int b200unet_conv_dgrad(const b200unet_conv_dgrad_args* a, void* stream) {
  B200_CHECK_ARG(a && a->dy && a->wt && a->dx, "conv_dgrad: null pointer");
  B200_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv_dgrad: stride %d unsupported", a->stride);
  // GEMM: M = input pixels, N = Cin, K = taps * Cout
  const int BK = pick_bk(a->Cout), BN = pick_bn_gconv(a->Cin, BK);
  if (!BK || !BN)
    return set_error(kErrUnsupported, "conv_dgrad: Cin=%d Cout=%d outside the tensor-core envelope", a->Cin, a->Cout);
  B200_CHECK_ARG(a->dx_pitch % 8 == 0 && a->dy_pitch % 8 == 0, "conv_dgrad: pitches must be multiples of 8 elements");
  if (nconv_supported(a->Cout, a->Cin, a->stride, a->W) && !a->dx2) {
    if (a->bs_part) {
      B200_CHECK_ARG(a->bs_y && a->bs_a && a->bs_b, "conv_dgrad: norm-backward sums need y, a and b");
      BwdSums bs{static_cast<const __nv_bfloat16*>(a->bs_y), a->bs_y_pitch, a->bs_a, a->bs_b, a->bs_slope};
      return nconv_launch(a->dy, a->dy_pitch, a->wt, a->dx, a->dx_pitch, a->bs_part, a->N, a->H, a->W, a->Cout, a->Cin, 1,
                          nconv_stat_slots(a->N, a->H, a->W), static_cast<cudaStream_t>(stream), &bs);
    }
    return nconv_launch(a->dy, a->dy_pitch, a->wt, a->dx, a->dx_pitch, nullptr, a->N, a->H, a->W, a->Cout, a->Cin, 1, 0,
                        static_cast<cudaStream_t>(stream));
  }
  B200_CHECK_ARG(!a->bs_part, "conv_dgrad: this shape does not produce norm-backward sums (b200unet_conv_dgrad_bwd_slots == 0)");
  const int s = a->stride;
  const int OH = (a->H - 1) / s + 1, OW = (a->W - 1) / s + 1;
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(a->dy);
  const __nv_bfloat16* dx = static_cast<const __nv_bfloat16*>(a->dx);
  const int OC = (BN % 64 == 0) ? 64 : 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  GConvMaps maps;
  const WSpec ws{a->wt, a->Cin, 9 * a->Cout};
  SrcLattice lat[1] = {SrcLattice{dy, a->dy_pitch, a->N, OH, OW, a->Cout, 1, 1, 0, 0}};
  if (s == 1) {
    GConvParams p{};
    p.N = a->N;
    p.OH = a->H;
    p.OW = a->W;
    p.mt = pick_mt(BK, BN, a->Cin, a->Cout, 9);
    p.tiles_w = ceil_div(a->W, kTW);
    p.tiles_h = ceil_div(a->H, kTH * p.mt);
    p.cin = a->Cout;
    p.cout = a->Cin;
    p.stats = nullptr;
    // dx[ih,iw] += dy[ih + 1 - kh, iw + 1 - kw] * W[kh,kw]
    TapSpec taps[9];
    int n = 0;
    for (int kw = 0; kw < 3; ++kw)
      for (int kh = 2; kh >= 0; --kh) taps[n++] = TapSpec{0, 1 - kh, 1 - kw, (kh * 3 + kw) * a->Cout};  // row offsets 0,1,2
    if ((rc = build_loads(taps, 9, lat, BK, &p, &maps))) return rc;
    if (a->dx2) {
      // Two output tensors: channels [0, dx_split) go to dx, the rest to dx2 -- the two halves of the gradient of a
      // decoder concat buffer ([upsampled | skip], unet.py:228) written as two DENSE tensors.  At the 512^2 level the
      // concat gradient has a 192-byte pixel pitch: its 128-byte and 64-byte halves straddle / half-fill 128-byte
      // lines, and each consumer (upsample backward; the two passes of the encoder's norm backward) paid up to 2x the
      // DRAM bytes it used (ncu, profiles/r2_step_launches.md: 2.15 GB read for a 1.6 GB pass).
      B200_CHECK_ARG(a->dx_split > 0 && a->dx_split < a->Cin && a->dx_split % OC == 0 && a->Cin / OC <= 8 && BN == a->Cin,
                     "conv_dgrad: dx_split=%d must be a multiple of %d inside (0, %d), one N tile", a->dx_split, OC, a->Cin);
      B200_CHECK_ARG(a->dx2_pitch % 8 == 0, "conv_dgrad: dx2 pitch must be a multiple of 8 elements");
      p.out_split = OC;
      const __nv_bfloat16* dx2 = static_cast<const __nv_bfloat16*>(a->dx2);
      for (int j = 0; j < a->Cin / OC; ++j) {
        const int c0 = j * OC;
        const bool first = c0 < a->dx_split;
        if ((rc = make_act_map(&maps.out[j], first ? dx + c0 : dx2 + (c0 - a->dx_split), first ? a->dx_pitch : a->dx2_pitch,
                               a->N, a->H, a->W, OC, 1, 1, 0, 0, OC, kTW, kTH)))
          return rc;
      }
      return dispatch_gconv(maps, p, ws, BK, BN, st);
    }
    if ((rc = make_act_map(&maps.out[0], dx, a->dx_pitch, a->N, a->H, a->W, a->Cin, 1, 1, 0, 0, OC, kTW, kTH))) return rc;
    return dispatch_gconv(maps, p, ws, BK, BN, st);
  }
  B200_CHECK_ARG(!a->dx2, "conv_dgrad: a split output needs stride 1");
  // stride 2: one launch per parity class (ph,pw) of the input pixel; ih = 2a + ph receives
  //   ph = 0: kh = 1 from oh = a;      ph = 1: kh = 0 from oh = a + 1 and kh = 2 from oh = a   (same along w)
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int Hs = (a->H - ph + 1) / 2, Ws = (a->W - pw + 1) / 2;
      if (Hs <= 0 || Ws <= 0) continue;
      GConvParams p{};
      p.N = a->N;
      p.OH = Hs;
      p.OW = Ws;
      p.mt = pick_mt(BK, BN, a->Cin, a->Cout, (ph ? 2 : 1) * (pw ? 2 : 1));
      p.tiles_w = ceil_div(Ws, kTW);
      p.tiles_h = ceil_div(Hs, kTH * p.mt);
      p.cin = a->Cout;
      p.cout = a->Cin;
      p.stats = nullptr;
      int khs[2], dhs[2], nkh, kws[2], dws[2], nkw;
      if (ph == 0) { nkh = 1; khs[0] = 1; dhs[0] = 0; } else { nkh = 2; khs[0] = 0; dhs[0] = 1; khs[1] = 2; dhs[1] = 0; }
      if (pw == 0) { nkw = 1; kws[0] = 1; dws[0] = 0; } else { nkw = 2; kws[0] = 0; dws[0] = 1; kws[1] = 2; dws[1] = 0; }
      TapSpec taps[4];
      int n = 0;
      for (int j = 0; j < nkw; ++j)
        for (int i = 0; i < nkh; ++i) taps[n++] = TapSpec{0, dhs[i], dws[j], (khs[i] * 3 + kws[j]) * a->Cout};
      if ((rc = build_loads(taps, n, lat, BK, &p, &maps))) return rc;
      if ((rc = make_act_map(&maps.out[0], dx, a->dx_pitch, a->N, a->H, a->W, a->Cin, 2, 2, ph, pw, OC, kTW, kTH)))
        return rc;
      if ((rc = dispatch_gconv(maps, p, ws, BK, BN, st))) return rc;
    }
  return 0;
}

// ------------------------------------------------------------------------------ stride-2 data gradient, parities on N
// The four parity classes (ph, pw) of the input pixel of a stride-2 conv use 1, 2, 2 and 4 of the 9 taps; run as four
// launches each is a narrow GEMM (N = Cin: 25 % of the tensor array for Cin = 32) that re-reads dy.  For Cin <= 64 the
// classes are STACKED ON N instead: one launch over the coarse grid (a, b) with
//     D[(a, b), (ph, pw, ci)] = sum over the 4 shifts (dh, dw) in {0,1}^2 and co of dy[a + dh, b + dw, co] * Ws[(ph, pw, ci)][(dh, dw)][co]
// N = 4 * Cin (128 / 256: full tensor rate), dy read once; Ws holds W[kh(ph, dh)][kw(pw, dw)] where the class uses that
// shift and zero elsewhere (7 of 16 blocks); the epilogue stores channel block (ph, pw) through that class's
// sub-lattice tensor map.   dx[2a + ph] = sum_dh dy[a + dh] * W[kh]:  ph = 0: (dh = 0, kh = 1);  ph = 1: (dh = 0, kh = 2), (dh = 1, kh = 0).
__global__ void pack_s2_dgrad_weights_kernel(const __nv_bfloat16* __restrict__ wt, __nv_bfloat16* __restrict__ ws, int Cin,
                                             int Cout) {
  // wt [Cin][3][3][Cout] -> ws [4*Cin][4][Cout]
  const int64_t total = static_cast<int64_t>(4) * Cin * 4 * Cout;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % Cout);
  int64_t r = i / Cout;
  const int shift = static_cast<int>(r % 4);
  r /= 4;
  const int ci = static_cast<int>(r % Cin);
  const int par = static_cast<int>(r / Cin);
  const int ph = par >> 1, pw = par & 1, dh = shift >> 1, dw = shift & 1;
  const int kh = ph == 0 ? (dh == 0 ? 1 : -1) : (dh == 0 ? 2 : 0);
  const int kw = pw == 0 ? (dw == 0 ? 1 : -1) : (dw == 0 ? 2 : 0);
  __nv_bfloat16 v = __float2bfloat16_rn(0.f);
  if (kh >= 0 && kw >= 0) v = wt[((static_cast<int64_t>(ci) * 3 + kh) * 3 + kw) * Cout + co];
  ws[i] = v;
}

extern "C" int b200unet_conv_dgrad_s2_supported(int Cin, int Cout) {
  return (Cin == 32 || Cin == 64) && Cout % 32 == 0 && Cout > 0;
}

extern "C" int b200unet_pack_s2_dgrad_weights(const void* wt, void* ws, int Cin, int Cout, void* stream) {
  B200_CHECK_ARG(wt && ws, "pack_s2_dgrad_weights: null pointer");
  const int64_t total = static_cast<int64_t>(16) * Cin * Cout;
  pack_s2_dgrad_weights_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(wt), static_cast<__nv_bfloat16*>(ws), Cin, Cout);
  B200_LAUNCH_CHECK("pack_s2_dgrad_weights_kernel");
  return 0;
}

extern "C" int b200unet_conv_dgrad_s2(const b200unet_conv_dgrad_args* a, void* stream) {
  B200_CHECK_ARG(a && a->dy && a->wt && a->dx, "conv_dgrad_s2: null pointer");
  B200_CHECK_ARG(a->stride == 2 && b200unet_conv_dgrad_s2_supported(a->Cin, a->Cout), "conv_dgrad_s2: needs stride 2, Cin in {32, 64}");
  B200_CHECK_ARG(a->dx_pitch % 8 == 0 && a->dy_pitch % 8 == 0, "conv_dgrad_s2: pitches must be multiples of 8 elements");
  B200_CHECK_ARG(a->H >= 2 && a->W >= 2, "conv_dgrad_s2: input must be at least 2x2");
  const int OH = (a->H - 1) / 2 + 1, OW = (a->W - 1) / 2 + 1;
  const int BK = pick_bk(a->Cout), BN = 4 * a->Cin;
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(a->dy);
  const __nv_bfloat16* dx = static_cast<const __nv_bfloat16*>(a->dx);
  GConvParams p{};
  GConvMaps maps;
  p.N = a->N;
  p.OH = OH;  // the coarse grid: one tile pixel = one 2x2 block of dx
  p.OW = OW;
  p.mt = 1;
  p.tiles_w = ceil_div(OW, kTW);
  p.tiles_h = ceil_div(OH, kTH);
  p.cin = a->Cout;
  p.cout = BN;
  p.stats = nullptr;
  p.out_split = a->Cin;
  int rc;
  SrcLattice lat[1] = {SrcLattice{dy, a->dy_pitch, a->N, OH, OW, a->Cout, 1, 1, 0, 0}};
  TapSpec taps[4];
  int n = 0;
  for (int dw = 0; dw < 2; ++dw)
    for (int dh = 0; dh < 2; ++dh) taps[n++] = TapSpec{0, dh, dw, (dh * 2 + dw) * a->Cout};
  if ((rc = build_loads(taps, 4, lat, BK, &p, &maps))) return rc;
  const WSpec ws{a->wt, BN, 4 * a->Cout};
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw)
      if ((rc = make_act_map(&maps.out[ph * 2 + pw], dx, a->dx_pitch, a->N, a->H, a->W, a->Cin, 2, 2, ph, pw, a->Cin, kTW,
                             kTH)))
        return rc;
  return dispatch_gconv(maps, p, ws, BK, BN, static_cast<cudaStream_t>(stream));
}
