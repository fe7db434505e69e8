// Host helpers shared by the tensor-core convolution translation units: TMA descriptors over (pitched) NHWC
// activations and packed weight matrices, tile-shape selection.
#pragma once
#include "common.cuh"

namespace b200 {

// Warp roles of the tensor-core conv kernels.  A cp.async.bulk.tensor issue blocks its thread for ~420 cycles (4-D
// box) / ~200 cycles (2-D box) on B200 regardless of the box size (tools/micro/tma_rate.cu), so ONE producer thread
// caps a CTA at one box per ~420 cycles (< 20 B/cycle/SM with 8 KB boxes): the loads are spread over four producer
// warps (one elected lane each), which issue concurrently.
// Per-role cycle counters in the conv kernels (B200UNET_GCONV_DEBUG=1 prints them) exist only in a build with
// -DB200UNET_INSTRUMENT (make INSTRUMENT=1): even predicated-off clock reads in the single MMA-issuing thread cost
// ~15 % on the short-tile kernels.
#ifdef B200UNET_INSTRUMENT
constexpr bool kInstr = true;
#else
constexpr bool kInstr = false;
#endif
constexpr int kProducerWarps = 4;                 // warps 0..3: TMA producers
constexpr int kMmaWarp = 4;                       // warp 4: TMEM owner + single-thread MMA issuer
constexpr int kEpiWarp0 = 5;                      // warps 5..8: epilogue (TMEM lane quarter = warp % 4)
constexpr int kConvThreads = 32 * (kEpiWarp0 + 4);

static inline int make_act_map(CUtensorMap* m, const __nv_bfloat16* base, int64_t pitch, int N, int H, int W, int C,
                        int hstep, int wstep, int hoff, int woff, int boxC, int boxW, int boxH) {
  // 4-D view (C, W', H', N) of the sub-lattice {(hoff + hstep*i, woff + wstep*j)} of an NHWC tensor
  const int Hs = (H - hoff + hstep - 1) / hstep;
  const int Ws = (W - woff + wstep - 1) / wstep;
  if (Hs <= 0 || Ws <= 0) return set_error(kErrInvalid, "empty sub-lattice");
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)(wstep * pitch * 2), (uint64_t)(hstep * (int64_t)W * pitch * 2),
                         (uint64_t)((int64_t)H * W * pitch * 2)};
  uint32_t box[4] = {(uint32_t)boxC, (uint32_t)boxW, (uint32_t)boxH, 1};
  const CUtensorMapSwizzle swz = (boxC * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (boxC * 2 == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                    : CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swz == CU_TENSOR_MAP_SWIZZLE_NONE) return set_error(kErrInvalid, "box channel count %d unsupported", boxC);
  return make_tmap_bf16(m, base + ((int64_t)hoff * W + woff) * pitch, 4, dims, strides, box, swz);
}

static inline int make_weight_map(CUtensorMap* m, const void* w, int rows, int K, int boxK, int boxRows) {
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  uint64_t strides[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {(uint32_t)boxK, (uint32_t)boxRows};
  return make_tmap_bf16(m, w, 2, dims, strides, box,
                        boxK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

static inline int pick_bn(int n_total) {
  if (n_total % 256 == 0) return 256;
  if (n_total % 128 == 0) return 128;
  if (n_total % 64 == 0) return 64;
  if (n_total % 32 == 0) return 32;
  return 0;
}
// N tile of the fprop/dgrad kernel: additionally 96 (one tile instead of three 32-wide ones, 75 % MMA rate)
// and 192 (N = 192 / 384: the data gradients of the decoder convs that read a concat buffer -- one or two full-rate
// tiles instead of three 64- or 128-wide passes over the same A operand)
static inline int pick_bn_gconv(int n_total, int bk) {
  if (n_total == 96 && bk == 32) return 96;
  if ((n_total == 192 || n_total == 384) && bk == 64) return 192;
  return pick_bn(n_total);
}
static inline int pick_bk(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 0); }


// fp32 split-K partials [S][9][cin][cout] -> OIHW gradient (conv_tc.cu)
int launch_wgrad_finalize(const float* partial, float* dw, int S, int cin, int cout, cudaStream_t st);

// Producer-side InstanceNorm-backward sums of a data-gradient kernel (b200unet_conv_dgrad_args.bs_*): the epilogue reads
// the consuming unit's raw output y next to the gradient tile it has just staged and accumulates, per image,
// T1 = sum gm, T2raw = sum gm * y with gm = dx_stored * lrelu'(a*y + b) into `part` ([N][P][C][2], the layout of the
// forward statistics).
struct BwdSums {
  const __nv_bfloat16* y;
  int64_t y_pitch;
  const float* a;
  const float* b;
  float slope;
};

// Narrow-output forward / data gradient with the column taps stacked on N (conv_narrow.cu): stride 1, N-side
// channels in {32, 64}, weights resident in shared memory, image at least 64 pixels wide.
bool nconv_supported(int k_channels, int n_channels, int stride, int W);
int nconv_stat_slots(int N, int H, int W);
int nconv_launch(const void* src, int64_t src_pitch, const void* wpack, void* out, int64_t out_pitch, float* stats, int N,
                 int H, int W, int k_channels, int n_channels, int rev, int stat_slots, cudaStream_t st,
                 const BwdSums* bs = nullptr);
// 32 -> 32 channels on dense tensors with pixel pairs as 128-byte operand rows (conv_pair.cu); taken by nconv_launch
// when it applies.
bool pconv_supported(int k_channels, int n_channels, int stride, int W, int64_t src_pitch, int64_t out_pitch);
int pconv_stat_slots(int N, int H, int W);
int pconv_launch(const void* src, const void* wpack, void* out, float* stats, int N, int H, int W, int rev, int stat_slots,
                 cudaStream_t st, const BwdSums* bs = nullptr);
// Partial-sum slots per image of the fprop statistics buffer [N][P][Cout][2]: one value for both fprop kernels
// (conv_fprop_dgrad.cu), so that the caller can size the buffer without knowing which kernel will run.
int conv_stat_slots(int N, int OH, int OW, int Cout);

// Narrow-output weight gradient (conv_wgrad_narrow.cu): stride 1, Cout in {32, 64}, Cin a multiple of 32.
bool wgradn_supported(int Cin, int Cout, int stride);
int64_t wgradn_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int wgradn_launch(const b200unet_conv_wgrad_args* a, cudaStream_t st);

}  // namespace b200
