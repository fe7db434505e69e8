// The two steps either side of the hot path in the reference's trainer:
//  * the optimizer step -- torch.optim.SGD(momentum, nesterov=True, weight_decay) over the model's 90 parameter
//    tensors (Our_UNet/src/train.py:431-451) as ONE multi-tensor launch instead of ~5 foreach launches per step;
//  * the validation metric -- argmax over the 3 logits + per-class intersection / prediction / target counts over
//    the valid pixels (train.py:554-572, nine .item() host syncs per batch in the reference) as one pass with integer
//    counters (bit-exact) and no host sync.
#include "common.cuh"

namespace b200 {

constexpr int kSgdMaxTensors = 112;  // pointers travel in the kernel parameters (4 KB limit)
constexpr int kSgdThreads = 256, kSgdPerThread = 4;

struct SgdBatch {
  float* p[kSgdMaxTensors];
  const float* g[kSgdMaxTensors];
  float* buf[kSgdMaxTensors];
  int first_block[kSgdMaxTensors + 1];  // prefix sum of the blocks of each tensor
  int numel[kSgdMaxTensors];
  int count;
};

// Arithmetic mirrors torch's foreach SGD on CUDA operation by operation (each of its `a + alpha * b` kernels is one
// fused multiply-add; `buf.mul_(momentum)` and the following add are two roundings):
//   g1 = g + wd * p;  buf = first ? g1 : (buf * momentum) + g1;  g2 = g1 + momentum * buf;  p = p - lr * g2
__global__ void __launch_bounds__(kSgdThreads) sgd_nesterov_kernel(const __grid_constant__ SgdBatch B, float lr,
                                                                    float momentum, float wd, int nesterov,
                                                                    int first_step) {
  // which tensor owns this block: binary search in the prefix table
  int lo = 0, hi = B.count;
  const int b = blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (B.first_block[mid] <= b) lo = mid; else hi = mid;
  }
  const int t = lo;
  const int n = B.numel[t];
  float* __restrict__ p = B.p[t];
  const float* __restrict__ g = B.g[t];
  float* __restrict__ buf = B.buf[t];
  const int base = (b - B.first_block[t]) * kSgdThreads * kSgdPerThread;
#pragma unroll
  for (int k = 0; k < kSgdPerThread; ++k) {
    const int i = base + k * kSgdThreads + threadIdx.x;
    if (i >= n) break;
    const float pv = p[i];
    float gv = g[i];
    if (wd != 0.f) gv = fmaf(wd, pv, gv);
    float upd = gv;
    if (momentum != 0.f) {
      const float bv = first_step ? gv : __fadd_rn(__fmul_rn(buf[i], momentum), gv);
      buf[i] = bv;
      upd = nesterov ? fmaf(momentum, bv, gv) : bv;
    }
    p[i] = fmaf(-lr, upd, pv);
  }
}

// ------------------------------------------------------------------------------------------------ validation
constexpr int kEvalThreads = 256, kEvalPerThread = 8;

// counts[c][0] = #(pred == c & target == c & valid), [c][1] = #(pred == c & valid), [c][2] = #(target == c & valid)
__global__ void __launch_bounds__(kEvalThreads) argmax_counts_kernel(const float* __restrict__ logits,
                                                                      const int64_t* __restrict__ target,
                                                                      int ignore_index, int64_t* __restrict__ pred,
                                                                      unsigned long long* __restrict__ counts,
                                                                      int64_t HW) {
  __shared__ unsigned int sc[9];
  if (threadIdx.x < 9) sc[threadIdx.x] = 0;
  __syncthreads();
  const int n = blockIdx.y;
  const float* z = logits + static_cast<int64_t>(n) * 3 * HW;
  unsigned int c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kEvalThreads * kEvalPerThread;
#pragma unroll
  for (int k = 0; k < kEvalPerThread; ++k) {
    const int64_t px = base + static_cast<int64_t>(k) * kEvalThreads + threadIdx.x;
    if (px >= HW) break;
    const float a0 = z[px], a1 = z[HW + px], a2 = z[2 * HW + px];
    // torch.argmax: first maximal index (ties -> lowest index); NaN is maximal, like torch
    int am = 0;
    float best = a0;
    if (a1 > best || (a1 != a1 && best == best)) { am = 1; best = a1; }
    if (a2 > best || (a2 != a2 && best == best)) { am = 2; best = a2; }
    if (pred) pred[static_cast<int64_t>(n) * HW + px] = am;
    const int64_t t = target[static_cast<int64_t>(n) * HW + px];
    if (t == ignore_index) continue;
    c[am * 3 + 1] += 1;
    if (t >= 0 && t < 3) {
      c[static_cast<int>(t) * 3 + 2] += 1;
      if (t == am) c[am * 3 + 0] += 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    unsigned int v = c[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[i], v);
  }
  __syncthreads();
  if (threadIdx.x < 9 && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], static_cast<unsigned long long>(sc[threadIdx.x]));
}


// ------------------------------------------------------------------------------------------------ input pipeline
// PetSegmentationDataset.__getitem__ (Our_UNet/src/train.py:299-311) on the device: uint8 HWC image -> float / 255 ->
// (x - mean) / std -> fp32 NCHW, and uint8 mask -> clean-up (values > 2 other than 255 become 0) -> int64.  The batch
// crosses PCIe as uint8 (4 bytes per pixel instead of 20).  Same IEEE operations in the same order as the reference's
// torch CPU ops: bit-exact.
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                                             float* __restrict__ out, int64_t* __restrict__ mask_out,
                                                             float m0, float m1, float m2, float s0, float s1, float s2,
                                                             int64_t HW) {
  const int n = blockIdx.y;
  const int64_t px = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (px >= HW) return;
  if (img) {
    const uint8_t* p = img + (static_cast<int64_t>(n) * HW + px) * 3;
    float* o = out + static_cast<int64_t>(n) * 3 * HW + px;
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(p[0]), 255.f), m0), s0);
    o[HW] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(p[1]), 255.f), m1), s1);
    o[2 * HW] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(p[2]), 255.f), m2), s2);
  }
  if (mask) {
    const uint8_t v = mask[static_cast<int64_t>(n) * HW + px];
    mask_out[static_cast<int64_t>(n) * HW + px] = (v > 2 && v != 255) ? 0 : v;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200unet_sgd_max_tensors(void) { return kSgdMaxTensors; }

extern "C" int b200unet_sgd_nesterov_step(float* const* params, const float* const* grads, float* const* momentum_bufs,
                                          const int64_t* numels, int count, float lr, float momentum,
                                          float weight_decay, int nesterov, int first_step, void* stream) {
  B200_CHECK_ARG(params && grads && numels && count >= 0, "sgd_nesterov_step: null pointer");
  B200_CHECK_ARG(momentum == 0.f || momentum_bufs, "sgd_nesterov_step: momentum buffers required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int t0 = 0; t0 < count; t0 += kSgdMaxTensors) {
    SgdBatch B;
    B.count = count - t0 < kSgdMaxTensors ? count - t0 : kSgdMaxTensors;
    int blocks = 0;
    for (int i = 0; i < B.count; ++i) {
      const int64_t n = numels[t0 + i];
      B200_CHECK_ARG(n > 0 && n < (1ll << 31), "sgd_nesterov_step: tensor %d has %lld elements", t0 + i, (long long)n);
      B200_CHECK_ARG(params[t0 + i] && grads[t0 + i], "sgd_nesterov_step: tensor %d has a null pointer", t0 + i);
      B.p[i] = params[t0 + i];
      B.g[i] = grads[t0 + i];
      B.buf[i] = momentum_bufs ? momentum_bufs[t0 + i] : nullptr;
      B.numel[i] = static_cast<int>(n);
      B.first_block[i] = blocks;
      blocks += static_cast<int>(ceil_div64(n, kSgdThreads * kSgdPerThread));
    }
    B.first_block[B.count] = blocks;
    if (blocks == 0) continue;
    sgd_nesterov_kernel<<<blocks, kSgdThreads, 0, st>>>(B, lr, momentum, weight_decay, nesterov, first_step);
    B200_LAUNCH_CHECK("sgd_nesterov_kernel");
  }
  return 0;
}

extern "C" int b200unet_argmax_counts(const float* logits_nchw, const int64_t* target, int ignore_index,
                                      int64_t* pred_or_null, int64_t* counts9, int N, int64_t HW, void* stream) {
  B200_CHECK_ARG(logits_nchw && target && counts9, "argmax_counts: null pointer");
  B200_CHECK_ARG(N > 0 && N <= 65535 && HW > 0, "argmax_counts: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CUDA(cudaMemsetAsync(counts9, 0, 9 * sizeof(int64_t), st));
  const int blocks = static_cast<int>(ceil_div64(HW, kEvalThreads * kEvalPerThread));
  argmax_counts_kernel<<<dim3(blocks, N), kEvalThreads, 0, st>>>(logits_nchw, target, ignore_index, pred_or_null,
                                                                 reinterpret_cast<unsigned long long*>(counts9), HW);
  B200_LAUNCH_CHECK("argmax_counts_kernel");
  return 0;
}

extern "C" int b200unet_preprocess_u8(const void* image_u8_nhwc, const void* mask_u8, float* image_out_nchw,
                                      int64_t* mask_out, const float* mean3, const float* std3, int N, int64_t HW,
                                      void* stream) {
  B200_CHECK_ARG((image_u8_nhwc && image_out_nchw && mean3 && std3) || (mask_u8 && mask_out), "preprocess_u8: nothing to do");
  B200_CHECK_ARG(N > 0 && N <= 65535 && HW > 0, "preprocess_u8: bad sizes");
  const float m[3] = {mean3 ? mean3[0] : 0.f, mean3 ? mean3[1] : 0.f, mean3 ? mean3[2] : 0.f};
  const float sd[3] = {std3 ? std3[0] : 1.f, std3 ? std3[1] : 1.f, std3 ? std3[2] : 1.f};
  preprocess_u8_kernel<<<dim3((unsigned)ceil_div64(HW, 256), N), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(image_u8_nhwc), static_cast<const uint8_t*>(mask_u8), image_out_nchw, mask_out, m[0], m[1],
      m[2], sd[0], sd[1], sd[2], HW);
  B200_LAUNCH_CHECK("preprocess_u8_kernel");
  return 0;
}
