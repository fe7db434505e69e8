// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32) with BOTH operands in shared memory,
// K-major 128B-swizzled tiles, as a function of N, for cta_group::1 (M=128) and cta_group::2 (M=256).
// Answers: how much does a narrow N (Cout = 32/64) cost on B200 when A is re-read from smem for every MMA?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../unet-implementations_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200;

__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int STAGES = 4;
constexpr int BK = 64;

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = (CG == 2) ? N / 2 : N;                  // B rows held by this CTA
  const int stage_bytes = 128 * BK * 2 + ((nb * BK * 2 + 1023) / 1024) * 1024;
  // fill smem with something finite
  for (int i = threadIdx.x; i < STAGES * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    if (CG == 1) { tmem_alloc(&tmem_holder, 512); tmem_relinquish(); }
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const bool leader = (CG == 1) || cluster_ctarank() == 0;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && lane == 0 && leader) {
    const uint32_t idesc = umma_idesc_bf16(CG == 2 ? 256 : 128, N, 0, 0);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      const uint32_t sb = sa + 128 * BK * 2;
      // two accumulators alternate so consecutive MMAs are independent tiles like a real double-buffered kernel
      const uint32_t d = tmem_base + ((it & 1) ? 256 : 0);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) {
        const uint64_t ad = umma_smem_desc(sa + k * 32, 16, 1024, kSwz128);
        const uint64_t bd = umma_smem_desc(sb + k * 32, 16, 1024, kSwz128);
        if (CG == 1) umma_bf16(d, ad, bd, idesc, (it >= 2 || k) ? 1u : 0u);
        else umma_bf16_cg2(d, ad, bd, idesc, (it >= 2 || k) ? 1u : 0u);
      }
    }
    if (CG == 1) umma_commit(&bar); else umma_commit_cg2(&bar, 3);
    mbar_wait(&bar, 0);
    t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  } else if (CG == 2 && warp == 1 && lane == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 1) tmem_dealloc(tmem_base, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int CG>
void run(int N, int iters) {
  long long* d;
  int ctas = 148;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaMemset(d, 0, ctas * sizeof(long long));
  const int nb = (CG == 2) ? N / 2 : N;
  const int stage_bytes = 128 * BK * 2 + ((nb * BK * 2 + 1023) / 1024) * 1024;
  int smem = STAGES * stage_bytes + 1024;
  cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, N, iters, d);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("CG%d N=%d failed: %s\n", CG, N, cudaGetErrorString(cudaGetLastError())); exit(1); }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < ctas; ++i) if (h[i] > mx) mx = h[i];
  const double mmas = (double)iters * (BK / 16);
  const int Mtot = CG == 2 ? 256 : 128;
  const double flops = 2.0 * Mtot * N * 16 * mmas * (ctas / CG);
  printf("cta_group::%d M=%3d N=%3d  cycles/MMA %7.2f  (floor %5.1f)  chip %7.1f TFLOP/s  (%.3f ms)\n", CG, Mtot, N, mx / mmas,
         128.0 * N / 256.0 , flops / (ms * 1e-3) / 1e12, ms);
  cudaFree(d);
}

int main() {
  const int iters = 20000;
  for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) run<1>(N, iters);
  for (int N : {32, 64, 96, 128, 192, 256}) run<2>(N, iters);
  return 0;
}
