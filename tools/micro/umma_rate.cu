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

template <int CG, int MN>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long* cycles, int commit_every, int M1) {
  __shared__ __align__(8) uint64_t scratch_bar[4];
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = (CG == 2) ? N / 2 : N;                  // B rows held by this CTA
  const int stage_bytes = 128 * BK * 2 + ((nb * BK * 2 + 1023) / 1024) * 1024;
  // fill smem with something finite
  for (int i = threadIdx.x; i < STAGES * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&scratch_bar[i], 1); fence_barrier_init(); }
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
    const uint32_t idesc = umma_idesc_bf16(CG == 2 ? 256 : M1, N, MN, MN);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      const uint32_t sb = sa + 128 * BK * 2;
      // two accumulators alternate so consecutive MMAs are independent tiles like a real double-buffered kernel
      const uint32_t d = tmem_base + ((it & 1) ? 256 : 0);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) {
        // K-major: 128-byte rows of 64 K-elements, advance 32 bytes per K=16.  MN-major: rows of 64 MN-elements, one row
        // per K index: advance 16 rows per K=16, LBO = byte stride between 64-element MN blocks (64 K-rows each).
        const uint64_t ad = MN ? umma_smem_desc(sa + k * 16 * 128, 64 * 128, 1024, kSwz128) : umma_smem_desc(sa + k * 32, 16, 1024, kSwz128);
        const uint64_t bd = MN ? umma_smem_desc(sb + k * 16 * 128, 64 * 128, 1024, kSwz128) : umma_smem_desc(sb + k * 32, 16, 1024, kSwz128);
        if (CG == 1) umma_bf16(d, ad, bd, idesc, (it >= 2 || k) ? 1u : 0u);
        else umma_bf16_cg2(d, ad, bd, idesc, (it >= 2 || k) ? 1u : 0u);
      }
      // optional: a tcgen05.commit every `commit_every` iterations (4 MMAs each) onto a barrier nobody waits on
      if (CG == 1 && commit_every > 0 && (it % commit_every) == 0) umma_commit(&scratch_bar[it & 3]);
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

static int g_commit_every = 0;
template <int CG, int MN>
void run(int N, int iters, int M1 = 128) {
  long long* d;
  int ctas = 148;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaMemset(d, 0, ctas * sizeof(long long));
  const int nb = (CG == 2) ? N / 2 : N;
  const int stage_bytes = 128 * BK * 2 + ((nb * BK * 2 + 1023) / 1024) * 1024;
  int smem = STAGES * stage_bytes + 1024;
  cudaFuncSetAttribute(rate_kernel<CG, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, rate_kernel<CG, MN>, N, iters, d, g_commit_every, M1);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("CG%d N=%d failed: %s\n", CG, N, cudaGetErrorString(cudaGetLastError())); exit(1); }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < ctas; ++i) if (h[i] > mx) mx = h[i];
  const double mmas = (double)iters * (BK / 16);
  const int Mtot = CG == 2 ? 256 : M1;
  const double flops = 2.0 * Mtot * N * 16 * mmas * (ctas / CG);
  printf("commit/%d %s cta_group::%d M=%3d N=%3d  cycles/MMA %7.2f  (floor %5.1f)  chip %7.1f TFLOP/s  (%.3f ms)\n", g_commit_every, MN ? "MN-major" : "K-major ", CG, Mtot, N, mx / mmas,
         128.0 * N / 256.0 , flops / (ms * 1e-3) / 1e12, ms);
  cudaFree(d);
}


// Lean issue loop: descriptors are precomputed per stage, the K advance is a 64-bit add of 2 (32 bytes >> 4), the
// instruction descriptor and the accumulate predicate are constants.  Measures how fast ONE thread can issue.
template <int UNROLL, int COMMIT>
__global__ void __launch_bounds__(128, 1) lean_kernel(int N, int iters, long long* cycles) {
  __shared__ __align__(8) uint64_t sbar[4];
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = 128 * BK * 2 + ((N * BK * 2 + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < STAGES * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&sbar[i], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_holder, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t sb0 = smem_u32(&sbar[0]);
    uint64_t ad[STAGES], bd[STAGES];
    for (int s = 0; s < STAGES; ++s) {
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      ad[s] = umma_smem_desc(sa, 16, 1024, kSwz128);
      bd[s] = umma_smem_desc(sa + 128 * BK * 2, 16, 1024, kSwz128);
    }
    // first touch
    umma_bf16(tmem_base, ad[0], bd[0], idesc, 0u);
    umma_bf16(tmem_base + 256, ad[0], bd[0], idesc, 0u);
    long long t0 = clock64();
    for (int it = 0; it < iters; it += UNROLL) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int s = u % STAGES;
        const uint32_t d = tmem_base + ((u & 1) ? 256 : 0);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(ad[s] + 2 * k), "l"(bd[s] + 2 * k), "r"(idesc) : "memory");
        }
        // COMMIT = n: one tcgen05.commit after every n-th group of 4 MMAs, onto a barrier nobody waits on
        if (COMMIT > 0 && (u % COMMIT) == COMMIT - 1) umma_commit_u32(sb0 + (u & 3) * 8);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int COMMIT>
void run_lean(int N, int iters) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int stage_bytes = 128 * BK * 2 + ((N * BK * 2 + 1023) / 1024) * 1024;
  int smem = STAGES * stage_bytes + 1024;
  cudaFuncSetAttribute(lean_kernel<8, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    lean_kernel<8, COMMIT><<<148, 128, smem>>>(N, iters, d);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("lean N=%d failed: %s\n", N, cudaGetErrorString(cudaGetLastError())); exit(1); }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
  const double mmas = (double)iters * (BK / 16);
  printf("lean issue commit-every-%d-groups M=128 N=%3d  cycles/MMA %7.2f  (N/2 = %5.1f)  chip %7.1f TFLOP/s\n", COMMIT, N, mx / mmas, N / 2.0,
         2.0 * 128 * N * 16 * mmas * 148 / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

int main() {
  const int iters = 20000;
  for (int N : {32, 64, 128}) { run_lean<0>(N, iters); run_lean<8>(N, iters); run_lean<2>(N, iters); run_lean<1>(N, iters); }
  // round 2: K-major against MN-major operands (the weight-gradient kernels read both operands MN-major), N up to 256, M = 64
  for (int N : {64, 128, 192, 256}) { run<1, 0>(N, iters); run<1, 1>(N, iters); }
  for (int N : {128, 192, 256}) { run<1, 0>(N, iters, 64); run<1, 1>(N, iters, 64); }
  return 0;
}
