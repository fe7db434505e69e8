// Micro-benchmark: does TMA multicast across a 2-CTA cluster raise the rate at which L2-resident WEIGHT tiles reach an
// SM's shared memory?  The streamed-weight conv layers (gconv<64,128/192/256>) ask 60-90 B/cycle/SM of their L2->SM
// stream and sit at 65-84 % tensor pipe.  Two modes, same bytes LANDING per SM and iteration:
//   unicast   : every CTA loads the whole [rows x 64] bf16 tile (128 B swizzle) itself;
//   multicast : the two CTAs of a cluster each load HALF of the tile with .multicast::cluster to both (every byte leaves
//               L2 once per pair); a slot is reused when BOTH CTAs have released it (remote mbarrier arrive).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../unet-implementations_b200/csrc mcast_rate.cu ../../unet-implementations_b200/csrc/api.cu -o mcast_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "common.cuh"
#include "ptx.cuh"
#include "conv_common.cuh"
using namespace b200;

constexpr int kStages = 4, kProd = 2;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// rows x 64 bf16 tiles of a [total_rows x 64] matrix; producer p of kProd handles iterations i = p, p + kProd, ...
template <bool MC>
__global__ void __launch_bounds__(256, 1) mcast_kernel(const __grid_constant__ CUtensorMap map_full,
                                                        const __grid_constant__ CUtensorMap map_half, int rows, int iters,
                                                        int total_tiles, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t rank = MC ? cluster_ctarank() : 0;
  const int tile_bytes = rows * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC ? 2 : 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (MC) cluster_sync_all();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = MC ? blockIdx.x >> 1 : blockIdx.x;
  if (warp < kProd && lane == 0) {
    long long t0 = clock64();
    for (int i = warp; i < iters; i += kProd) {
      const int s = i % kStages;
      mbar_wait(&empty_bar[s], ((i / kStages) & 1) ^ 1);
      mbar_expect_tx(&full_bar[s], tile_bytes);
      const int t = (pair * 37 + i) % total_tiles;
      if (MC)
        tma_load_2d_mcast(smem + s * tile_bytes + rank * (tile_bytes / 2), &map_half, &full_bar[s], 0,
                          t * rows + rank * (rows / 2), 3);
      else
        tma_load_2d(smem + s * tile_bytes, &map_full, &full_bar[s], 0, t * rows);
    }
    if (warp == 0) cycles[blockIdx.x] = clock64() - t0;
  } else if (warp == 4 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % kStages;
      mbar_wait(&full_bar[s], (i / kStages) & 1);
      if (MC) {
        mbar_arrive_remote(&empty_bar[s], 0);
        mbar_arrive_remote(&empty_bar[s], 1);
      } else {
        mbar_arrive(&empty_bar[s]);
      }
    }
  }
  __syncthreads();
  if (MC) cluster_sync_all();
}

int main() {
  const int total_rows = 64 * 1024;  // 8 MB matrix: L2 resident
  __nv_bfloat16* buf;
  cudaMalloc(&buf, (size_t)total_rows * 128);
  cudaMemset(buf, 0, (size_t)total_rows * 128);
  long long* d;
  cudaMalloc(&d, 148 * 8);
  for (int rows : {128, 256}) {
    CUtensorMap mf, mh;
    uint64_t dims[2] = {64, (uint64_t)total_rows};
    uint64_t strides[1] = {128};
    uint32_t bf[2] = {64, (uint32_t)rows}, bh[2] = {64, (uint32_t)(rows / 2)};
    if (make_tmap_bf16(&mf, buf, 2, dims, strides, bf, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap_bf16(&mh, buf, 2, dims, strides, bh, CU_TENSOR_MAP_SWIZZLE_128B)) {
      printf("map failed: %s\n", b200unet_last_error());
      return 1;
    }
    const int smem = kStages * rows * 128 + 1024, iters = 4000, total_tiles = total_rows / rows;
    for (int mc = 0; mc < 2; ++mc) {
      auto kern = mc ? mcast_kernel<true> : mcast_kernel<false>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = mc ? 1 : 0;
      int max_clusters = -1;
      if (mc) cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, kern, mf, mh, rows, iters, total_tiles, d);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) {
          printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
          return 1;
        }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long h[148], mx = 0;
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)iters * rows * 128;
      printf("tile %3d x 64 (%5d B) %s: %7.1f cycles/tile/SM, %6.1f B/cycle LANDING per SM, chip %6.2f TB/s landing%s\n", rows,
             rows * 128, mc ? "multicast x2" : "unicast     ", (double)mx / iters, bytes / (double)mx,
             148.0 * bytes / (ms * 1e-3) / 1e12, mc ? "" : "");
      if (mc) printf("   (max active 2-CTA clusters with this shared-memory footprint: %d of 74)\n", max_clusters);
    }
  }
  return 0;
}
