// Micro-benchmark: per-SM TMA load throughput for NHWC activation boxes as a function of the inner (channel) extent,
// boxes per barrier (G), boxes in flight (stages) and number of independent producer warps (P).
// One CTA per SM.  Producer warp w (lane 0) issues loads into its own ring; consumer warp 4+w frees slots as they land.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../unet-implementations_b200/csrc tma_rate.cu ../../unet-implementations_b200/csrc/api.cu -o tma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "common.cuh"
#include "ptx.cuh"
#include "conv_common.cuh"
using namespace b200;

constexpr int MAXS = 8, MAXP = 8;

__global__ void __launch_bounds__(512, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int G,
                                                      int stages, int P, int iters, int tiles_w, int tiles_h, int boxH,
                                                      int nimg, long long* cycles, long long* issue_cycles, int rank2) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAXP][MAXS], empty_bar[MAXP][MAXS];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) {
    for (int p = 0; p < P; ++p)
      for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[p][s], 1); mbar_init(&empty_bar[p][s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_img = tiles_w * tiles_h;
  const int total = per_img * nimg;
  if (warp < P && lane == 0) {
    const int p = warp;
    uint8_t* base = smem + (size_t)p * stages * G * box_bytes;
    int t = (int)(((long long)(blockIdx.x * P + p) * iters * G) % total);
    long long t0 = clock64(), iss = 0;
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mbar_wait(&empty_bar[p][s], ((i / stages) & 1) ^ 1);
      mbar_expect_tx(&full_bar[p][s], box_bytes * G);
      long long a = clock64();
      for (int g = 0; g < G; ++g) {
        const int n = t / per_img, r = t - n * per_img;
        if (rank2)
          tma_load_2d(base + (size_t)(s * G + g) * box_bytes, &map, &full_bar[p][s], 0, t * 16 * boxH);
        else
          tma_load_4d(base + (size_t)(s * G + g) * box_bytes, &map, &full_bar[p][s], 0, (r % tiles_w) * 16,
                      (r / tiles_w) * boxH, n);
        if (++t == total) t = 0;
      }
      iss += clock64() - a;
    }
    cycles[blockIdx.x * MAXP + p] = clock64() - t0;
    issue_cycles[blockIdx.x * MAXP + p] = iss;
  } else if (warp >= 8 && warp < 8 + P && lane == 0) {
    const int p = warp - 8;
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mbar_wait(&full_bar[p][s], (i / stages) & 1);
      mbar_arrive(&empty_bar[p][s]);
    }
  }
  __syncthreads();
}

int main() {
  const int N = 32, H = 512, W = 512;
  long long *d, *d2;
  cudaMalloc(&d, 148 * MAXP * 8);
  cudaMalloc(&d2, 148 * MAXP * 8);
  for (int C : {32, 64}) {
    __nv_bfloat16* buf;
    size_t bytes = (size_t)N * H * W * C * 2;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    for (int nimg : {32, 1})
    for (int rank2 : {0, 1})
      for (int boxH : {4, 8, 16}) {
        CUtensorMap map;
        const int box_bytes = C * 2 * 16 * boxH;
        if (rank2) {
          uint64_t dims[2] = {(uint64_t)C, (uint64_t)N * H * W};
          uint64_t strides[1] = {(uint64_t)C * 2};
          uint32_t box[2] = {(uint32_t)C, (uint32_t)(16 * boxH)};
          if (make_tmap_bf16(&map, buf, 2, dims, strides, box, C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) { printf("map2 failed: %s\n", b200unet_last_error()); return 1; }
        } else if (make_act_map(&map, buf, C, N, H, W, C, 1, 1, 0, 0, C, 16, boxH)) { printf("map failed: %s\n", b200unet_last_error()); return 1; }
        for (int P : {1, 4, 8})
          for (int G : {1, 2}) {
            const int stages = 4;
            const int smem = P * stages * G * box_bytes + 1024;
            if (smem > 220 * 1024) continue;
            cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            const int tiles_w = W / 16, tiles_h = H / boxH;
            const int iters = 2000;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
              cudaEventRecord(e0);
              tma_kernel<<<148, 512, smem>>>(map, box_bytes, G, stages, P, iters, tiles_w, tiles_h, boxH, nimg, d, d2, rank2);
              cudaEventRecord(e1);
              if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
              cudaEventElapsedTime(&ms, e0, e1);
            }
            long long h[148 * MAXP], h2[148 * MAXP];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost);
            long long mx = 0, is = 0;
            for (int i = 0; i < 148; ++i) for (int p = 0; p < P; ++p) { if (h[i * MAXP + p] > mx) mx = h[i * MAXP + p]; if (h2[i * MAXP + p] > is) is = h2[i * MAXP + p]; }
            const double boxes = (double)iters * G * P;
            printf("imgs=%2d C=%2d rank%d box %3d rows (%5d B) P=%d G=%d: %7.1f cycles/box/SM  issue %6.1f cycles/box  %6.1f B/cycle/SM  chip %6.2f TB/s\n",
                   nimg, C, rank2 ? 2 : 4, 16 * boxH, box_bytes, P, G, (double)mx / (iters * G * P), (double)is / (iters * G),
                   boxes * box_bytes / (double)mx, 148.0 * boxes * box_bytes / (ms * 1e-3) / 1e12);
          }
      }
    cudaFree(buf);
  }
  return 0;
}
