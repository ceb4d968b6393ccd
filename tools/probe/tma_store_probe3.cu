// Micro-benchmark 3: does the PLANE STRIDE matter for the build epilogue's store pattern?  Same visit pattern as
// tma_store_probe2.cu (persistent CTAs, W warps, a warp owns groups of 32 planes and writes the 256-byte band of each
// of them patch by patch), rows issued as plain coalesced st.global lines (mode 5 there: as fast as the TMA boxes),
// with the distance between planes a run-time parameter (28672 bytes = cfg2 level 0; padded variants).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_store_probe3 tools/probe/tma_store_probe3.cu
//   ./tma_store_probe3 <plane stride bytes> <warps> [row bytes: 128 | 256 | 512]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int PLANES = 56320, PLANE_DATA = 28672;

__global__ void __launch_bounds__(512, 1)
probe(float* gbase, int stride4, int row4, int groups_per_warp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gw = blockIdx.x * nwarps + warp;
  const int visits = PLANE_DATA / 4 / row4;  // rows of row4 floats per plane
  for (int g = 0; g < groups_per_warp; ++g) {
    const int u = gw + g * 148 * nwarps;
    const int pg = u >> 2;
    if (pg >= PLANES / 32) break;
    for (int v = (u & 3) * (visits / 4); v < ((u & 3) + 1) * (visits / 4); ++v) {
      float* dst = gbase + (size_t)pg * 32 * stride4 + (size_t)v * row4;
#pragma unroll 4
      for (int r = 0; r < 32; ++r)
        for (int c = lane; c < row4; c += 32) dst[(size_t)r * stride4 + c] = (float)(v + r);
    }
  }
}

int main(int argc, char** argv) {
  const int stride = argc > 1 ? atoi(argv[1]) : PLANE_DATA;
  const int warps = argc > 2 ? atoi(argv[2]) : 8;
  const int row = argc > 3 ? atoi(argv[3]) : 256;
  size_t total = (size_t)PLANES * stride;
  void* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 0, total));
  const int groups = PLANES / 32;
  const int per_warp = (groups * 4 + 148 * warps - 1) / (148 * warps);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) probe<<<148, warps * 32>>>((float*)buf, stride / 4, row / 4, per_warp);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int it = 0; it < reps; ++it) probe<<<148, warps * 32>>>((float*)buf, stride / 4, row / 4, per_warp);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double bytes = (double)PLANES * PLANE_DATA;
  printf("stride %6d (+%5d), row %3d B, %2d warps: %.1f us, %.0f GB/s\n", stride, stride - PLANE_DATA, row, warps,
         ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
  return 0;
}
