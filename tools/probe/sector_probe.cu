// Is DRAM read traffic 32-byte-sector or 64-byte granular?  Reads the first 32 bytes of every 64-byte (or 128-byte)
// block of a 2 GB buffer once (no reuse, far larger than L2); run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./sector_probe
// and compare dram__bytes_read with the bytes requested (printed).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/sector_probe tools/probe/sector_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void read_part(const uint4* __restrict__ p, size_t nblocks, int block16, int take16, float* sink) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nblocks * take16; i += (size_t)gridDim.x * blockDim.x) {
    const size_t blk = i / take16, part = i % take16;
    const uint4 v = __ldg(p + blk * block16 + part);
    acc += __uint_as_float(v.x) + __uint_as_float(v.y) + __uint_as_float(v.z) + __uint_as_float(v.w);
  }
  if (acc == 123.456f) *sink = acc;
}

int main() {
  const size_t bytes = 2ull << 30;
  void* buf; float* sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, bytes);
  struct { int block16, take16; const char* what; } cases[] = {
      {4, 4, "64 of every 64 bytes"}, {4, 2, "first 32 of every 64 bytes"}, {4, 1, "first 16 of every 64 bytes"},
      {8, 2, "first 32 of every 128 bytes"}, {8, 4, "first 64 of every 128 bytes"}};
  for (auto& c : cases) {
    const size_t nblocks = bytes / (16 * c.block16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    read_part<<<148 * 16, 256>>>((const uint4*)buf, nblocks, c.block16, c.take16, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-30s requested %.1f MB  %.1f us  %.0f GB/s of requested bytes\n", c.what, nblocks * 16.0 * c.take16 / 1e6,
           ms * 1e3, nblocks * 16.0 * c.take16 / ms / 1e6);
  }
  return 0;
}
