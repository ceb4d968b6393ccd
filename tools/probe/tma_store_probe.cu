// Micro-benchmark: what does the TMA store path of one SM sustain for the build kernel's store shape?
// Persistent CTAs (one per SM), W warps each issuing cp.async.bulk.tensor stores of [ROWS][ROWB bytes] boxes from a
// per-warp staging ring (contents irrelevant) to a [planes][bytes_per_plane] tensor, box rows = consecutive planes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_store_probe tools/probe/tma_store_probe.cu -lcuda
//   ./tma_store_probe <row_bytes 64|128|256> <warps> <ring depth> <alias planes (0 = none)> <plane_bytes> [run_bytes streams]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DEPTH>
__device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory"); }

__global__ void __launch_bounds__(512, 1)
probe(const __grid_constant__ CUtensorMap map, int row_bytes, int depth, long long boxes_per_warp, int planes_total,
      int cols_per_plane /* boxes along a plane */, int alias, int run_bytes, int streams, int variant, float* gbase,
      int plane_bytes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int box_bytes = 32 * row_bytes;
  unsigned char* ring = smem + (size_t)warp * depth * box_bytes;
  // each warp owns a group of 32 planes; successive boxes walk along the plane (like successive patches)
  long long gw = (long long)blockIdx.x * nwarps + warp;
  int buf = 0;
  for (long long i = 0; i < boxes_per_warp; ++i) {
    long long id = gw * boxes_per_warp + i;
    int col = (int)(id % cols_per_plane);
    long long pg = (id / cols_per_plane) % (planes_total / 32);
    if (run_bytes > 0) {
      // `streams` interleaved sequential streams of `run_bytes` runs (the stores of a 4 x (run_bytes/16)-target patch
      // whose vertical neighbour follows it: streams = 2; a patch row swept left to right: streams = 1); every warp
      // of every CTA walks the same in-plane offsets in lockstep on its own 32 planes (plane = 14 tile rows x 2 KB)
      long long unit = i / cols_per_plane, j = i % cols_per_plane;
      const int bpr = run_bytes / 128, nrun = 2048 / run_bytes;
      int k = (int)(j % bpr); j /= bpr;
      int half = (int)(j % streams); j /= streams;
      int spx = (int)(j % nrun); j /= nrun;
      int byte_off = ((int)j * streams + half) * 2048 + spx * run_bytes + k * 128;
      col = byte_off / row_bytes;
      pg = (gw + unit * 148LL * nwarps) % (planes_total / 32);
    } else if (alias < 0) {
      // the build kernel's pattern: every warp of every CTA walks the SAME in-plane offsets in lockstep (patch by
      // patch: band 0 half 0, band 0 half 1, band 1 half 0, band 1 half 1), each on its own 32 planes
      long long unit = i / cols_per_plane, j = i % cols_per_plane;  // j-th box of this plane group
      int patch = (int)(j >> 2), band = (int)((j >> 1) & 1), half = (int)(j & 1);
      // alias = -S: S stagger groups, group g starts its sweep g * (56 / S) patches later (CTA-level stagger)
      const int S = -alias;
      patch = (patch + (blockIdx.x % S) * (56 / S)) % 56;
      int px = patch % 8, py = patch / 8;                           // 8 patches per patch row (W = 128)
      int byte_off = (py * 2 + band) * 2048 + px * 256 + half * 128;
      col = byte_off / row_bytes;
      pg = (gw + unit * 148LL * nwarps) % (planes_total / 32);
    } else if (alias) pg %= alias;
    if (variant == 4) {
      // transposed-epilogue emulation: no staging, no TMA -- the warp writes one full 128-byte line per plane with
      // plain coalesced st.global (lane = 4 bytes of the line), 32 planes per "box"
      float* dst = gbase + ((long long)pg * 32) * (plane_bytes / 4) + col * (row_bytes / 4) + lane;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) dst[(long long)r * (plane_bytes / 4)] = (float)(i + r);
      continue;
    }
    if (lane == 0) {
      switch (depth) {
        case 1: wait_read<1>(); break;
        case 2: wait_read<2>(); break;
        case 4: wait_read<4>(); break;
        default: wait_read<8>(); break;
      }
    }
    __syncwarp();
    // touch the buffer like the epilogue does (one 16-byte store per lane) and make it visible to the async proxy
    if (variant != 1 && variant != 2)
      *reinterpret_cast<float4*>(ring + buf * box_bytes + lane * 16) = make_float4(1.f, 2.f, 3.f, (float)i);
    if (variant != 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map),
                   "r"(smem_u32(ring + buf * box_bytes)), "r"(col * (row_bytes / 4)), "r"((int)(pg * 32))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (++buf == depth) buf = 0;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  int row_bytes = argc > 1 ? atoi(argv[1]) : 128;
  int warps = argc > 2 ? atoi(argv[2]) : 4;
  int depth = argc > 3 ? atoi(argv[3]) : 2;
  int alias = argc > 4 ? atoi(argv[4]) : 0;
  int plane_bytes = argc > 5 ? atoi(argv[5]) : 28672;
  int run_bytes = argc > 6 ? atoi(argv[6]) : 0;   // > 0: "runs" pattern (row_bytes must be 128, plane_bytes 28672)
  int streams = argc > 7 ? atoi(argv[7]) : 2;
  int variant = argc > 8 ? atoi(argv[8]) : 0;     // 1: no staging write, no fence; 2: fence only; 4: coalesced st.global
  const int planes = 56320;  // cfg2: 8 x 7040 query planes
  size_t total = (size_t)planes * plane_bytes;
  void* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 0, total));
  // driver entry point
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)plane_bytes / 4, (cuuint64_t)planes};
  cuuint64_t str[1] = {(cuuint64_t)plane_bytes};
  cuuint32_t box[2] = {(cuuint32_t)row_bytes / 4, 32};
  cuuint32_t es[2] = {1, 1};
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                          : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  int cols = plane_bytes / row_bytes;
  long long boxes_total = (long long)(planes / 32) * cols;       // covers the buffer exactly once
  long long per_warp = boxes_total / (148LL * warps);
  size_t smem = (size_t)warps * depth * 32 * row_bytes;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) probe<<<148, warps * 32, smem>>>(map, row_bytes, depth, per_warp, planes, cols, alias, run_bytes, streams, variant, (float*)buf, plane_bytes);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int it = 0; it < reps; ++it) probe<<<148, warps * 32, smem>>>(map, row_bytes, depth, per_warp, planes, cols, alias, run_bytes, streams, variant, (float*)buf, plane_bytes);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double bytes = (double)per_warp * 148 * warps * 32 * row_bytes;
  printf("v%d row %3d B, %2d warps, depth %d, alias %4d, run %4d x %d: %.1f us, %.0f GB/s, %.1f B/ns/SM\n", variant, row_bytes, warps, depth, alias, run_bytes, streams,
         ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9, bytes / (ms / reps * 1e-3) / 1e9 / 148);
  return 0;
}
