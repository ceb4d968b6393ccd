// Micro-benchmark 2: how should the build epilogue hand a query's 256-byte tile band to the memory system?
// (tma_store_probe.cu showed that 128-byte rows cap at ~5 TB/s however they are issued, 256-byte rows reach 7.7 TB/s.)
// Persistent CTAs (one per SM), W warps; every warp owns groups of 32 planes (28672 bytes each, cfg2 level 0) and
// visits, patch by patch like the build kernel (56 patches x 2 bands), the 256-byte band of each of its 32 planes:
//   mode 0  two TMA stores of [32 planes][128 B] (SWIZZLE_128B) back to back          (what the round-1 kernel does)
//   mode 1  one TMA store of [32 planes][256 B], no swizzle                           (needs conflict-prone staging)
//   mode 2  one TMA store, 3-D box {128 B, 2 halves, 32 planes}, SWIZZLE_128B         (halves = adjacent smem rows)
//   mode 3  one TMA store, 3-D box {128 B, 32 planes, 2 halves}, SWIZZLE_128B         (halves 32 rows apart)
//   mode 4  st.global: one 128-byte line per instruction, half 0 of all planes, then half 1
//   mode 5  st.global: both 128-byte lines of a plane back to back
//   mode 6  st.global.v2: lane = 8 bytes, one instruction = 256 bytes of one plane
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_store_probe2 tools/probe/tma_store_probe2.cu -lcuda
//   ./tma_store_probe2 <mode> <warps> <ring depth>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int PLANE_BYTES = 28672, PLANES = 56320, VISITS = 112;  // 56 patches x 2 bands of 256 bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DEPTH>
__device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory"); }

__global__ void __launch_bounds__(512, 1)
probe(const __grid_constant__ CUtensorMap map, int mode, int depth, int groups_per_warp, float* gbase) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int box_bytes = mode == 0 ? 4096 : 8192;
  unsigned char* ring = smem + (size_t)warp * depth * box_bytes;
  const int gw = blockIdx.x * nwarps + warp;
  int buf = 0;
  for (int g = 0; g < groups_per_warp; ++g) {
    // work unit = a quarter of the visits of one plane group (32 planes), units round-robin over all warps
    const int u = gw + g * 148 * nwarps;
    const int pg = u >> 2;
    if (pg >= PLANES / 32) break;
    for (int v = (u & 3) * (VISITS / 4); v < ((u & 3) + 1) * (VISITS / 4); ++v) {
      const int patch = v >> 1, band = v & 1, px = patch & 7, py = patch >> 3;
      const int off = (py * 2 + band) * 2048 + px * 256;  // byte offset of the band inside a plane
      if (mode >= 4) {
        float* dst = gbase + (size_t)pg * 32 * (PLANE_BYTES / 4) + off / 4;
        if (mode == 4) {
#pragma unroll 1
          for (int h = 0; h < 2; ++h)
#pragma unroll 8
            for (int r = 0; r < 32; ++r) dst[(size_t)r * (PLANE_BYTES / 4) + h * 32 + lane] = (float)(v + r);
        } else if (mode == 5) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            dst[(size_t)r * (PLANE_BYTES / 4) + lane] = (float)(v + r);
            dst[(size_t)r * (PLANE_BYTES / 4) + 32 + lane] = (float)(v - r);
          }
        } else {
#pragma unroll 8
          for (int r = 0; r < 32; ++r)
            *reinterpret_cast<float2*>(dst + (size_t)r * (PLANE_BYTES / 4) + 2 * lane) = make_float2((float)v, (float)r);
        }
        continue;
      }
      const int nbox = mode == 0 ? 2 : 1;
      for (int h = 0; h < nbox; ++h) {
        if (lane == 0) {
          switch (depth) {
            case 1: wait_read<1>(); break;
            case 2: wait_read<2>(); break;
            case 4: wait_read<4>(); break;
            default: wait_read<8>(); break;
          }
        }
        __syncwarp();
        unsigned char* sb = ring + buf * box_bytes;
        // touch the buffer like the epilogue does and make it visible to the async proxy
        *reinterpret_cast<float4*>(sb + lane * 16) = make_float4(1.f, 2.f, 3.f, (float)v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (mode == 0 || mode == 1) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map),
                         "r"(smem_u32(sb)), "r"((off + h * 128) / 4), "r"(pg * 32)
                         : "memory");
          } else if (mode == 2) {  // dims {32 floats, halves, planes}
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map),
                         "r"(smem_u32(sb)), "r"(0), "r"(off / 128), "r"(pg * 32)
                         : "memory");
          } else {  // mode 3: dims {32 floats, planes, halves}
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map),
                         "r"(smem_u32(sb)), "r"(0), "r"(pg * 32), "r"(off / 128)
                         : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (++buf == depth) buf = 0;
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int warps = argc > 2 ? atoi(argv[2]) : 4;
  const int depth = argc > 3 ? atoi(argv[3]) : 2;
  size_t total = (size_t)PLANES * PLANE_BYTES;
  void* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 0, total));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  CUtensorMap map;
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  if (mode == 0 || mode == 1 || mode >= 4) {
    const int row = mode == 1 ? 256 : 128;
    cuuint64_t dims[2] = {PLANE_BYTES / 4, PLANES};
    cuuint64_t str[1] = {PLANE_BYTES};
    cuuint32_t box[2] = {(cuuint32_t)row / 4, 32};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            mode == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (mode == 2) {
    cuuint64_t dims[3] = {32, PLANE_BYTES / 128, PLANES};
    cuuint64_t str[2] = {128, PLANE_BYTES};
    cuuint32_t box[3] = {32, 2, 32};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {32, PLANES, PLANE_BYTES / 128};
    cuuint64_t str[2] = {PLANE_BYTES, 128};
    cuuint32_t box[3] = {32, 32, 2};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int groups = PLANES / 32;                                   // 1760 plane groups
  const int per_warp = (groups * 4 + 148 * warps - 1) / (148 * warps);
  size_t smem = (size_t)warps * depth * (mode == 0 ? 4096 : 8192);
  if (mode >= 4) smem = 0;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) probe<<<148, warps * 32, smem>>>(map, mode, depth, per_warp, (float*)buf);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int it = 0; it < reps; ++it) probe<<<148, warps * 32, smem>>>(map, mode, depth, per_warp, (float*)buf);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double bytes = (double)groups * 32 * VISITS * 256;
  printf("mode %d, %2d warps, depth %d: %.1f us, %.0f GB/s (%.2f GB written)\n", mode, warps, depth, ms / reps * 1e3,
         bytes / (ms / reps * 1e-3) / 1e9, bytes / 1e9);
  return 0;
}
