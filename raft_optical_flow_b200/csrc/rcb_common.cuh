// Shared device/host helpers for libraftcorr_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/raft_corr_b200.h"

#define RCB_DEVINL __device__ __forceinline__

namespace rcb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Debug / profiling hooks (timing experiments of tools/time_*.py: skip phases of a kernel, cycle counters written
// through a raw device pointer, alternative tunings).  They exist ONLY in a library compiled with -DRCB_DEBUG
// (`python -m raft_optical_flow_b200.build --debug` -> libraftcorr_b200_debug.so); the release library never
// reads the environment, so a stray variable in a job's environment cannot change results or touch memory.
#ifdef RCB_DEBUG
#include <cstdlib>
inline int debug_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
inline unsigned long long* debug_env_ptr(const char* name) {
  const char* e = getenv(name);
  return e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr;
}
#else
inline int debug_env_int(const char*, int dflt) { return dflt; }
inline unsigned long long* debug_env_ptr(const char*) { return nullptr; }
#endif

inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? RCB_OK : (int)e;
}

// Pyramid geometry handed to kernels by value.
struct PyramidDev {
  const void* ptr[RCB_MAX_LEVELS];
  int H[RCB_MAX_LEVELS];
  int W[RCB_MAX_LEVELS];
  int tiles_x[RCB_MAX_LEVELS];
  long long plane_stride[RCB_MAX_LEVELS];
};

// Offset (in fp32 elements) of plane element (y, x) in the 4x4-tiled layout (see rcb_pyramid_layout).
__host__ __device__ __forceinline__ long long tile_off(int y, int x, int tiles_x) {
  return ((long long)((y >> 2) * tiles_x + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3);
}

inline PyramidDev make_pyramid_dev(const void* const* ptrs, const rcb_pyramid_layout& lay) {
  PyramidDev pd;
  for (int l = 0; l < RCB_MAX_LEVELS; ++l) {
    pd.ptr[l] = (ptrs && l < lay.levels) ? ptrs[l] : nullptr;
    pd.H[l] = lay.H[l];
    pd.W[l] = lay.W[l];
    pd.tiles_x[l] = lay.tiles_x[l];
    pd.plane_stride[l] = lay.plane_stride[l];
  }
  return pd;
}

inline int fill_layout(int B, int H, int W, int levels, int dtype, rcb_pyramid_layout* lay) {
  if (!lay || B <= 0 || H <= 0 || W <= 0) return RCB_ERR_INVALID_ARGUMENT;
  if (levels < 1 || levels > RCB_MAX_LEVELS) return RCB_ERR_UNSUPPORTED;
  if (dtype != RCB_F32 && dtype != RCB_F16) return RCB_ERR_UNSUPPORTED;
  const int esize = dtype == RCB_F32 ? 4 : 2;
  const int tile_w = 16 / esize;  // 4 rows x tile_w columns = 64 bytes
  lay->levels = levels;
  lay->dtype = dtype;
  lay->tile_w = tile_w;
  lay->reserved = 0;
  int h = H, w = W;
  for (int l = 0; l < RCB_MAX_LEVELS; ++l) {
    if (l < levels) {
      if (h < 1 || w < 1) return RCB_ERR_INVALID_ARGUMENT;  // pooled away (reference would fail too)
      lay->H[l] = h;
      lay->W[l] = w;
      lay->tiles_x[l] = (w + tile_w - 1) / tile_w;
      lay->tiles_y[l] = (h + 3) / 4;
      lay->plane_stride[l] = (long long)lay->tiles_x[l] * lay->tiles_y[l] * 4 * tile_w;
      lay->level_bytes[l] = (long long)B * H * W * lay->plane_stride[l] * esize;
      h /= 2;
      w /= 2;
    } else {
      lay->H[l] = lay->W[l] = lay->tiles_x[l] = lay->tiles_y[l] = 0;
      lay->plane_stride[l] = lay->level_bytes[l] = 0;
    }
  }
  return RCB_OK;
}

// ---- async copy (LDGSTS) with zero fill -------------------------------------------------
RCB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte global->shared copy; bytes beyond src_bytes are written as zero.
// .cg bypasses L1 (every lane's 16 bytes is its own L2 request); .ca goes through the L1 tag stage, which merges
// the lanes of one instruction that hit the same 32-byte sector into one request.
RCB_DEVINL void cp_async16_zfill(uint32_t dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "r"(src_bytes)
               : "memory");
}
RCB_DEVINL void cp_async16_ca_zfill(uint32_t dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "r"(src_bytes)
               : "memory");
}
RCB_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
RCB_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

RCB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rcb

// Launchers implemented in the individual .cu files (called from c_abi.cu).
namespace rcb {
int launch_build_simt(const float* f1, const float* f2, void* const* pyr, const rcb_pyramid_layout& lay, int B,
                      int C, int H, int W, cudaStream_t s);
int launch_pool_levels(void* const* pyr, const rcb_pyramid_layout& lay, int B, int H, int W, cudaStream_t s);
size_t build_tc_workspace_bytes(int B, int C, int H, int W, int mode);
int launch_build_tc(const float* f1, const float* f2, void* const* pyr, const rcb_pyramid_layout& lay, int B,
                    int C, int H, int W, int mode, void* ws, size_t ws_bytes, cudaStream_t s);
int launch_pack_tc(const float* f1, const float* f2, int B, int C, int H, int W, int mode, void* ws, size_t ws_bytes,
                   cudaStream_t s);
int launch_build_tc_packed(const void* ws, size_t ws_bytes, void* const* pyr, const rcb_pyramid_layout& lay, int B,
                           int C, int H, int W, int mode, cudaStream_t s);
int launch_lookup(const void* const* pyr, const rcb_pyramid_layout& lay, const float* coords, float* out, int B,
                  int H, int W, int radius, cudaStream_t s);
size_t lookup_plan_bytes();
int lookup_plan_init(void* plan, size_t plan_bytes, const void* const* pyr, const rcb_pyramid_layout& lay, int B,
                     int H, int W, int radius);
int launch_lookup_planned(const void* plan, const float* coords, float* out, cudaStream_t s);
int lookup_plan_set_lanes(void* plan, int lanes);
int launch_lookup_backward(const void* const* pyr, const rcb_pyramid_layout& lay, const float* coords,
                           const float* grad_out, float* const* dpyr, float* dcoords, int B, int H, int W,
                           int radius, cudaStream_t s);
int launch_pool_backward(float* const* dpyr, const rcb_pyramid_layout& lay, int B, int H, int W, cudaStream_t s);
int launch_contract_backward(const float* f1, const float* f2, const float* dvol0, float* df1, float* df2, int B,
                             int C, int H, int W, cudaStream_t s);
size_t contract_backward_tc_workspace_bytes(int B, int C, int H, int W);
int launch_contract_backward_tc(const float* f1, const float* f2, const float* dvol0, float* df1, float* df2, int B,
                                int C, int H, int W, void* ws, size_t ws_bytes, cudaStream_t s);
int launch_altcorr_forward(const float* f1, const float* f2, const float* coords, float* corr, int B, int N, int H1,
                           int W1, int H2, int W2, int C, int r, cudaStream_t s);
int launch_altcorr_backward(const float* f1, const float* f2, const float* coords, const float* cg, float* g1,
                            float* g2, float* gc, int B, int N, int H1, int W1, int H2, int W2, int C, int r,
                            int true_cg, cudaStream_t s);
int launch_upsample_flow(const float* flow, const float* mask, float* out, int N, int H, int W, cudaStream_t s);
int launch_upsample_flow_backward(const float* flow, const float* mask, const float* gout, float* dflow, float* dmask,
                                  float* workspace, int N, int H, int W, cudaStream_t s);
size_t convc1_pack_bytes(int cout, int levels, int radius);
int launch_convc1_pack(const float* weight, void* wpack, int cout, int levels, int radius, cudaStream_t s);
int launch_lookup_convc1(const void* plan, const float* coords, const void* wpack, const float* bias, float* out,
                         int cout, int relu, cudaStream_t s);
int launch_altcorr_prepare(const float* f1, const float* f2, float* f1n, float* const* f2n, int B, int C, int H,
                           int W, int levels, cudaStream_t s);
int launch_altcorr_pyramid_forward(const float* f1n, const float* const* f2n, const float* coords, float* out,
                                   int B, int C, int H, int W, int levels, int r, float scale, cudaStream_t s);
}  // namespace rcb
