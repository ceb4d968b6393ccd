// extern "C" surface of libraftcorr_b200.so (declared in include/raft_corr_b200.h): argument
// validation and dispatch to the kernel launchers.  No allocation, no global state.
#include "rcb_common.cuh"

using namespace rcb;

namespace {
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool radius_ok(int r) { return r >= 1 && r <= RCB_MAX_RADIUS; }
}  // namespace

extern "C" {

int rcb_abi_version(void) { return RCB_ABI_VERSION; }

const char* rcb_status_string(int status) {
  switch (status) {
    case RCB_OK: return "ok";
    case RCB_ERR_INVALID_ARGUMENT: return "invalid argument (null/misaligned pointer or non-positive size)";
    case RCB_ERR_UNSUPPORTED: return "unsupported radius / level count / channel count / dtype / mode";
    case RCB_ERR_WORKSPACE: return "workspace too small";
    case RCB_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown status";
  }
}

int rcb_pyramid_layout_query(int B, int H, int W, int levels, int dtype, rcb_pyramid_layout* layout) {
  return fill_layout(B, H, W, levels, dtype, layout);
}

size_t rcb_corr_build_workspace_bytes(int B, int C, int H, int W, int mode) {
  if (mode == RCB_BUILD_FP32_SIMT) return 0;
  return build_tc_workspace_bytes(B, C, H, W, mode);
}

int rcb_corr_build(const float* fmap1, const float* fmap2, void* const* pyr, int B, int C, int H, int W,
                   int levels, int mode, int pyr_dtype, void* workspace, size_t workspace_bytes,
                   rcb_stream_t stream) {
  if (!fmap1 || !fmap2 || !pyr || C <= 0) return RCB_ERR_INVALID_ARGUMENT;
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, levels, pyr_dtype, &lay);
  if (st != RCB_OK) return st;
  if (!aligned16(fmap1) || !aligned16(fmap2)) return RCB_ERR_INVALID_ARGUMENT;
  for (int l = 0; l < levels; ++l)
    if (!pyr[l] || !aligned16(pyr[l])) return RCB_ERR_INVALID_ARGUMENT;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (mode) {
    case RCB_BUILD_FP32_SIMT: return launch_build_simt(fmap1, fmap2, pyr, lay, B, C, H, W, s);
    case RCB_BUILD_BF16X3:
    case RCB_BUILD_BF16:
    case RCB_BUILD_F16F8: return launch_build_tc(fmap1, fmap2, pyr, lay, B, C, H, W, mode, workspace, workspace_bytes, s);
    default: return RCB_ERR_UNSUPPORTED;
  }
}

int rcb_corr_pack_fmaps(const float* fmaps, void* packed, size_t packed_bytes, int B, int C, int H, int W, int mode,
                        rcb_stream_t stream) {
  if (!fmaps || !packed || B <= 0 || C <= 0 || H <= 0 || W <= 0 || !aligned16(fmaps)) return RCB_ERR_INVALID_ARGUMENT;
  if (mode != RCB_BUILD_BF16X3 && mode != RCB_BUILD_BF16 && mode != RCB_BUILD_F16F8) return RCB_ERR_UNSUPPORTED;
  return launch_pack_tc(fmaps, fmaps + (size_t)B * C * H * W, B, C, H, W, mode, packed, packed_bytes,
                        reinterpret_cast<cudaStream_t>(stream));
}

int rcb_corr_build_packed(const void* packed, size_t packed_bytes, void* const* pyr, int B, int C, int H, int W,
                          int levels, int mode, int pyr_dtype, rcb_stream_t stream) {
  if (!packed || !pyr || C <= 0) return RCB_ERR_INVALID_ARGUMENT;
  if (mode != RCB_BUILD_BF16X3 && mode != RCB_BUILD_BF16 && mode != RCB_BUILD_F16F8) return RCB_ERR_UNSUPPORTED;
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, levels, pyr_dtype, &lay);
  if (st != RCB_OK) return st;
  for (int l = 0; l < levels; ++l)
    if (!pyr[l] || !aligned16(pyr[l])) return RCB_ERR_INVALID_ARGUMENT;
  return launch_build_tc_packed(packed, packed_bytes, pyr, lay, B, C, H, W, mode, reinterpret_cast<cudaStream_t>(stream));
}

int rcb_corr_lookup(const void* const* pyr, const float* coords, float* out, int B, int H, int W, int levels,
                    int radius, int pyr_dtype, rcb_stream_t stream) {
  if (!pyr || !coords || !out) return RCB_ERR_INVALID_ARGUMENT;
  if (!radius_ok(radius)) return RCB_ERR_UNSUPPORTED;
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, levels, pyr_dtype, &lay);
  if (st != RCB_OK) return st;
  for (int l = 0; l < levels; ++l)
    if (!pyr[l] || !aligned16(pyr[l])) return RCB_ERR_INVALID_ARGUMENT;
  return launch_lookup(pyr, lay, coords, out, B, H, W, radius, reinterpret_cast<cudaStream_t>(stream));
}

size_t rcb_corr_lookup_plan_bytes(void) { return lookup_plan_bytes(); }

int rcb_corr_lookup_plan_init(void* plan, size_t plan_bytes, const void* const* pyr, int B, int H, int W,
                              int levels, int radius, int pyr_dtype) {
  if (!plan || !pyr) return RCB_ERR_INVALID_ARGUMENT;
  if (!radius_ok(radius)) return RCB_ERR_UNSUPPORTED;
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, levels, pyr_dtype, &lay);
  if (st != RCB_OK) return st;
  for (int l = 0; l < levels; ++l)
    if (!pyr[l] || !aligned16(pyr[l])) return RCB_ERR_INVALID_ARGUMENT;
  return lookup_plan_init(plan, plan_bytes, pyr, lay, B, H, W, radius);
}

int rcb_corr_lookup_plan_set_lanes(void* plan, int lanes_per_query) { return lookup_plan_set_lanes(plan, lanes_per_query); }

int rcb_corr_lookup_planned(const void* plan, const float* coords, float* out, rcb_stream_t stream) {
  if (!plan || !coords || !out) return RCB_ERR_INVALID_ARGUMENT;
  return launch_lookup_planned(plan, coords, out, reinterpret_cast<cudaStream_t>(stream));
}

int rcb_corr_lookup_backward(const void* const* pyr, const float* coords, const float* grad_out,
                             float* const* dpyr, float* dcoords, int B, int H, int W, int levels, int radius,
                             int pyr_dtype, rcb_stream_t stream) {
  if (!pyr || !coords || !grad_out || (!dpyr && !dcoords)) return RCB_ERR_INVALID_ARGUMENT;
  if (!radius_ok(radius)) return RCB_ERR_UNSUPPORTED;
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, levels, pyr_dtype, &lay);
  if (st != RCB_OK) return st;
  for (int l = 0; l < levels; ++l)
    if (!pyr[l] || (dpyr && !dpyr[l])) return RCB_ERR_INVALID_ARGUMENT;
  return launch_lookup_backward(pyr, lay, coords, grad_out, dpyr, dcoords, B, H, W, radius,
                                reinterpret_cast<cudaStream_t>(stream));
}

int rcb_corr_pool_backward(float* const* dpyr, int B, int H, int W, int levels, rcb_stream_t stream) {
  if (!dpyr) return RCB_ERR_INVALID_ARGUMENT;
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, levels, RCB_F32, &lay);
  if (st != RCB_OK) return st;
  for (int l = 0; l < levels; ++l)
    if (!dpyr[l] || !aligned16(dpyr[l])) return RCB_ERR_INVALID_ARGUMENT;
  return launch_pool_backward(dpyr, lay, B, H, W, reinterpret_cast<cudaStream_t>(stream));
}

int rcb_corr_contract_backward(const float* fmap1, const float* fmap2, const float* dvol0, float* dfmap1,
                               float* dfmap2, int B, int C, int H, int W, rcb_stream_t stream) {
  if (!fmap1 || !fmap2 || !dvol0 || !dfmap1 || !dfmap2 || B <= 0 || C <= 0 || H <= 0 || W <= 0)
    return RCB_ERR_INVALID_ARGUMENT;
  return launch_contract_backward(fmap1, fmap2, dvol0, dfmap1, dfmap2, B, C, H, W,
                                  reinterpret_cast<cudaStream_t>(stream));
}

size_t rcb_corr_contract_backward_tc_workspace_bytes(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return contract_backward_tc_workspace_bytes(B, C, H, W);
}

int rcb_corr_contract_backward_tc(const float* fmap1, const float* fmap2, const float* dvol0, float* dfmap1,
                                  float* dfmap2, int B, int C, int H, int W, void* workspace,
                                  size_t workspace_bytes, rcb_stream_t stream) {
  if (!fmap1 || !fmap2 || !dvol0 || !dfmap1 || !dfmap2 || B <= 0 || C <= 0 || H <= 0 || W <= 0)
    return RCB_ERR_INVALID_ARGUMENT;
  if (!aligned16(dvol0)) return RCB_ERR_INVALID_ARGUMENT;
  return launch_contract_backward_tc(fmap1, fmap2, dvol0, dfmap1, dfmap2, B, C, H, W, workspace, workspace_bytes,
                                     reinterpret_cast<cudaStream_t>(stream));
}

int rcb_altcorr_forward(const float* fmap1, const float* fmap2, const float* coords, float* corr, int B, int N,
                        int H1, int W1, int H2, int W2, int C, int radius, rcb_stream_t stream) {
  if (!fmap1 || !fmap2 || !coords || !corr) return RCB_ERR_INVALID_ARGUMENT;
  if (B <= 0 || N <= 0 || H1 <= 0 || W1 <= 0 || H2 <= 0 || W2 <= 0 || C <= 0) return RCB_ERR_INVALID_ARGUMENT;
  if (!radius_ok(radius)) return RCB_ERR_UNSUPPORTED;
  if (!aligned16(fmap1) || !aligned16(fmap2) || !aligned16(coords)) return RCB_ERR_INVALID_ARGUMENT;
  return launch_altcorr_forward(fmap1, fmap2, coords, corr, B, N, H1, W1, H2, W2, C, radius,
                                reinterpret_cast<cudaStream_t>(stream));
}

int rcb_altcorr_backward(const float* fmap1, const float* fmap2, const float* coords, const float* corr_grad,
                         float* fmap1_grad, float* fmap2_grad, float* coords_grad, int B, int N, int H1, int W1,
                         int H2, int W2, int C, int radius, int true_coords_grad, rcb_stream_t stream) {
  if (!fmap1 || !fmap2 || !coords || !corr_grad || !fmap1_grad || !fmap2_grad || !coords_grad)
    return RCB_ERR_INVALID_ARGUMENT;
  if (B <= 0 || N <= 0 || H1 <= 0 || W1 <= 0 || H2 <= 0 || W2 <= 0 || C <= 0) return RCB_ERR_INVALID_ARGUMENT;
  if (!radius_ok(radius)) return RCB_ERR_UNSUPPORTED;
  if (!aligned16(fmap1) || !aligned16(fmap2) || !aligned16(coords) || !aligned16(fmap1_grad) ||
      !aligned16(fmap2_grad) || !aligned16(coords_grad))
    return RCB_ERR_INVALID_ARGUMENT;
  return launch_altcorr_backward(fmap1, fmap2, coords, corr_grad, fmap1_grad, fmap2_grad, coords_grad, B, N, H1, W1,
                                 H2, W2, C, radius, true_coords_grad, reinterpret_cast<cudaStream_t>(stream));
}

int rcb_altcorr_prepare(const float* fmap1, const float* fmap2, float* fmap1_nhwc, float* const* fmap2_nhwc, int B,
                        int C, int H, int W, int levels, rcb_stream_t stream) {
  if (!fmap1 || !fmap2 || !fmap1_nhwc || !fmap2_nhwc || B <= 0 || C <= 0 || H <= 0 || W <= 0)
    return RCB_ERR_INVALID_ARGUMENT;
  if (levels < 1 || levels > RCB_MAX_LEVELS) return RCB_ERR_UNSUPPORTED;
  for (int l = 0; l < levels; ++l)
    if (!fmap2_nhwc[l]) return RCB_ERR_INVALID_ARGUMENT;
  return launch_altcorr_prepare(fmap1, fmap2, fmap1_nhwc, fmap2_nhwc, B, C, H, W, levels,
                                reinterpret_cast<cudaStream_t>(stream));
}

int rcb_altcorr_pyramid_forward(const float* fmap1_nhwc, const float* const* fmap2_nhwc, const float* coords,
                                float* out, int B, int C, int H, int W, int levels, int radius, float scale,
                                rcb_stream_t stream) {
  if (!fmap1_nhwc || !fmap2_nhwc || !coords || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0)
    return RCB_ERR_INVALID_ARGUMENT;
  if (levels < 1 || levels > RCB_MAX_LEVELS || !radius_ok(radius)) return RCB_ERR_UNSUPPORTED;
  if ((H >> (levels - 1)) < 1 || (W >> (levels - 1)) < 1) return RCB_ERR_INVALID_ARGUMENT;
  for (int l = 0; l < levels; ++l)
    if (!fmap2_nhwc[l] || !aligned16(fmap2_nhwc[l])) return RCB_ERR_INVALID_ARGUMENT;
  if (!aligned16(fmap1_nhwc)) return RCB_ERR_INVALID_ARGUMENT;
  return launch_altcorr_pyramid_forward(fmap1_nhwc, fmap2_nhwc, coords, out, B, C, H, W, levels, radius, scale,
                                        reinterpret_cast<cudaStream_t>(stream));
}

int rcb_upsample_flow(const float* flow, const float* mask, float* out, int N, int H, int W, rcb_stream_t stream) {
  if (!flow || !mask || !out || N <= 0 || H <= 0 || W <= 0) return RCB_ERR_INVALID_ARGUMENT;
  if (!aligned16(out)) return RCB_ERR_INVALID_ARGUMENT;
  return launch_upsample_flow(flow, mask, out, N, H, W, reinterpret_cast<cudaStream_t>(stream));
}

size_t rcb_upsample_flow_backward_workspace_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)N * 2 * 9 * H * W * sizeof(float);
}

int rcb_upsample_flow_backward(const float* flow, const float* mask, const float* grad_out, float* dflow,
                               float* dmask, void* workspace, size_t workspace_bytes, int N, int H, int W,
                               rcb_stream_t stream) {
  if (!flow || !mask || !grad_out || !dflow || !dmask || N <= 0 || H <= 0 || W <= 0) return RCB_ERR_INVALID_ARGUMENT;
  if (!aligned16(grad_out)) return RCB_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < rcb_upsample_flow_backward_workspace_bytes(N, H, W)) return RCB_ERR_WORKSPACE;
  return launch_upsample_flow_backward(flow, mask, grad_out, dflow, dmask, static_cast<float*>(workspace), N, H, W,
                                       reinterpret_cast<cudaStream_t>(stream));
}

size_t rcb_corr_convc1_pack_bytes(int cout, int levels, int radius) { return convc1_pack_bytes(cout, levels, radius); }

int rcb_corr_convc1_pack(const float* weight, void* wpack, int cout, int levels, int radius, rcb_stream_t stream) {
  return launch_convc1_pack(weight, wpack, cout, levels, radius, reinterpret_cast<cudaStream_t>(stream));
}

int rcb_corr_lookup_convc1(const void* plan, const float* coords, const void* wpack, const float* bias, float* out,
                           int cout, int relu, rcb_stream_t stream) {
  return launch_lookup_convc1(plan, coords, wpack, bias, out, cout, relu, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
