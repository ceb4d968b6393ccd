/*
 * raft_corr_b200.h -- C ABI of libraftcorr_b200.so, the B200-native (sm_100a) implementation of
 * the RAFT correlation hot path.
 *
 * The reference (wangty537/raft_optical_flow) exposes this path only through Python
 * (core/corr.py) and one pybind11 torch extension (alt_cuda_corr/correlation.cpp:51-54).  This
 * header is the torch-free boundary a binding for either of them calls into: plain device
 * pointers, sizes and a cudaStream_t.  Every entry point
 *   - takes DEVICE pointers to contiguous fp32 tensors (unless a dtype argument says otherwise),
 *   - allocates nothing and keeps no global mutable state (the caller owns every buffer; safe to
 *     call concurrently from one host thread per GPU, as nn.DataParallel does, train.py:172),
 *   - enqueues its kernels on `stream` of the CURRENT device and returns without synchronising,
 *   - returns RCB_OK (0), a negative RCB_ERR_* code for a rejected argument, or a positive
 *     cudaError_t reported by the launch.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Reference paths cited below are relative to the reference checkout.
 */
#ifndef RAFT_CORR_B200_H_
#define RAFT_CORR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCB_ABI_VERSION 9
#define RCB_MAX_LEVELS 4 /* core/raft.py:46-53 fixes corr_levels = 4 */
#define RCB_MAX_RADIUS 4 /* core/raft.py:47,53: radius 3 (small) / 4 (full) */

#define RCB_API __attribute__((visibility("default")))

typedef struct CUstream_st* rcb_stream_t; /* == cudaStream_t */

enum rcb_status {
  RCB_OK = 0,
  RCB_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, misaligned buffer */
  RCB_ERR_UNSUPPORTED = -2,      /* radius / levels / channel count outside the compiled range */
  RCB_ERR_WORKSPACE = -3,        /* workspace smaller than rcb_corr_build_workspace_bytes() */
  RCB_ERR_NO_DEVICE = -4,        /* no sm_100 device / driver entry point missing */
};

/* Storage type of the correlation pyramid. */
enum rcb_dtype {
  RCB_F32 = 0, /* reference behaviour (core/corr.py keeps everything fp32) */
  RCB_F16 = 1, /* fast mode: halves build stores and lookup reads; tensor-core build modes and
                  the lookup only -- the backward entry points and RCB_BUILD_FP32_SIMT return RCB_ERR_UNSUPPORTED */
};

/* Arithmetic of the all-pairs contraction (core/corr.py:121 is a true-fp32 cuBLAS SGEMM). */
enum rcb_build_mode {
  RCB_BUILD_FP32_SIMT = 0, /* fp32 FMA on the SIMT pipe; fp32 accumulate */
  RCB_BUILD_BF16X3 = 1,    /* tcgen05: hi/lo bf16 split, 3 products, fp32 accumulate in TMEM;
                              error ~2e-6 of max-abs -- the default fp32-parity mode */
  RCB_BUILD_BF16 = 2,      /* tcgen05: single bf16 pass, fp32 accumulate; ~1e-3 of max-abs */
  RCB_BUILD_F16F8 = 3,     /* tcgen05: hi = fp16(x) on kind::f16 (one pass) + the two cross terms hi*lo, lo*hi in
                              8-bit e4m3 on kind::f8f6f4 (twice the rate), fp32 accumulate in TMEM: two
                              pass-equivalents instead of three; error ~1e-5 of max-abs -- the default
                              fp32-parity mode.  Feature magnitudes must stay below the fp16 range (65504);
                              use RCB_BUILD_BF16X3 for unbounded inputs */
};

/* Memory layout of the pyramid the build writes and the lookup reads.
 * Level l holds, for every query q in [0, B*H*W), one plane of H_l x W_l correlations
 * (H_l = H_{l-1}/2, W_l = W_{l-1}/2, floor -- core/corr.py:52-54).  A plane is stored as a row-major grid of
 * 64-byte TILES of 4 rows x tile_w columns (tile_w = 4 for fp32, 8 for fp16): the (2r+2)^2 window the lookup
 * gathers then touches ~10.6 DRAM atoms (64 B) per level instead of the ~16 a row-major plane costs, and the
 * build writes 128..256 contiguous bytes per query and tile band.
 *   element (q, y, x) of level l lives at (in elements)
 *     q * plane_stride[l] + ((y / 4) * tiles_x[l] + x / tile_w) * (4 * tile_w) + (y % 4) * tile_w + x % tile_w
 * Elements of edge tiles that lie outside H_l x W_l are padding: never read as data, contents unspecified. */
typedef struct rcb_pyramid_layout {
  int32_t levels;
  int32_t dtype;                       /* enum rcb_dtype */
  int32_t tile_w;                      /* columns per tile (tile = 4 rows x tile_w columns = 64 bytes) */
  int32_t reserved;
  int32_t H[RCB_MAX_LEVELS];           /* logical rows of each level */
  int32_t W[RCB_MAX_LEVELS];           /* logical columns */
  int32_t tiles_x[RCB_MAX_LEVELS];     /* tiles per tile-row: ceil(W / tile_w) */
  int32_t tiles_y[RCB_MAX_LEVELS];     /* tile rows: ceil(H / 4) */
  int64_t plane_stride[RCB_MAX_LEVELS];/* elements between consecutive queries = tiles_x * tiles_y * 4 * tile_w */
  int64_t level_bytes[RCB_MAX_LEVELS]; /* bytes to allocate for the level: B*H*W planes */
} rcb_pyramid_layout;

RCB_API int rcb_abi_version(void);
RCB_API const char* rcb_status_string(int status);

/* Fills `layout` for a [B, C, H, W] feature pair.  Pure host arithmetic. */
RCB_API int rcb_pyramid_layout_query(int B, int H, int W, int levels, int dtype, rcb_pyramid_layout* layout);

/* ---- K1: all-pairs volume + pooled pyramid ------------------------------------------------
 * Replaces CorrBlock.corr (core/corr.py:96-127: view, matmul, / sqrt(C)) and the pyramid loop of
 * CorrBlock.__init__ (core/corr.py:44-54: reshape, 3 x avg_pool2d(2, stride=2)).
 *   fmap1, fmap2 : [B, C, H, W] fp32 (core/raft.py:181-182)
 *   pyr[l]       : level-l buffer laid out as rcb_pyramid_layout says, l < levels
 *   level l > 0 equals the 2x2 floor-mode mean of level l-1.
 * `workspace` holds the packed operands of the tensor-core modes (unused for FP32_SIMT), 128-byte aligned. */
RCB_API size_t rcb_corr_build_workspace_bytes(int B, int C, int H, int W, int mode);
RCB_API int rcb_corr_build(const float* fmap1, const float* fmap2, void* const* pyr, int B, int C, int H, int W,
                   int levels, int mode, int pyr_dtype, void* workspace, size_t workspace_bytes,
                   rcb_stream_t stream);

/* ---- next row: the build split at its operand boundary ---------------------------------
 * rcb_corr_build(fmap1, fmap2, ...) == rcb_corr_pack_fmaps + rcb_corr_build_packed (tensor-core modes only).
 * rcb_corr_pack_fmaps consumes the feature encoder's output for the concatenated frame pair as ONE tensor
 *   fmaps [2B, C, H, W] fp32 -- what fnet returns before torch.split (core/extractor.py:184-190; core/raft.py:178-182
 *   then calls .float() on the two halves): the first B maps are the queries (fmap1), the last B the targets (fmap2)
 *   -- and writes the tensor-core operands of `mode` into `packed` (rcb_corr_build_workspace_bytes(B, C, H, W, mode)
 *   bytes, 128-byte aligned): per query pixel the image of its tensor-memory lane, per target patch ready-made
 *   SWIZZLE_128B tile images.  A producer that emits this layout itself (a fused epilogue of the encoder's last
 *   1x1 convolution, core/extractor.py:144) can skip the call; the layout is described in csrc/corr_build_tc.cu.
 * rcb_corr_build_packed runs the volume + pyramid kernel on such operands (same B, C, H, W, mode). */
RCB_API int rcb_corr_pack_fmaps(const float* fmaps, void* packed, size_t packed_bytes, int B, int C, int H, int W,
                                int mode, rcb_stream_t stream);
RCB_API int rcb_corr_build_packed(const void* packed, size_t packed_bytes, void* const* pyr, int B, int C, int H, int W,
                                  int levels, int mode, int pyr_dtype, rcb_stream_t stream);

/* ---- K2: fused multi-level window lookup ------------------------------------------------
 * Replaces CorrBlock.__call__ (core/corr.py:56-94) including bilinear_sampler
 * (core/utils/utils.py:57-71 -> F.grid_sample, bilinear, zeros padding, align_corners=True), the
 * level concatenation and the permute(0,3,1,2).contiguous().
 *   coords : [B, 2, H, W], channel 0 = x, 1 = y (core/utils/utils.py:74-77)
 *   out    : [B, levels*(2r+1)^2, H, W]; channel = l*(2r+1)^2 + a*(2r+1) + b samples level l at
 *            (x/2^l + a - r, y/2^l + b - r)  -- the x offset is the slow index (core/corr.py:77-84). */
RCB_API int rcb_corr_lookup(const void* const* pyr, const float* coords, float* out, int B, int H, int W,
                    int levels, int radius, int pyr_dtype, rcb_stream_t stream);

/* Planned form: the lookup gathers every window with one TMA box, which needs tensor maps of the pyramid
 * levels.  rcb_corr_lookup encodes them on every call; a caller that looks the same pyramid up many times
 * (core/raft.py:214-219, once per GRU iteration) encodes them once into a host-side plan instead.
 *   plan : HOST memory owned by the caller, >= rcb_corr_lookup_plan_bytes() bytes, 64-byte aligned; plain data,
 *          may be copied/freed at will; valid while `pyr` buffers and geometry are unchanged. */
RCB_API size_t rcb_corr_lookup_plan_bytes(void);
RCB_API int rcb_corr_lookup_plan_init(void* plan, size_t plan_bytes, const void* const* pyr, int B, int H, int W,
                              int levels, int radius, int pyr_dtype);
RCB_API int rcb_corr_lookup_planned(const void* plan, const float* coords, float* out, rcb_stream_t stream);
/* Tuning knob of the fp32 lookup kernel: how many lanes share one query's window.  0 (what plan_init sets) lets the
 * launch choose -- 2 for grids of several waves (fewest instructions: the kernel is DRAM-bound and the board runs at
 * its power cap), 4 for grids that do not fill the GPU twice (latency-bound) --; 2 or 4 pins it.  Results are
 * identical bit for bit. */
RCB_API int rcb_corr_lookup_plan_set_lanes(void* plan, int lanes_per_query);

/* ---- K4: backward of the all-pairs path ---------------------------------------------------
 * Replaces what autograd records through core/corr.py:25-127 for train.py:212.
 * rcb_corr_lookup_backward: transposes the bilinear lookup of ONE call: scatter-adds grad_out into the
 *   dense fp32 pyramid gradient dpyr[l] (same layout as an RCB_F32 pyramid; accumulated, so several
 *   GRU iterations may share one buffer) and writes d out / d coords.  Either of `dpyr` / `dcoords`
 *   may be NULL to skip that half.
 * rcb_corr_pool_backward: folds dpyr[l] into dpyr[l-1] for l = levels-1 .. 1 (avg_pool2d backward).
 * rcb_corr_contract_backward: dF1[b,c,q] = sum_p dV0[b,q,p] F2[b,c,p] / sqrt(C),
 *                             dF2[b,c,p] = sum_q dV0[b,q,p] F1[b,c,q] / sqrt(C)   (bmm backward). */
RCB_API int rcb_corr_lookup_backward(const void* const* pyr, const float* coords, const float* grad_out,
                             float* const* dpyr, float* dcoords, int B, int H, int W, int levels,
                             int radius, int pyr_dtype, rcb_stream_t stream);
RCB_API int rcb_corr_pool_backward(float* const* dpyr, int B, int H, int W, int levels, rcb_stream_t stream);
RCB_API int rcb_corr_contract_backward(const float* fmap1, const float* fmap2, const float* dvol0, float* dfmap1,
                               float* dfmap2, int B, int C, int H, int W, rcb_stream_t stream);
/* The same two GEMMs on the tensor cores (tcgen05, hi/lo bf16 split with fp32 accumulation, ~5e-6 of max-abs;
 * C <= 256).  `workspace` (>= rcb_corr_contract_backward_tc_workspace_bytes(), 256-byte aligned) holds the packed
 * bf16 operands: dV0 and its transpose, fmap1 and fmap2. */
RCB_API size_t rcb_corr_contract_backward_tc_workspace_bytes(int B, int C, int H, int W);
RCB_API int rcb_corr_contract_backward_tc(const float* fmap1, const float* fmap2, const float* dvol0, float* dfmap1,
                                  float* dfmap2, int B, int C, int H, int W, void* workspace,
                                  size_t workspace_bytes, rcb_stream_t stream);

/* ---- K3 / K5: on-the-fly correlation (the alt_cuda_corr extension) ------------------------
 * rcb_altcorr_forward replaces alt_cuda_corr.forward (correlation.cpp:23-33,
 * correlation_kernel.cu:18-119,260-286):
 *   fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] (x,y) -> corr [B,N,(2r+1)^2,H1,W1],
 *   channel = iy + (2r+1)*ix, NOT divided by sqrt(C).  `corr` need not be zeroed by the caller.
 * rcb_altcorr_backward replaces alt_cuda_corr.backward (correlation.cpp:36-48,
 * correlation_kernel.cu:122-256,288-324): fmap1_grad [B,H1,W1,C], fmap2_grad [B,H2,W2,C],
 *   coords_grad [B,N,H1,W1,2].  The three gradient buffers are fully overwritten.
 *   true_coords_grad == 0 reproduces the reference (coords_grad is all zeros, correlation_kernel.cu:307);
 *   != 0 writes the real derivative, as autograd through CorrBlock does. */
RCB_API int rcb_altcorr_forward(const float* fmap1, const float* fmap2, const float* coords, float* corr, int B, int N,
                        int H1, int W1, int H2, int W2, int C, int radius, rcb_stream_t stream);
RCB_API int rcb_altcorr_backward(const float* fmap1, const float* fmap2, const float* coords, const float* corr_grad,
                         float* fmap1_grad, float* fmap2_grad, float* coords_grad, int B, int N, int H1, int W1,
                         int H2, int W2, int C, int radius, int true_coords_grad, rcb_stream_t stream);

/* Fused form used by AlternateCorrBlock (core/corr.py:140-198): one launch per GRU iteration for all
 * levels instead of 4 launches + 8 permute copies + stack + divide.
 * rcb_altcorr_prepare: NCHW fmap1 -> NHWC copy; NCHW fmap2 -> NHWC pooled feature pyramid
 *   (core/corr.py:157-161,183-184), once per frame pair.
 *     fmap1_nhwc [B,H,W,C]; fmap2_nhwc[l] [B,H_l,W_l,C] with H_l, W_l floor-halved.
 * rcb_altcorr_pyramid_forward: out [B, levels*(2r+1)^2, H, W] = stacked per-level on-the-fly
 *   correlation at coords/2^l, times `scale` (1/sqrt(C), core/corr.py:198). */
RCB_API int rcb_altcorr_prepare(const float* fmap1, const float* fmap2, float* fmap1_nhwc, float* const* fmap2_nhwc,
                        int B, int C, int H, int W, int levels, rcb_stream_t stream);
RCB_API int rcb_altcorr_pyramid_forward(const float* fmap1_nhwc, const float* const* fmap2_nhwc, const float* coords,
                                float* out, int B, int C, int H, int W, int levels, int radius, float scale,
                                rcb_stream_t stream);

/* ---- next row: convex upsampling of the flow -------------------------------
 * Replaces RAFT.upsample_flow (core/raft.py:112-142: view, softmax over the 9 neighbours, F.unfold(8*flow, 3x3,
 * padding=1), weighted sum, permute, reshape), called once per GRU iteration (core/raft.py:240).
 *   flow [N, 2, H, W], mask [N, 576, H, W] (channel = k*64 + i*8 + j)  ->  out [N, 2, 8H, 8W]
 * Backward: d flow [N, 2, H, W] and d mask [N, 576, H, W] from d out; `workspace` holds N*2*9*H*W floats. */
RCB_API int rcb_upsample_flow(const float* flow, const float* mask, float* out, int N, int H, int W, rcb_stream_t stream);
RCB_API size_t rcb_upsample_flow_backward_workspace_bytes(int N, int H, int W);
RCB_API int rcb_upsample_flow_backward(const float* flow, const float* mask, const float* grad_out, float* dflow,
                               float* dmask, void* workspace, size_t workspace_bytes, int N, int H, int W,
                               rcb_stream_t stream);

/* ---- lookup fused with the motion encoder's first layer ----------------------------------
 * Replaces the pair  corr = corr_fn(coords1)  (core/raft.py:219)  ->  cor = F.relu(self.convc1(corr))
 * (core/update.py:154 SmallMotionEncoder, :202 BasicMotionEncoder; convc1 = Conv2d(levels*(2r+1)^2, cout, 1),
 * core/update.py:136,182): out[b, n, y, x] = act(sum_k corr[b, k, y, x] * weight[n, k] + bias[n]) without ever
 * materialising corr.  fp16 operands on the tensor cores, fp32 accumulate (same class as the TF32 convolution the
 * reference gets from cuDNN by default): within 1e-3 of max-abs of the fp32 result.  fp32 pyramids, radius 3 or 4,
 * cout a multiple of 16 in 16..256; anything else returns RCB_ERR_UNSUPPORTED (0 bytes from the size query).
 *   rcb_corr_convc1_pack   weight [cout][levels*(2r+1)^2] fp32 (the Conv2d weight, contiguous) -> the kernel's
 *                          K-permuted, zero-padded fp16 operand; once per set of weights.  wpack: 128-byte aligned,
 *                          rcb_corr_convc1_pack_bytes() bytes.
 *   rcb_corr_lookup_convc1 plan: a lookup plan of the pyramid (rcb_corr_lookup_plan_init); bias may be NULL;
 *                          relu != 0 applies max(x, 0); out [B][cout][H][W] fp32. */
RCB_API size_t rcb_corr_convc1_pack_bytes(int cout, int levels, int radius);
RCB_API int rcb_corr_convc1_pack(const float* weight, void* wpack, int cout, int levels, int radius, rcb_stream_t stream);
RCB_API int rcb_corr_lookup_convc1(const void* plan, const float* coords, const void* wpack, const float* bias,
                                   float* out, int cout, int relu, rcb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RAFT_CORR_B200_H_ */
