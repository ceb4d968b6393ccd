// K2: fused multi-level correlation-window lookup (forward) and its transpose (backward).
//
// Replaces CorrBlock.__call__ (reference core/corr.py:56-94) and bilinear_sampler
// (core/utils/utils.py:57-71): per GRU iteration, for every query pixel and pyramid level a
// (2r+2)^2 tap window is gathered from that query's correlation plane and bilinearly resampled to the
// (2r+1)^2 outputs the update block consumes.  HBM-bound gather: see DESIGN.md "K2".
//
// The window of one (query, level) is exactly a box of whole 64-byte tiles of the pyramid layout:
// ny tile rows x nx tiles, nx/ny = 3 or 4 at r = 4 depending on the window's phase inside the tile grid.  One
// cp.async.bulk.tensor (TMA) per query fetches it -- 3..4 contiguous runs of 192..256 bytes -- straight into
// shared memory: address generation, the zeros padding outside the plane (TMA out-of-bounds fill) and the data
// movement cost no registers, no LSU instructions and no L1 data-pipe wavefronts (an earlier register-staged
// version of this kernel was bound by exactly those: ncu L1 data pipe 72 % busy, DRAM 43 %; this one runs at
// ~80 % of the measured copy bandwidth for the bytes it really moves).  Four tensor maps per level (the nx x ny
// combinations) live in a host-side plan that is encoded once per pyramid, not per call.
//   CTA   = 32 consecutive queries of ONE level, 2 lanes per query (64 threads), a warp owns 16 queries end to
//           end: the first lane of a query issues its box on the warp's own mbarrier, the warp waits, then does
//           the math; eight CTAs per SM keep ~100 KB of gathers in flight.
//   math  = the lanes of a query split the (2r+1) y offsets; a lane reads its 5-6 window rows as whole tile
//           rows (LDS.128), aligns them to the window's 4-byte phase with two select stages, applies the separable
//           bilinear weights horizontally, then vertically against the previous row (all (2r+1)^2 samples of a
//           level share one fractional offset because the window offsets are integers).
//   store = one instruction writes 2 channels x 16 consecutive queries (64-byte runs).
// Padding INSIDE edge tiles (rows >= H_l, columns >= W_l of the last tile row/column) is unspecified in the
// layout, so those taps are masked here; everything outside the tile grid is zero-filled by the TMA unit.
#include <cstdlib>

#include "lookup_common.cuh"
#include "rcb_common.cuh"
#include "tma_util.cuh"

namespace rcb {

// Programmatic dependent launch (see tma_util.cuh): only the lookup kernels of this file execute
// launch_dependents, and they write nothing but their own output tensor, so a following lookup can safely read the
// pyramid and its coordinates early; it executes pdl_wait() before its first store, so a stream-ordered allocator
// that hands the same output memory to two consecutive calls stays correct.  Any other preceding kernel never
// triggers, which leaves the ordinary stream order.

// r = 4 only (XROW): a window needs a 4th tile row exactly when it starts in the last pixel row of a tile (py = 3), and
// then it uses a single pixel row of it.  The TMA box is therefore always 3 tile rows (12 tiles, 768-byte slots instead
// of 1024) and that one pixel row -- 16 bytes per tile -- is fetched with cp.async into 64 bytes per query.  The
// smaller slots let 8 CTAs share an SM instead of 6 (36 -> 33 us at cfg2) and the window touches fewer sectors.
//
// Lanes per query of the fp32 lookup kernel (they split the (2r+1) window rows): 2 for grids of several waves, 4 for
// small ones (rcb_corr_lookup_plan_set_lanes pins it).  Two lanes halve the per-query prologue (coordinates, window phase, TMA issue) every lane executes: 26 %
// fewer instructions per launch.  In a short loop the kernel is DRAM-bound either way (33.2 us at cfg2); under
// sustained load the board runs at its power cap and the leaner kernel leaves the SM clock higher: whole step
// 1914 -> 1866 us on the same GPU.  A grid that does not fill the GPU twice is latency-bound
// instead (cfg1: one frame pair, 880 CTAs): there four lanes finish a query sooner (6.9 against 8.9 us per launch).
// RCB_LOOKUP_LPQ = 2 or 4 pins the choice (A/B builds).
#ifndef RCB_LOOKUP_LPQ
#define RCB_LOOKUP_LPQ 0
#endif
constexpr int kLookupLPQ = RCB_LOOKUP_LPQ;
static_assert(kLookupLPQ == 0 || kLookupLPQ == 2 || kLookupLPQ == 4, "lanes per query");

template <int R, int LPQ>
__global__ void __launch_bounds__(TmaCfg<R>::QT * LPQ, 8)
lookup_tma_kernel(const __grid_constant__ LookupMaps maps, const __grid_constant__ PyramidDev pyr,
                  const float* __restrict__ coords, float* __restrict__ out, int Q, int L, int dbg) {
  using Cfg = TmaCfg<R>;
  constexpr int RD = Cfg::RD, ROWS = Cfg::ROWS, NMIN = Cfg::NMIN, NMAX = Cfg::NMAX;
  constexpr int QT = Cfg::QT, THREADS = QT * LPQ;
  constexpr int NBMAX = (RD + LPQ - 1) / LPQ;    // output rows per lane
  constexpr bool XROW = R == 4;
  constexpr int NYBOX = XROW ? NMAX - 1 : NMAX;  // tile rows a box can have
  constexpr int SLOT16 = XROW ? (NMAX * NYBOX * 64 + 127) / 128 * 8 : Cfg::SLOT16;
  __shared__ __align__(128) float4 slots[QT * SLOT16];
  __shared__ __align__(16) float4 xrow[XROW ? QT * 4 : 1];  // [query][tile]: pixel row 0 of the 4th tile row
  __shared__ __align__(8) unsigned long long bars[THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = blockIdx.y;
  const int b = blockIdx.z;
  const int ql = tid / LPQ, sub = tid % LPQ;
  const int q = blockIdx.x * QT + ql;
  const bool q_ok = q < Q;
  // (pyr is a __grid_constant__ parameter: indexing it by the level is an indexed constant-bank load, not a local copy)
  const int Hl = pyr.H[l], Wl = pyr.W[l];
  float cx = -1.0e6f, cy = -1.0e6f;
  if (q_ok) {
    cx = __ldg(coords + (long long)(b * 2 + 0) * Q + q);
    cy = __ldg(coords + (long long)(b * 2 + 1) * Q + q);
  }
  const LevelCoord lc = level_coord<R>(cx, cy, l, Hl, Wl);
  const int ph = lc.xs & 3, py = lc.ys & 3;
  const int nx = (ph + ROWS + 3) >> 2, ny = (py + ROWS + 3) >> 2;
  const int nyb = ny < NYBOX ? ny : NYBOX;  // tile rows of the TMA box

  // Programmatic dependent launch: consecutive lookups of one GRU loop are independent (each writes its own
  // output), so the next launch may begin its gathers while this one drains; see pdl_wait() below.
  pdl_launch_dependents();
  const uint32_t bar = smem_u32(&bars[warp]);
  if (lane == 0) {
    mbar_init(bar, 32 / LPQ);
    fence_barrier_init();
  }
  __syncwarp();
  if (sub == 0) {
    if (q_ok && !(dbg & 2)) {
      mbar_expect_tx(bar, (uint32_t)(nx * nyb * 64));
      tma_load_3d(smem_u32(slots + ql * SLOT16), &maps.m[l * 4 + (nyb - NMIN) * 2 + (nx - NMIN)], bar,
                  (lc.xs >> 2) * 16, lc.ys >> 2, b * Q + q);
    } else {
      mbar_arrive(bar);
    }
  }
  if (XROW) {
    if (q_ok && ny > NYBOX && sub < nx) {  // the lanes of the query fetch the row pieces of its nx tiles
      const float* pl = static_cast<const float*>(pyr.ptr[l]);
      const int txs = pyr.tiles_x[l];
      const long long ps = pyr.plane_stride[l];
      const int ty = (lc.ys >> 2) + NYBOX;
      const float* plane = pl + ((long long)b * Q + q) * ps;
#pragma unroll
      for (int t = sub; t < NMAX; t += LPQ) {
        if (t < nx) {
          const int tx = (lc.xs >> 2) + t;
          const bool in = ty >= 0 && ty * 4 < Hl && tx >= 0 && tx < txs;  // tiles outside the grid read as zeros
          cp_async16_zfill(smem_u32(xrow + ql * 4 + t), plane + (in ? ((long long)(ty * txs + tx) << 4) : 0), in ? 16 : 0);
        }
      }
    }
    cp_async_commit();
  }
  if (!q_ok) return;
  const float fx = lc.fx, fy = lc.fy, gx = 1.0f - lc.fx, gy = 1.0f - lc.fy;
  const int b0 = (RD * sub) / LPQ, nb = (RD * (sub + 1)) / LPQ - b0;  // output rows [b0, b0 + nb), nb <= NBMAX
  float* o = out + (((long long)b * L + l) * RD * RD + b0) * Q + q;   // channel = a * RD + b
  int offa[RD];  // channel offset of x offset a (one level's window of one pair is < 2^31 elements, see plan_init)
#pragma unroll
  for (int a = 0; a < RD; ++a) offa[a] = a * RD * Q;
  const float4* slot = slots + ql * SLOT16;
  const bool ragged_w = (Wl & 3) != 0;
  mbar_wait(bar, 0);
  if (XROW) {
    cp_async_wait<0>();
    __syncwarp();  // the row pieces were fetched by the other lanes of the query
  }
  if (dbg & 1) return;  // timing experiment: gather only

  float hp[RD];
#pragma unroll
  for (int jj = 0; jj <= NBMAX; ++jj) {
    if (jj > nb) break;
    const int j = b0 + jj;     // window row
    const int ya = py + j;     // row inside the fetched box
    const bool row_ok = lc.ys + j < Hl;  // rows < 0 lie in tile rows the TMA zero-filled
    const bool ext = XROW && (ya >> 2) >= NYBOX;  // the single pixel row beyond the box
    const float4* rowp = ext ? xrow + ql * 4 : slot + ((ya >> 2) * nx) * 4 + (ya & 3);
    const int ks = ext ? 1 : 4;
    float w[4 * NMAX];
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok && k < nx) u = rowp[k * ks];
      w[4 * k + 0] = u.x; w[4 * k + 1] = u.y; w[4 * k + 2] = u.z; w[4 * k + 3] = u.w;
    }
    float v1[ROWS + 2];
#pragma unroll
    for (int i = 0; i < ROWS + 2; ++i) v1[i] = (ph & 1) ? w[i + 1] : w[i];
    float t[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) t[i] = (ph & 2) ? v1[i + 2] : v1[i];
    if (ragged_w) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i)
        if (lc.xs + i >= Wl) t[i] = 0.f;
    }
    float h[RD];
#pragma unroll
    for (int a = 0; a < RD; ++a) h[a] = gx * t[a] + fx * t[a + 1];
    if (jj > 0) {
      if (jj == 1) pdl_wait();  // nothing is written before the preceding kernel of the stream has completed
#pragma unroll
      for (int a = 0; a < RD; ++a) o[offa[a]] = gy * hp[a] + fy * h[a];
      o += Q;
    }
#pragma unroll
    for (int a = 0; a < RD; ++a) hp[a] = h[a];
  }
}

// fp16 pyramid (RCB_F16): tiles are 4 rows x 8 halfs, so a window spans 2-3 tiles in x (3-4 in y) and a tile row
// (16 bytes) holds 8 taps; the taps are widened to fp32 on load and everything after that is the fp32 kernel with
// one more select stage (the phase inside a tile row is 0..7).
template <int R>
struct TmaCfgH {
  static constexpr int RD = 2 * R + 1;
  static constexpr int ROWS = 2 * R + 2;
  static constexpr int NMINX = (ROWS + 7) >> 3, NMAXX = (ROWS + 14) >> 3;
  static constexpr int NMINY = (ROWS + 3) >> 2, NMAXY = (ROWS + 6) >> 2;
  static constexpr int SLOT_BYTES = (NMAXX * NMAXY * 64 + 127) / 128 * 128;
  static constexpr int SLOT16 = SLOT_BYTES / 16;
  static constexpr int NBMAX = (RD + 3) / 4;
  static constexpr int QT = 32, THREADS = 128;
};

template <int R, int LPQ>
__global__ void __launch_bounds__(TmaCfgH<R>::QT * LPQ, 8)
lookup_tma_f16_kernel(const __grid_constant__ LookupMaps maps, const __grid_constant__ PyramidDev pyr,
                      const float* __restrict__ coords, float* __restrict__ out, int Q, int L) {
  using Cfg = TmaCfgH<R>;
  constexpr int RD = Cfg::RD, ROWS = Cfg::ROWS, NMINX = Cfg::NMINX, NMAXX = Cfg::NMAXX, NMINY = Cfg::NMINY;
  constexpr int SLOT16 = Cfg::SLOT16, QT = Cfg::QT, THREADS = QT * LPQ;
  constexpr int NBMAX = (RD + LPQ - 1) / LPQ;  // output rows per lane
  __shared__ __align__(128) uint4 slots[QT * SLOT16];
  __shared__ __align__(8) unsigned long long bars[THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = blockIdx.y;
  const int b = blockIdx.z;
  const int ql = tid / LPQ, sub = tid % LPQ;
  const int q = blockIdx.x * QT + ql;
  const bool q_ok = q < Q;
  const int Hl = pyr.H[l], Wl = pyr.W[l];
  float cx = -1.0e6f, cy = -1.0e6f;
  if (q_ok) {
    cx = __ldg(coords + (long long)(b * 2 + 0) * Q + q);
    cy = __ldg(coords + (long long)(b * 2 + 1) * Q + q);
  }
  const LevelCoord lc = level_coord<R>(cx, cy, l, Hl, Wl);
  const int ph = lc.xs & 7, py = lc.ys & 3;
  const int nx = (ph + ROWS + 7) >> 3, ny = (py + ROWS + 3) >> 2;

  pdl_launch_dependents();
  const uint32_t bar = smem_u32(&bars[warp]);
  if (lane == 0) {
    mbar_init(bar, 32 / LPQ);
    fence_barrier_init();
  }
  __syncwarp();
  if (sub == 0) {
    if (q_ok) {
      mbar_expect_tx(bar, (uint32_t)(nx * ny * 64));
      tma_load_3d(smem_u32(slots + ql * SLOT16), &maps.m[l * 4 + (ny - NMINY) * 2 + (nx - NMINX)], bar,
                  (lc.xs >> 3) * 16, lc.ys >> 2, b * Q + q);
    } else {
      mbar_arrive(bar);
    }
  }
  if (!q_ok) return;
  const float fx = lc.fx, fy = lc.fy, gx = 1.0f - lc.fx, gy = 1.0f - lc.fy;
  const int b0 = (RD * sub) / LPQ, nb = (RD * (sub + 1)) / LPQ - b0;  // output rows [b0, b0 + nb), nb <= NBMAX
  float* o = out + (((long long)b * L + l) * RD * RD + b0) * Q + q;   // channel = a * RD + b
  int offa[RD];  // channel offset of x offset a (< 2^31 elements, see plan_init)
#pragma unroll
  for (int a = 0; a < RD; ++a) offa[a] = a * RD * Q;
  const uint4* slot = slots + ql * SLOT16;
  const bool ragged_w = (Wl & 7) != 0;
  constexpr int NP = (ROWS / 2 + 3) / 2;                 // 8-byte pairs that cover ROWS / 2 + 2 words
  const int pair_step = ph >> 2;                         // whole pairs the window starts into the row
  const int off_even = pair_step * 8, off_odd = pair_step * 56;
  const bool word_step = (ph >> 1) & 1;
  const uint32_t half_sel = (ph & 1) ? 0x5432u : 0x3210u;
  mbar_wait(bar, 0);

  float hp[RD];
#pragma unroll
  for (int jj = 0; jj <= NBMAX; ++jj) {
    if (jj > nb) break;
    const int j = b0 + jj;     // window row
    const int ya = py + j;     // row inside the fetched box
    const bool row_ok = lc.ys + j < Hl;  // rows < 0 lie in tile rows the TMA zero-filled
    // The ROWS taps start ph halfs (0..7) into the row.  Whole 8-byte pairs of that offset are folded into the load
    // addresses (a tile row is 16 bytes, the next tile 64 bytes further), the remaining 4-byte step is one select
    // stage on packed words and the odd half is a byte permute; only the ROWS taps are then widened to fp32.
    const unsigned char* rb = reinterpret_cast<const unsigned char*>(slot + ((ya >> 2) * nx) * 4 + (ya & 3));
    uint32_t x[2 * NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const uint2 pr = *reinterpret_cast<const uint2*>(rb + (k >> 1) * 64 + ((k & 1) ? 8 + off_odd : off_even));
      x[2 * k] = pr.x;
      x[2 * k + 1] = pr.y;
    }
    float t[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS / 2; ++i) {
      const uint32_t y0 = word_step ? x[i + 1] : x[i], y1 = word_step ? x[i + 2] : x[i + 1];
      uint32_t z = __byte_perm(y0, y1, half_sel);
      if (!row_ok) z = 0u;
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&z));
      t[2 * i] = f.x;
      t[2 * i + 1] = f.y;
    }
    if (ragged_w) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i)
        if (lc.xs + i >= Wl) t[i] = 0.f;
    }
    float h[RD];
#pragma unroll
    for (int a = 0; a < RD; ++a) h[a] = gx * t[a] + fx * t[a + 1];
    if (jj > 0) {
      if (jj == 1) pdl_wait();  // nothing is written before the preceding kernel of the stream has completed
#pragma unroll
      for (int a = 0; a < RD; ++a) o[offa[a]] = gy * hp[a] + fy * h[a];
      o += Q;
    }
#pragma unroll
    for (int a = 0; a < RD; ++a) hp[a] = h[a];
  }
}

static int plan_init(LookupPlan* plan, const void* const* pyr, const rcb_pyramid_layout& lay, int B, int H, int W,
                     int radius) {
  if (!encode_fn()) return RCB_ERR_NO_DEVICE;
  const bool f16 = lay.dtype == RCB_F16;
  const int esize = f16 ? 2 : 4;
  const int rows = 2 * radius + 2;
  const int nminx = f16 ? (rows + 7) >> 3 : (rows + 3) >> 2, nminy = (rows + 3) >> 2;
  const long long planes = (long long)B * H * W;
  // the kernels index the window channels of one pair and level with 32-bit offsets
  if ((long long)H * W * (2 * radius + 1) * (2 * radius + 1) > 0x7fffffffLL) return RCB_ERR_UNSUPPORTED;
  for (int l = 0; l < lay.levels; ++l) {
    for (int sel = 0; sel < 4; ++sel) {
      const int nx = nminx + (sel & 1), ny = nminy + (sel >> 1);
      // in 4-byte words: a 64-byte tile is 16 words for both element types
      cuuint64_t dims[3] = {(cuuint64_t)lay.tiles_x[l] * 16, (cuuint64_t)lay.tiles_y[l], (cuuint64_t)planes};
      cuuint64_t str[2] = {(cuuint64_t)lay.tiles_x[l] * 64, (cuuint64_t)lay.plane_stride[l] * esize};
      cuuint32_t box[3] = {(cuuint32_t)nx * 16, (cuuint32_t)ny, 1};
      if (!encode(&plan->maps.m[l * 4 + sel], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, pyr[l], dims, str, box,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE))
        return RCB_ERR_INVALID_ARGUMENT;
    }
    plan->ptr[l] = pyr[l];
  }
  for (int l = lay.levels; l < RCB_MAX_LEVELS; ++l) {
    plan->ptr[l] = nullptr;
    for (int sel = 0; sel < 4; ++sel) plan->maps.m[l * 4 + sel] = plan->maps.m[0];
  }
  plan->lay = lay;
  plan->B = B; plan->H = H; plan->W = W; plan->radius = radius;
  plan->lanes = 0;
  plan->magic = kPlanMagic;
  return RCB_OK;
}

static cudaLaunchConfig_t pdl_config(dim3 grid, int threads, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  return cfg;
}
// RCB_LOOKUP_PDL=0 (RCB_DEBUG builds only) launches with ordinary stream serialization (A/B timing)
static int pdl_attribute(cudaLaunchAttribute* attr) {
  static const bool on = debug_env_int("RCB_LOOKUP_PDL", 1) != 0;
  if (!on) return 0;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  return 1;
}

template <int R>
static int launch_lookup_tma_r(const LookupPlan& plan, const PyramidDev& pd, const float* coords, float* out,
                               cudaStream_t s) {
  using Cfg = TmaCfg<R>;
  const int Q = plan.H * plan.W;
  dim3 grid((Q + Cfg::QT - 1) / Cfg::QT, plan.lay.levels, plan.B);
  static const int dbg = debug_env_int("RCB_LOOKUP_DEBUG", 0);  // RCB_DEBUG builds only
  // two waves of 8 CTAs per SM on 148 SMs
  const bool small = (long long)grid.x * grid.y * grid.z <= 2LL * 148 * 8;
  const int lpq = kLookupLPQ ? kLookupLPQ : plan.lanes ? plan.lanes : (small ? 4 : 2);
  cudaLaunchConfig_t cfg = pdl_config(grid, Cfg::QT * lpq, s);
  cudaLaunchAttribute attr[1];
  cfg.numAttrs = pdl_attribute(attr);
  cfg.attrs = attr;
  cudaError_t e = lpq == 4 ? cudaLaunchKernelEx(&cfg, lookup_tma_kernel<R, 4>, plan.maps, pd, coords, out, Q,
                                                (int)plan.lay.levels, dbg)
                           : cudaLaunchKernelEx(&cfg, lookup_tma_kernel<R, 2>, plan.maps, pd, coords, out, Q,
                                                (int)plan.lay.levels, dbg);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

template <int R>
static int launch_lookup_tma_f16_r(const LookupPlan& plan, const PyramidDev& pd, const float* coords, float* out,
                                   cudaStream_t s) {
  using Cfg = TmaCfgH<R>;
  const int Q = plan.H * plan.W;
  dim3 grid((Q + Cfg::QT - 1) / Cfg::QT, plan.lay.levels, plan.B);
  const bool small = (long long)grid.x * grid.y * grid.z <= 2LL * 148 * 8;
  const int lpq = kLookupLPQ ? kLookupLPQ : plan.lanes ? plan.lanes : (small ? 4 : 2);
  cudaLaunchConfig_t cfg = pdl_config(grid, Cfg::QT * lpq, s);
  cudaLaunchAttribute attr[1];
  cfg.numAttrs = pdl_attribute(attr);
  cfg.attrs = attr;
  cudaError_t e = lpq == 4 ? cudaLaunchKernelEx(&cfg, lookup_tma_f16_kernel<R, 4>, plan.maps, pd, coords, out, Q,
                                                (int)plan.lay.levels)
                           : cudaLaunchKernelEx(&cfg, lookup_tma_f16_kernel<R, 2>, plan.maps, pd, coords, out, Q,
                                                (int)plan.lay.levels);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

size_t lookup_plan_bytes() { return sizeof(LookupPlan); }

int lookup_plan_init(void* plan, size_t plan_bytes, const void* const* pyr, const rcb_pyramid_layout& lay, int B,
                     int H, int W, int radius) {
  if (!plan || plan_bytes < sizeof(LookupPlan) || (reinterpret_cast<uintptr_t>(plan) & 63))
    return RCB_ERR_INVALID_ARGUMENT;
  return plan_init(static_cast<LookupPlan*>(plan), pyr, lay, B, H, W, radius);
}

int lookup_plan_set_lanes(void* plan_, int lanes) {
  LookupPlan* plan = static_cast<LookupPlan*>(plan_);
  if (!plan || (reinterpret_cast<uintptr_t>(plan_) & 63) || plan->magic != kPlanMagic) return RCB_ERR_INVALID_ARGUMENT;
  if (lanes != 0 && lanes != 2 && lanes != 4) return RCB_ERR_INVALID_ARGUMENT;
  plan->lanes = lanes;
  return RCB_OK;
}

int launch_lookup_planned(const void* plan_, const float* coords, float* out, cudaStream_t s) {
  const LookupPlan* plan = static_cast<const LookupPlan*>(plan_);
  if (!plan || (reinterpret_cast<uintptr_t>(plan_) & 63) || plan->magic != kPlanMagic) return RCB_ERR_INVALID_ARGUMENT;
  const PyramidDev pd = make_pyramid_dev(plan->ptr, plan->lay);
  if (plan->lay.dtype == RCB_F16) {
    switch (plan->radius) {
      case 1: return launch_lookup_tma_f16_r<1>(*plan, pd, coords, out, s);
      case 2: return launch_lookup_tma_f16_r<2>(*plan, pd, coords, out, s);
      case 3: return launch_lookup_tma_f16_r<3>(*plan, pd, coords, out, s);
      case 4: return launch_lookup_tma_f16_r<4>(*plan, pd, coords, out, s);
      default: return RCB_ERR_UNSUPPORTED;
    }
  }
  switch (plan->radius) {
    case 1: return launch_lookup_tma_r<1>(*plan, pd, coords, out, s);
    case 2: return launch_lookup_tma_r<2>(*plan, pd, coords, out, s);
    case 3: return launch_lookup_tma_r<3>(*plan, pd, coords, out, s);
    case 4: return launch_lookup_tma_r<4>(*plan, pd, coords, out, s);
    default: return RCB_ERR_UNSUPPORTED;
  }
}

// Unplanned entry: encodes the tensor maps on every call (a few microseconds of host time).
int launch_lookup(const void* const* pyr, const rcb_pyramid_layout& lay, const float* coords, float* out, int B,
                  int H, int W, int radius, cudaStream_t s) {
  alignas(64) LookupPlan plan;
  const int st = plan_init(&plan, pyr, lay, B, H, W, radius);
  if (st != RCB_OK) return st;
  return launch_lookup_planned(&plan, coords, out, s);
}

// ---------------------------------------------------------------------------------------------
// Backward of one lookup call (K4, first half): transpose of the bilinear resampling.
//   dpyr[l][q, y, x] += sum over window entries that touch tap (y, x) of  grad_out * weight
//   dcoords[q]       =  sum_l 2^-l * sum_{a,b} grad_out * d(sample)/d(coordinate)
// One warp per (query, level): lanes stride over the (2r+1)^2 window entries; the four taps of an entry
// are scattered with red.global.add.f32 (planes of different queries never alias, and within a plane
// at most 4 entries hit one tap, so contention is negligible); coords gradients are warp-reduced.
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(128)
lookup_backward_kernel(PyramidDev pyr, PyramidDev dpyr, const float* __restrict__ coords,
                       const float* __restrict__ grad_out, float* __restrict__ dcoords, int B, int H, int W,
                       int L, int want_dpyr) {
  constexpr int RD = 2 * R + 1;
  const long long Q = (long long)H * W;
  const int lane = threadIdx.x & 31;
  const long long bq = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);  // one warp per query
  if (bq >= (long long)B * Q) return;
  const int b = (int)(bq / Q);
  const long long q = bq % Q;
  const float cx = __ldg(coords + ((long long)b * 2 + 0) * Q + q);
  const float cy = __ldg(coords + ((long long)b * 2 + 1) * Q + q);
  float gx = 0.0f, gy = 0.0f;
  for (int l = 0; l < L; ++l) {
    const int Hl = pyr.H[l], Wl = pyr.W[l], tw = pyr.tiles_x[l];
    const float* plane = static_cast<const float*>(pyr.ptr[l]) + bq * pyr.plane_stride[l];
    float* dplane = want_dpyr ? static_cast<float*>(const_cast<void*>(dpyr.ptr[l])) + bq * dpyr.plane_stride[l]
                              : nullptr;
    const LevelCoord lc = level_coord<R>(cx, cy, l, Hl, Wl);
    const float fx = lc.fx, fy = lc.fy;
    const float inv = 1.0f / (float)(1 << l);
    float lgx = 0.0f, lgy = 0.0f;
    for (int e = lane; e < RD * RD; e += 32) {
      const int a = e / RD, bb = e % RD;
      const float go = __ldg(grad_out + (((long long)b * L + l) * RD * RD + e) * Q + q);
      const int x0 = lc.xs + a, y0 = lc.ys + bb;
      const float wgt[4] = {(1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy};
      const float wdx[4] = {-(1 - fy), (1 - fy), -fy, fy};
      const float wdy[4] = {-(1 - fx), -fx, (1 - fx), fx};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x0 + (k & 1), y = y0 + (k >> 1);
        if (x >= 0 && x < Wl && y >= 0 && y < Hl) {
          const long long off = tile_off(y, x, tw);
          if (want_dpyr) atomicAdd(dplane + off, go * wgt[k]);
          const float v = __ldg(plane + off);
          lgx += go * wdx[k] * v;
          lgy += go * wdy[k] * v;
        }
      }
    }
    gx += lgx * inv;
    gy += lgy * inv;
  }
  gx = warp_sum(gx);
  gy = warp_sum(gy);
  if (lane == 0 && dcoords != nullptr) {
    dcoords[((long long)b * 2 + 0) * Q + q] = gx;
    dcoords[((long long)b * 2 + 1) * Q + q] = gy;
  }
}

// Pyramid-gradient half of the lookup backward, restructured like the forward kernel: CTA = 32 consecutive queries
// of ONE level, 4 lanes per query.  The resampling is separable, so is its transpose:
//   dh[j][a] = gy * G[a][j] + fy * G[a][j-1]        (G = grad_out of this level, (2r+1) x (2r+1))
//   dW[j][i] = gx * dh[j][i] + fx * dh[j][i-1]      (dW = gradient of the (2r+2)^2 tap window)
// The 4 lanes of a query split the window rows; a row is aligned to the 16-byte tile rows of the plane with two
// select stages and added with red.global.add.v4.f32 -- one 16-byte reduction per tile row instead of four scalar
// atomics per window entry (planes of different queries never alias; several GRU iterations add into one buffer).
template <int R>
__global__ void __launch_bounds__(128)
lookup_backward_dpyr_kernel(PyramidDev dpyr, const float* __restrict__ coords, const float* __restrict__ grad_out,
                            int Q, int L) {
  constexpr int RD = 2 * R + 1, ROWS = 2 * R + 2;
  constexpr int NCH = (ROWS + 6) >> 2;  // 16-byte chunks a window row can overlap
  const int tid = threadIdx.x;
  const int l = blockIdx.y, b = blockIdx.z;
  const int ql = tid >> 2, sub = tid & 3;
  const int q = blockIdx.x * 32 + ql;
  if (q >= Q) return;
  const int Hl = l == 0 ? dpyr.H[0] : l == 1 ? dpyr.H[1] : l == 2 ? dpyr.H[2] : dpyr.H[3];
  const int Wl = l == 0 ? dpyr.W[0] : l == 1 ? dpyr.W[1] : l == 2 ? dpyr.W[2] : dpyr.W[3];
  const int tw = l == 0 ? dpyr.tiles_x[0] : l == 1 ? dpyr.tiles_x[1] : l == 2 ? dpyr.tiles_x[2] : dpyr.tiles_x[3];
  const long long ps = l == 0 ? dpyr.plane_stride[0] : l == 1 ? dpyr.plane_stride[1]
                     : l == 2 ? dpyr.plane_stride[2] : dpyr.plane_stride[3];
  float* base = static_cast<float*>(const_cast<void*>(l == 0 ? dpyr.ptr[0] : l == 1 ? dpyr.ptr[1]
                                                      : l == 2 ? dpyr.ptr[2] : dpyr.ptr[3]));
  float* dplane = base + ((long long)b * Q + q) * ps;
  const float cx = __ldg(coords + (long long)(b * 2 + 0) * Q + q);
  const float cy = __ldg(coords + (long long)(b * 2 + 1) * Q + q);
  const LevelCoord lc = level_coord<R>(cx, cy, l, Hl, Wl);
  const int ph = lc.xs & 3, xa = lc.xs - ph;
  const float fx = lc.fx, fy = lc.fy, gx = 1.0f - lc.fx, gy = 1.0f - lc.fy;
  const int j0 = (ROWS * sub) >> 2, j1 = (ROWS * (sub + 1)) >> 2;  // window rows [j0, j1) of this lane
  constexpr int NJ = (ROWS + 3) / 4;
  const float* g = grad_out + (((long long)b * L + l) * RD * RD) * Q + q;  // G[a][bb] at g[(a * RD + bb) * Q]
  // grad_out rows bb = j0 - 1 .. j1 - 1 (all x offsets a): loaded once, each read is a 32-byte sector per 8 queries
  float G[NJ + 1][RD];
#pragma unroll
  for (int t = 0; t <= NJ; ++t) {
    const int bb = j0 - 1 + t;
    const bool ok = bb >= 0 && bb < RD && bb < j1;
#pragma unroll
    for (int a = 0; a < RD; ++a) G[t][a] = ok ? __ldg(g + (long long)(a * RD + bb) * Q) : 0.f;
  }
#pragma unroll
  for (int t = 0; t < NJ; ++t) {
    const int j = j0 + t;
    if (j >= j1) break;
    const int y = lc.ys + j;
    if (y < 0 || y >= Hl) continue;
    float dw[ROWS + 6];  // dW[i] at dw[3 + i], zeros around it
#pragma unroll
    for (int i = 0; i < ROWS + 6; ++i) dw[i] = 0.f;
    {
      float dh[RD];
#pragma unroll
      for (int a = 0; a < RD; ++a) dh[a] = gy * G[t + 1][a] + fy * G[t][a];  // G[t+1] = row bb = j, G[t] = row j - 1
#pragma unroll
      for (int i = 0; i < ROWS; ++i)
        dw[3 + i] = (i < RD ? gx * dh[i] : 0.f) + (i >= 1 ? fx * dh[i - 1] : 0.f);
    }
    // w[k] = dW[k - ph] for chunk-aligned position k = 0 .. 4 * NCH - 1
    float s1[4 * NCH + 1], w[4 * NCH];
#pragma unroll
    for (int k = 0; k < 4 * NCH + 1; ++k) {  // s1[k] = dW[k - 1 - (ph & 2)]
      const float u0 = (k + 2 < ROWS + 6) ? dw[k + 2] : 0.f;   // dW[k - 1]
      const float u2 = (k < ROWS + 6) ? dw[k] : 0.f;           // dW[k - 3]
      s1[k] = (ph & 2) ? u2 : u0;
    }
#pragma unroll
    for (int k = 0; k < 4 * NCH; ++k) w[k] = (ph & 1) ? s1[k] : s1[k + 1];  // dW[k - ph]
    float* rowp = dplane + (((long long)(y >> 2) * tw) << 4) + ((y & 3) << 2);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int xc = xa + 4 * k;
      if (xc < 0 || xc >= Wl) continue;
      float4 v = make_float4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
      if (xc + 3 >= Wl) {  // the chunk hangs over the right edge: nothing is added to the padding
        if (xc + 1 >= Wl) v.y = 0.f;
        if (xc + 2 >= Wl) v.z = 0.f;
        v.w = 0.f;
      }
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rowp + ((xc >> 2) << 4)), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w)
                     : "memory");
    }
  }
}

int launch_lookup_backward(const void* const* pyr, const rcb_pyramid_layout& lay, const float* coords,
                           const float* grad_out, float* const* dpyr, float* dcoords, int B, int H, int W,
                           int radius, cudaStream_t s) {
  if (lay.dtype != RCB_F32) return RCB_ERR_UNSUPPORTED;
  const PyramidDev pd = make_pyramid_dev(pyr, lay);
  const PyramidDev dd = make_pyramid_dev(reinterpret_cast<const void* const*>(dpyr), lay);
  const long long nq = (long long)B * H * W;
  const unsigned grid = (unsigned)((nq + 3) / 4);
  const int Q = H * W;
  // pyramid gradient: the tiled red.v4 kernel; coords gradient (needs the forward taps): the per-entry kernel
  if (dpyr != nullptr) {
    dim3 g3((Q + 31) / 32, lay.levels, B);
    switch (radius) {
      case 1: lookup_backward_dpyr_kernel<1><<<g3, 128, 0, s>>>(dd, coords, grad_out, Q, lay.levels); break;
      case 2: lookup_backward_dpyr_kernel<2><<<g3, 128, 0, s>>>(dd, coords, grad_out, Q, lay.levels); break;
      case 3: lookup_backward_dpyr_kernel<3><<<g3, 128, 0, s>>>(dd, coords, grad_out, Q, lay.levels); break;
      case 4: lookup_backward_dpyr_kernel<4><<<g3, 128, 0, s>>>(dd, coords, grad_out, Q, lay.levels); break;
      default: return RCB_ERR_UNSUPPORTED;
    }
  }
  if (dcoords != nullptr) {
    switch (radius) {
      case 1: lookup_backward_kernel<1><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, 0); break;
      case 2: lookup_backward_kernel<2><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, 0); break;
      case 3: lookup_backward_kernel<3><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, 0); break;
      case 4: lookup_backward_kernel<4><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, 0); break;
      default: return RCB_ERR_UNSUPPORTED;
    }
  }
  return launch_status();
}

}  // namespace rcb
