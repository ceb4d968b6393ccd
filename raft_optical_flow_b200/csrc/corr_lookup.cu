// K2: fused multi-level correlation-window lookup (forward) and its transpose (backward).
//
// Replaces CorrBlock.__call__ (reference core/corr.py:56-94) and bilinear_sampler
// (core/utils/utils.py:57-71): per GRU iteration, for every query pixel and pyramid level a
// (2r+2)^2 tap window is gathered from that query's correlation plane and bilinearly resampled to the
// (2r+1)^2 outputs the update block consumes.  HBM-bound gather: see DESIGN.md "K2".
//
// Work decomposition
//   CTA   = 32 consecutive query pixels in flattened (h, w) order (so every output channel is written
//           as one 128-byte line) of ONE level; 128 threads, ~21 KB of shared memory, so ~9 CTAs are
//           resident per SM and their fetch / math / store phases overlap.
//   fetch = all threads issue 16-byte cp.async (L2-only, zero-fill) for the aligned chunks that cover
//           each window row; out-of-plane rows/chunks are zero-filled by the copy itself, which IS the
//           zeros-padding of grid_sample.  Planes are stored as 4x4 tiles of 64 bytes (one DRAM atom), so
//           a (2r+2)^2 window touches ~3.25^2 atoms instead of ~1.6 per row.
//   math  = lane = query, the 4 warps split the output rows.  The window of a lane is read back with
//           conflict-free LDS.128 (per-query block stride is an odd number of 16-byte units), aligned
//           with two select stages, separable bilinear weights (all (2r+1)^2 samples of one level share
//           the same fractional offset because the window offsets are integers).
//   store = lane = query => 32 lanes write 32 consecutive floats of one output channel.
#include "rcb_common.cuh"

namespace rcb {

template <int R>
struct LookupCfg {
  static constexpr int RD = 2 * R + 1;
  static constexpr int ROWS = 2 * R + 2;           // taps per axis
  static constexpr int NCH = (ROWS + 3 + 3) / 4;   // 16-byte chunks covering ROWS floats at any 4-byte phase
  static constexpr int BLK16 = ROWS * NCH + 1;     // 16-byte units per (level, query) block -- odd
  static constexpr int QT = 32;                    // queries per CTA
  static constexpr int THREADS = 128;
  static_assert((BLK16 & 1) == 1, "block stride must be odd for conflict-free LDS.128");
};

struct LevelCoord {
  int xs, ys;    // integer position of tap (0,0)
  float fx, fy;  // fractional offset shared by the whole window
};

// coords/2^l, floor and fraction.  Coordinates far outside the plane are clamped so that the integer
// conversion is defined; every tap of such a window is out of bounds and contributes zero either way.
template <int R>
RCB_DEVINL LevelCoord level_coord(float cx, float cy, int l, int Hl, int Wl) {
  const float inv = 1.0f / (float)(1 << l);  // exact power of two (core/corr.py:82: coords / 2**i)
  float x = cx * inv, y = cy * inv;
  x = fminf(fmaxf(x, -(float)(R + 8)), (float)(Wl + R + 8));
  y = fminf(fmaxf(y, -(float)(R + 8)), (float)(Hl + R + 8));
  const float x0 = floorf(x), y0 = floorf(y);
  LevelCoord c;
  c.fx = x - x0;
  c.fy = y - y0;
  c.xs = (int)x0 - R;
  c.ys = (int)y0 - R;
  return c;
}

template <int R>
__global__ void __launch_bounds__(LookupCfg<R>::THREADS)
lookup_f32_kernel(PyramidDev pyr, const float* __restrict__ coords, float* __restrict__ out, int Q, int L) {
  using Cfg = LookupCfg<R>;
  constexpr int RD = Cfg::RD, ROWS = Cfg::ROWS, NCH = Cfg::NCH, BLK16 = Cfg::BLK16, QT = Cfg::QT;
  __shared__ float4 win[QT * BLK16];  // one level of 32 query windows, 16-byte chunks
  __shared__ float s_cx[QT], s_cy[QT];

  const int tid = threadIdx.x;
  const int l = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = blockIdx.x * QT;
  const int Hl = pyr.H[l], Wl = pyr.W[l], tw = pyr.tiles_x[l];
  const long long ps = pyr.plane_stride[l];
  const float* __restrict__ base = static_cast<const float*>(pyr.ptr[l]);

  if (tid < QT) {
    const int q = q0 + tid;
    float cx = -1.0e6f, cy = -1.0e6f;  // lanes past the end of the image fetch nothing
    if (q < Q) {
      cx = __ldg(coords + (long long)(b * 2 + 0) * Q + q);
      cy = __ldg(coords + (long long)(b * 2 + 1) * Q + q);
    }
    s_cx[tid] = cx;
    s_cy[tid] = cy;
  }
  __syncthreads();

  // ---- fetch: zero-filling 16-byte async copies of every window row ------------------------
  // 4 consecutive lanes cover the (up to) 64-byte span of one window row; chunks the window does not reach
  // are not fetched at all.
  const long long q_base = (long long)b * Q + q0;
  const uint32_t dst0 = smem_u32(win);
  for (int i = tid; i < QT * ROWS * NCH; i += Cfg::THREADS) {
    const int c = i % NCH;
    const int t = i / NCH;
    const int j = t % ROWS;
    const int q = t / ROWS;
    const LevelCoord lc = level_coord<R>(s_cx[q], s_cy[q], l, Hl, Wl);
    const int ph = lc.xs & 3;
    if (4 * c >= ph + ROWS) continue;   // chunk lies beyond the last tap of this row
    const int xa = lc.xs - ph;          // 16-byte aligned start (may be negative)
    const int y = lc.ys + j;
    const int xc = xa + 4 * c;
    const bool ok = (y >= 0) && (y < Hl) && (xc >= 0) && (xc < Wl);
    const int nvalid = min(4, Wl - xc);
    // an aligned 4-float chunk of a row is exactly one 16-byte row of one 4x4 tile
    const float* src = ok ? base + (q_base + q) * ps + tile_off(y, xc, tw) : base;
    cp_async16_zfill(dst0 + (uint32_t)((q * BLK16 + j * NCH + c) * 16), src, ok ? nvalid * 4 : 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // ---- math + store: lane = query, the 4 warps split the (2r+1) output rows ------------------
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = Cfg::THREADS / 32;
  const int b_begin = (RD * warp) / NW, b_end = (RD * (warp + 1)) / NW;  // output rows (y offsets) of this warp
  if (b_begin == b_end) return;
  const bool q_ok = q0 + lane < Q;
  const LevelCoord lc = level_coord<R>(s_cx[lane], s_cy[lane], l, Hl, Wl);
  const int ph = lc.xs & 3;
  const float fx = lc.fx, fy = lc.fy, gx = 1.0f - lc.fx, gy = 1.0f - lc.fy;
  const float4* blk = win + lane * BLK16;
  float* o = out + ((long long)b * L + l) * RD * RD * Q + q0 + lane;
  float prev[RD];
  for (int j = b_begin; j <= b_end; ++j) {  // input rows b_begin .. b_end
    float wv[4 * NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float4 v = blk[j * NCH + c];
      wv[4 * c + 0] = v.x;
      wv[4 * c + 1] = v.y;
      wv[4 * c + 2] = v.z;
      wv[4 * c + 3] = v.w;
    }
    float v1[ROWS + 2];
#pragma unroll
    for (int i = 0; i < ROWS + 2; ++i) v1[i] = (ph & 1) ? wv[i + 1] : wv[i];
    float s[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) s[i] = (ph & 2) ? v1[i + 2] : v1[i];
    float t[RD];
#pragma unroll
    for (int a = 0; a < RD; ++a) t[a] = gx * s[a] + fx * s[a + 1];
    if (j > b_begin && q_ok) {
#pragma unroll
      for (int a = 0; a < RD; ++a) o[(long long)(a * RD + (j - 1)) * Q] = gy * prev[a] + fy * t[a];
    }
#pragma unroll
    for (int a = 0; a < RD; ++a) prev[a] = t[a];
  }
}

template <int R>
static int launch_lookup_r(const PyramidDev& pd, const float* coords, float* out, int B, int H, int W, int L,
                           cudaStream_t s) {
  using Cfg = LookupCfg<R>;
  const int Q = H * W;
  dim3 grid((Q + Cfg::QT - 1) / Cfg::QT, L, B);
  lookup_f32_kernel<R><<<grid, Cfg::THREADS, 0, s>>>(pd, coords, out, Q, L);
  return launch_status();
}

int launch_lookup(const void* const* pyr, const rcb_pyramid_layout& lay, const float* coords, float* out, int B,
                  int H, int W, int radius, cudaStream_t s) {
  if (lay.dtype != RCB_F32) return RCB_ERR_UNSUPPORTED;
  const PyramidDev pd = make_pyramid_dev(pyr, lay);
  switch (radius) {
    case 1: return launch_lookup_r<1>(pd, coords, out, B, H, W, lay.levels, s);
    case 2: return launch_lookup_r<2>(pd, coords, out, B, H, W, lay.levels, s);
    case 3: return launch_lookup_r<3>(pd, coords, out, B, H, W, lay.levels, s);
    case 4: return launch_lookup_r<4>(pd, coords, out, B, H, W, lay.levels, s);
    default: return RCB_ERR_UNSUPPORTED;
  }
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

int launch_lookup_backward(const void* const* pyr, const rcb_pyramid_layout& lay, const float* coords,
                           const float* grad_out, float* const* dpyr, float* dcoords, int B, int H, int W,
                           int radius, cudaStream_t s) {
  if (lay.dtype != RCB_F32) return RCB_ERR_UNSUPPORTED;
  const PyramidDev pd = make_pyramid_dev(pyr, lay);
  const PyramidDev dd = make_pyramid_dev(reinterpret_cast<const void* const*>(dpyr), lay);
  const long long nq = (long long)B * H * W;
  const unsigned grid = (unsigned)((nq + 3) / 4);
  switch (radius) {
    case 1: lookup_backward_kernel<1><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, dpyr != nullptr); break;
    case 2: lookup_backward_kernel<2><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, dpyr != nullptr); break;
    case 3: lookup_backward_kernel<3><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, dpyr != nullptr); break;
    case 4: lookup_backward_kernel<4><<<grid, 128, 0, s>>>(pd, dd, coords, grad_out, dcoords, B, H, W, lay.levels, dpyr != nullptr); break;
    default: return RCB_ERR_UNSUPPORTED;
  }
  return launch_status();
}

}  // namespace rcb
