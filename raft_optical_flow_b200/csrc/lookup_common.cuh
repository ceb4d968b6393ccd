// Window geometry, the host-side lookup plan and the TMA gather configuration shared by the lookup kernels
// (corr_lookup.cu) and the fused lookup + 1x1 convolution (corr_lookup_conv.cu).
#pragma once
#include "rcb_common.cuh"
#include "tma_util.cuh"

namespace rcb {

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

struct LookupMaps {
  CUtensorMap m[RCB_MAX_LEVELS * 4];  // [level][ny_sel * 2 + nx_sel]
};

struct LookupPlan {  // host-side blob behind rcb_corr_lookup_plan_*
  LookupMaps maps;
  rcb_pyramid_layout lay;
  const void* ptr[RCB_MAX_LEVELS];
  int B, H, W, radius;
  int lanes;  // lanes per query of the fp32 kernel: 0 = chosen per launch from the grid size, 2 or 4 = pinned
  uint32_t magic;
};
constexpr uint32_t kPlanMagic = 0x52434250u;  // "RCBP"

template <int R>
struct TmaCfg {
  static constexpr int RD = 2 * R + 1;
  static constexpr int ROWS = 2 * R + 2;
  static constexpr int NMIN = (ROWS + 3) >> 2;  // tiles per axis a window overlaps: NMIN or NMIN + 1
  static constexpr int NMAX = (ROWS + 6) >> 2;
  static constexpr int SLOT_BYTES = (NMAX * NMAX * 64 + 127) / 128 * 128;  // TMA destinations are 128-byte aligned
  static constexpr int SLOT16 = SLOT_BYTES / 16;
  static constexpr int NBMAX = (RD + 3) / 4;    // output rows per lane
  static constexpr int QT = 32, THREADS = 128;
};

}  // namespace rcb
