/*
 * corr_oracle.c -- CPU restatement of the RAFT correlation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * kernels in raft_optical_flow_b200/csrc/.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path never calls into it and has no CPU fallback.
 *
 * Parity status: the reference ships NO tests/golden vectors for this path
 * (SURVEY.md section 4), so the oracle is pinned against outputs of the reference
 * itself: tests/golden/make_golden.py imports /root/reference/core/corr.py
 * (CorrBlock, torch CPU ops) and stores input/output fixtures under
 * tests/golden/; tests/test_oracle_golden.py checks every function below
 * against them.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout).  The arithmetic of CorrBlock lives in three PyTorch
 * ops (torch.matmul, F.avg_pool2d, F.grid_sample; PyTorch is an un-vendored
 * dependency, installed version 2.11.0); their published semantics are restated
 * here in plain C.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -fopenmp -shared).
 * All tensors are contiguous float32, C order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EXPORT __attribute__((visibility("default")))

ORC_EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORC_EXPORT void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------- *
 * All-pairs volume.  core/corr.py:96-127 (CorrBlock.corr):
 *   fmap [N,C,H,W] viewed as [N,C,Q]; corr = matmul(fmap1^T, fmap2)  (:121)
 *   corr / sqrt(tensor(dim).float())                                   (:127)
 * vol[b, q, p] = (sum_c f1[b,c,q] * f2[b,c,p]) / sqrtf(C)
 * acc64 != 0 accumulates the contraction in double (checker mode); acc64 == 0
 * accumulates in float like an SGEMM (CPU-baseline timing mode).
 * ------------------------------------------------------------------------- */
#define QB 8    /* queries per register block */
#define PB 512  /* targets per cache block     */

static void volume_block_f64(const float* f1b, const float* f2b, int C, int Q,
                             int q0, int qn, int p0, int pn, float div, float* volb) {
  double acc[QB][PB];
  for (int i = 0; i < qn; ++i)
    for (int j = 0; j < pn; ++j) acc[i][j] = 0.0;
  for (int c = 0; c < C; ++c) {
    const float* row = f2b + (size_t)c * Q + p0;
    for (int i = 0; i < qn; ++i) {
      const double a = (double)f1b[(size_t)c * Q + q0 + i];
      double* ai = acc[i];
      for (int j = 0; j < pn; ++j) ai[j] += a * (double)row[j];
    }
  }
  for (int i = 0; i < qn; ++i)
    for (int j = 0; j < pn; ++j)
      volb[(size_t)(q0 + i) * Q + p0 + j] = (float)acc[i][j] / div;
}

static void volume_block_f32(const float* f1b, const float* f2b, int C, int Q,
                             int q0, int qn, int p0, int pn, float div, float* volb) {
  float acc[QB][PB];
  for (int i = 0; i < qn; ++i)
    for (int j = 0; j < pn; ++j) acc[i][j] = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float* row = f2b + (size_t)c * Q + p0;
    for (int i = 0; i < qn; ++i) {
      const float a = f1b[(size_t)c * Q + q0 + i];
      float* ai = acc[i];
      for (int j = 0; j < pn; ++j) ai[j] += a * row[j];
    }
  }
  for (int i = 0; i < qn; ++i)
    for (int j = 0; j < pn; ++j)
      volb[(size_t)(q0 + i) * Q + p0 + j] = acc[i][j] / div;
}

ORC_EXPORT void orc_corr_volume(const float* f1, const float* f2, int B, int C, int H, int W,
                                float* vol, int acc64) {
  const int Q = H * W;
  const float div = sqrtf((float)C); /* corr.py:127 divides by a float32 sqrt */
  const int nqb = (Q + QB - 1) / QB;
  const int npb = (Q + PB - 1) / PB;
  const long ntask = (long)B * npb * nqb;
#pragma omp parallel for schedule(dynamic, 4)
  for (long t = 0; t < ntask; ++t) {
    const int b = (int)(t / ((long)npb * nqb));
    const long rem = t % ((long)npb * nqb);
    const int pb = (int)(rem / nqb);
    const int qb = (int)(rem % nqb);
    const int q0 = qb * QB, p0 = pb * PB;
    const int qn = (Q - q0 < QB) ? Q - q0 : QB;
    const int pn = (Q - p0 < PB) ? Q - p0 : PB;
    const float* f1b = f1 + (size_t)b * C * Q;
    const float* f2b = f2 + (size_t)b * C * Q;
    float* volb = vol + (size_t)b * Q * Q;
    if (acc64)
      volume_block_f64(f1b, f2b, C, Q, q0, qn, p0, pn, div, volb);
    else
      volume_block_f32(f1b, f2b, C, Q, q0, qn, p0, pn, div, volb);
  }
}

/* ------------------------------------------------------------------------- *
 * 2x2 stride-2 mean pooling, floor mode, no padding.
 * core/corr.py:52-54 (F.avg_pool2d(corr, 2, stride=2) on [N*Q,1,H,W]) and
 * core/corr.py:157-161 (same op on the feature maps for AlternateCorrBlock).
 * in [n, H, W] -> out [n, H/2, W/2]; an odd trailing row/column is dropped.
 * ATen's CPU kernel sums the window in the accumulate type (double for float
 * inputs on CPU) and divides by the window size.
 * ------------------------------------------------------------------------- */
ORC_EXPORT void orc_avg_pool2(const float* in, long n, int H, int W, float* out) {
  const int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    const float* src = in + (size_t)i * H * W;
    float* dst = out + (size_t)i * Ho * Wo;
    for (int y = 0; y < Ho; ++y)
      for (int x = 0; x < Wo; ++x) {
        const float* s = src + (size_t)(2 * y) * W + 2 * x;
        const double sum = (double)s[0] + (double)s[1] + (double)s[W] + (double)s[W + 1];
        dst[(size_t)y * Wo + x] = (float)(sum / 4.0);
      }
  }
}

/* ------------------------------------------------------------------------- *
 * One bilinear sample with zero padding, align_corners=True, pixel coordinates
 * passed through the reference's normalise / un-normalise round trip.
 * core/utils/utils.py:57-71 (bilinear_sampler):
 *     xgrid = 2*xgrid/(W-1) - 1 ; ygrid = 2*ygrid/(H-1) - 1 ; F.grid_sample(..., align_corners=True)
 * ATen grid_sampler_2d (bilinear, padding_mode=zeros, align_corners=True):
 *     ix = ((x + 1) / 2) * (W - 1); nw = floor; weights (ix_se-ix)*(iy_se-iy) ...; taps outside the
 *     plane contribute nothing.
 * When roundtrip == 0 the pixel coordinate is used directly (what the CUDA kernels do); the two
 * differ by a few ulp of the coordinate only.
 * ------------------------------------------------------------------------- */
static inline float plane_at(const float* plane, int Hl, int Wl, int y, int x) {
  return (y >= 0 && y < Hl && x >= 0 && x < Wl) ? plane[(size_t)y * Wl + x] : 0.0f;
}

static inline float sample_bilinear(const float* plane, int Hl, int Wl, float x, float y, int roundtrip) {
  float ix = x, iy = y;
  if (roundtrip) {
    const float xg = 2.0f * x / (float)(Wl - 1) - 1.0f;
    const float yg = 2.0f * y / (float)(Hl - 1) - 1.0f;
    ix = ((xg + 1.0f) / 2.0f) * (float)(Wl - 1);
    iy = ((yg + 1.0f) / 2.0f) * (float)(Hl - 1);
  }
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const float x1f = fx0 + 1.0f, y1f = fy0 + 1.0f;
  const float w_nw = (x1f - ix) * (y1f - iy);
  const float w_ne = (ix - fx0) * (y1f - iy);
  const float w_sw = (x1f - ix) * (iy - fy0);
  const float w_se = (ix - fx0) * (iy - fy0);
  float out = 0.0f;
  out += plane_at(plane, Hl, Wl, y0, x0) * w_nw;
  out += plane_at(plane, Hl, Wl, y0, x0 + 1) * w_ne;
  out += plane_at(plane, Hl, Wl, y0 + 1, x0) * w_sw;
  out += plane_at(plane, Hl, Wl, y0 + 1, x0 + 1) * w_se;
  return out;
}

/* ------------------------------------------------------------------------- *
 * Pyramid window lookup.  core/corr.py:56-94 (CorrBlock.__call__):
 *   for level i: delta = stack(meshgrid(dy, dx), -1)              (:77-79)
 *                centroid = coords.reshape(N*H*W,1,1,2) / 2**i     (:82)
 *                coords_lvl = centroid + delta                     (:84)
 *                bilinear_sampler(corr_pyramid[i], coords_lvl)     (:87)
 *   cat over levels, permute to [N, L*rd*rd, H, W]                 (:92-94)
 * meshgrid(dy,dx) is 'ij' indexed, so delta[a,b] = (dy[a], dx[b]) is added to
 * (x, y): window entry (a,b) samples (x + (a-r), y + (b-r)) -- the x offset is
 * the slow index.  Output channel = i*rd*rd + a*rd + b.
 * coords [B,2,H,W] (channel 0 = x, 1 = y, core/utils/utils.py:74-77);
 * pyr[i] [B*H*W, Hs[i], Ws[i]]; out [B, L*rd*rd, H, W].
 * ------------------------------------------------------------------------- */
ORC_EXPORT void orc_lookup(const float* const* pyr, const int* Hs, const int* Ws, const float* coords,
                           int B, int H, int W, int L, int r, float* out, int roundtrip) {
  const int rd = 2 * r + 1;
  const int Q = H * W;
  const int CH = L * rd * rd;
#pragma omp parallel for schedule(static)
  for (long bq = 0; bq < (long)B * Q; ++bq) {
    const int b = (int)(bq / Q), q = (int)(bq % Q);
    const float cx = coords[((size_t)b * 2 + 0) * Q + q];
    const float cy = coords[((size_t)b * 2 + 1) * Q + q];
    for (int l = 0; l < L; ++l) {
      const float* plane = pyr[l] + (size_t)bq * Hs[l] * Ws[l];
      const float scale = (float)(1 << l);
      const float x = cx / scale, y = cy / scale;
      for (int a = 0; a < rd; ++a)
        for (int bb = 0; bb < rd; ++bb) {
          const float sx = x + (float)(a - r);
          const float sy = y + (float)(bb - r);
          const float v = sample_bilinear(plane, Hs[l], Ws[l], sx, sy, roundtrip);
          out[((size_t)b * CH + (size_t)l * rd * rd + a * rd + bb) * Q + q] = v;
        }
    }
  }
}

/* ------------------------------------------------------------------------- *
 * On-the-fly correlation, forward.
 * alt_cuda_corr/correlation_kernel.cu:18-119 (corr_forward_kernel) and launcher :260-286.
 *   fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] -> corr [B,N,rd*rd,H1,W1], zero-initialised (:273).
 *   For every integer tap (iy,ix) in [0,rd]^2 around floor(coords)-r (:75-76) the dot product of the
 *   query vector with fmap2 at the tap (0 outside fmap2, :80-83), accumulated over 32-channel chunks
 *   (:44,:88-90), is splatted into up to four output channels with weights
 *   dy*dx, dy*(1-dx), (1-dy)*dx, (1-dy)*(1-dx) (:92-114).  Channel index = iy + rd*ix (:92-95).
 *   The result is NOT divided by sqrt(C); the caller does that (core/corr.py:198).
 * ------------------------------------------------------------------------- */
ORC_EXPORT void orc_altcorr_forward(const float* fmap1, const float* fmap2, const float* coords,
                                    int B, int N, int H1, int W1, int H2, int W2, int C, int r, float* corr) {
  const int rd = 2 * r + 1;
  const size_t plane = (size_t)H1 * W1;
  memset(corr, 0, sizeof(float) * (size_t)B * N * rd * rd * plane);
#pragma omp parallel for schedule(static)
  for (long bq = 0; bq < (long)B * H1 * W1; ++bq) {
    const int b = (int)(bq / (H1 * W1));
    const int hw = (int)(bq % (H1 * W1));
    const float* q = fmap1 + (size_t)bq * C;
    for (int n = 0; n < N; ++n) {
      const float* cp = coords + (((size_t)b * N + n) * plane + hw) * 2;
      const float x = cp[0], y = cp[1];
      const float fx0 = floorf(x), fy0 = floorf(y);
      const float dx = x - fx0, dy = y - fy0;
      float* out = corr + ((size_t)b * N + n) * rd * rd * plane + hw;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int cn = (C - c0 < 32) ? C - c0 : 32;
        for (int iy = 0; iy < rd + 1; ++iy)
          for (int ix = 0; ix < rd + 1; ++ix) {
            const int h2 = (int)fy0 - r + iy;
            const int w2 = (int)fx0 - r + ix;
            float s = 0.0f;
            if (h2 >= 0 && h2 < H2 && w2 >= 0 && w2 < W2) {
              const float* t = fmap2 + (((size_t)b * H2 + h2) * W2 + w2) * C + c0;
              for (int k = 0; k < cn; ++k) s += q[c0 + k] * t[k];
            }
            const float nw = s * dy * dx, ne = s * dy * (1 - dx);
            const float sw = s * (1 - dy) * dx, se = s * (1 - dy) * (1 - dx);
            if (iy > 0 && ix > 0) out[(size_t)((iy - 1) + rd * (ix - 1)) * plane] += nw;
            if (iy > 0 && ix < rd) out[(size_t)((iy - 1) + rd * ix) * plane] += ne;
            if (iy < rd && ix > 0) out[(size_t)(iy + rd * (ix - 1)) * plane] += sw;
            if (iy < rd && ix < rd) out[(size_t)(iy + rd * ix) * plane] += se;
          }
      }
    }
  }
}

/* ------------------------------------------------------------------------- *
 * On-the-fly correlation, backward.
 * alt_cuda_corr/correlation_kernel.cu:122-256 (corr_backward_kernel), launcher :288-324.
 *   g(tap) = bilinear-weighted sum of the (up to) four corr_grad entries the tap was splatted to (:204-222)
 *   fmap1_grad[q,:]   += g * fmap2[tap,:]        (:224-226, :245-254)
 *   fmap2_grad[tap,:] += g * fmap1[q,:]  (atomicAdd; taps outside fmap2 skipped, :229-238)
 *   coords_grad is allocated zero and never written (:130,:307,:323).
 * want_coords_grad != 0 additionally fills coords_grad with the TRUE derivative
 *   d corr / d coords (what autograd through CorrBlock yields) instead of the
 *   reference kernel's zeros:  d/dx of the bilinear weights applied to the tap dot products.
 * Accumulation is in double, results rounded to float (order-independent checker).
 * ------------------------------------------------------------------------- */
ORC_EXPORT void orc_altcorr_backward(const float* fmap1, const float* fmap2, const float* coords,
                                     const float* corr_grad, int B, int N, int H1, int W1, int H2, int W2,
                                     int C, int r, float* fmap1_grad, float* fmap2_grad, float* coords_grad,
                                     int want_coords_grad) {
  const int rd = 2 * r + 1;
  const size_t plane = (size_t)H1 * W1;
  const size_t n1 = (size_t)B * H1 * W1 * C, n2 = (size_t)B * H2 * W2 * C;
  double* g1 = (double*)calloc(n1, sizeof(double));
  double* g2 = (double*)calloc(n2, sizeof(double));
  memset(coords_grad, 0, sizeof(float) * (size_t)B * N * plane * 2);
  /* serial over queries: fmap2_grad is a scatter (the reference uses atomics) */
  for (long bq = 0; bq < (long)B * H1 * W1; ++bq) {
    const int b = (int)(bq / (H1 * W1));
    const int hw = (int)(bq % (H1 * W1));
    const float* q = fmap1 + (size_t)bq * C;
    for (int n = 0; n < N; ++n) {
      const float* cp = coords + (((size_t)b * N + n) * plane + hw) * 2;
      const float x = cp[0], y = cp[1];
      const float fx0 = floorf(x), fy0 = floorf(y);
      const float dx = x - fx0, dy = y - fy0;
      const float* gp = corr_grad + ((size_t)b * N + n) * rd * rd * plane + hw;
      double gcx = 0.0, gcy = 0.0;
      for (int iy = 0; iy < rd + 1; ++iy)
        for (int ix = 0; ix < rd + 1; ++ix) {
          const int h2 = (int)fy0 - r + iy;
          const int w2 = (int)fx0 - r + ix;
          float g = 0.0f;   /* d loss / d s(tap) */
          float gdx = 0.0f; /* sum_k grad_k * d weight_k / d dx */
          float gdy = 0.0f;
          if (iy > 0 && ix > 0) {
            const float go = gp[(size_t)((iy - 1) + rd * (ix - 1)) * plane];
            g += go * dy * dx; gdx += go * dy; gdy += go * dx;
          }
          if (iy > 0 && ix < rd) {
            const float go = gp[(size_t)((iy - 1) + rd * ix) * plane];
            g += go * dy * (1 - dx); gdx -= go * dy; gdy += go * (1 - dx);
          }
          if (iy < rd && ix > 0) {
            const float go = gp[(size_t)(iy + rd * (ix - 1)) * plane];
            g += go * (1 - dy) * dx; gdx += go * (1 - dy); gdy -= go * dx;
          }
          if (iy < rd && ix < rd) {
            const float go = gp[(size_t)(iy + rd * ix) * plane];
            g += go * (1 - dy) * (1 - dx); gdx -= go * (1 - dy); gdy -= go * (1 - dx);
          }
          if (!(h2 >= 0 && h2 < H2 && w2 >= 0 && w2 < W2)) continue;
          const size_t toff = (((size_t)b * H2 + h2) * W2 + w2) * C;
          const float* t = fmap2 + toff;
          double s = 0.0;
          for (int k = 0; k < C; ++k) {
            g1[(size_t)bq * C + k] += (double)g * (double)t[k];
            g2[toff + k] += (double)g * (double)q[k];
            s += (double)q[k] * (double)t[k];
          }
          gcx += s * (double)gdx;
          gcy += s * (double)gdy;
        }
      if (want_coords_grad) {
        float* cg = coords_grad + (((size_t)b * N + n) * plane + hw) * 2;
        cg[0] = (float)gcx;
        cg[1] = (float)gcy;
      }
    }
  }
  for (size_t i = 0; i < n1; ++i) fmap1_grad[i] = (float)g1[i];
  for (size_t i = 0; i < n2; ++i) fmap2_grad[i] = (float)g2[i];
  free(g1);
  free(g2);
}

/* ------------------------------------------------------------------------- *
 * Backward of CorrBlock (what autograd records for train.py:212 through
 * core/corr.py:25-127 and core/utils/utils.py:57-71): grid_sampler_2d_backward per level,
 * avg_pool2d_backward x(L-1), bmm backward x2 and the 1/sqrt(C) scale.
 *   grad_out [B, L*rd*rd, H, W]  ->  dF1, dF2 [B,C,H,W], dcoords [B,2,H,W]
 * pyr[i] are the forward pyramid planes (needed for the coords gradient).
 * dvol_scratch[i] must hold B*Q*Hs[i]*Ws[i] doubles (caller-allocated, zeroed here).
 * The coords gradient is d/d(pixel coordinate): the (W-1)/2 of grid_sample's un-normalise and the
 * 2/(W-1) of bilinear_sampler's normalise cancel, leaving the 1/2^i level factor.
 * Small sizes only (dense double scratch).
 * ------------------------------------------------------------------------- */
ORC_EXPORT void orc_corrblock_backward(const float* f1, const float* f2, const float* const* pyr,
                                       const int* Hs, const int* Ws, const float* coords,
                                       const float* grad_out, int B, int C, int H, int W, int L, int r,
                                       double* const* dvol_scratch, float* df1, float* df2,
                                       float* dcoords) {
  const int rd = 2 * r + 1;
  const int Q = H * W;
  const int CH = L * rd * rd;
  for (int l = 0; l < L; ++l)
    memset(dvol_scratch[l], 0, sizeof(double) * (size_t)B * Q * Hs[l] * Ws[l]);
  /* 1. bilinear backward: scatter into the dense per-level grads, coords grad */
  for (long bq = 0; bq < (long)B * Q; ++bq) {
    const int b = (int)(bq / Q), q = (int)(bq % Q);
    const float cx = coords[((size_t)b * 2 + 0) * Q + q];
    const float cy = coords[((size_t)b * 2 + 1) * Q + q];
    double gx = 0.0, gy = 0.0;
    for (int l = 0; l < L; ++l) {
      const int Hl = Hs[l], Wl = Ws[l];
      const float* plane = pyr[l] + (size_t)bq * Hl * Wl;
      double* dplane = dvol_scratch[l] + (size_t)bq * Hl * Wl;
      const float scale = (float)(1 << l);
      const float x = cx / scale, y = cy / scale;
      double lgx = 0.0, lgy = 0.0;
      for (int a = 0; a < rd; ++a)
        for (int bb = 0; bb < rd; ++bb) {
          const float go = grad_out[((size_t)b * CH + (size_t)l * rd * rd + a * rd + bb) * Q + q];
          const float sx = x + (float)(a - r), sy = y + (float)(bb - r);
          const float fx0 = floorf(sx), fy0 = floorf(sy);
          const int x0 = (int)fx0, y0 = (int)fy0;
          const double tx = (double)sx - (double)fx0, ty = (double)sy - (double)fy0;
          const double w[4] = {(1 - tx) * (1 - ty), tx * (1 - ty), (1 - tx) * ty, tx * ty};
          const double wx[4] = {-(1 - ty), (1 - ty), -ty, ty};
          const double wy[4] = {-(1 - tx), -tx, (1 - tx), tx};
          const int ys[4] = {y0, y0, y0 + 1, y0 + 1};
          const int xs[4] = {x0, x0 + 1, x0, x0 + 1};
          for (int k = 0; k < 4; ++k) {
            if (ys[k] < 0 || ys[k] >= Hl || xs[k] < 0 || xs[k] >= Wl) continue;
            const size_t off = (size_t)ys[k] * Wl + xs[k];
            dplane[off] += (double)go * w[k];
            lgx += (double)go * wx[k] * (double)plane[off];
            lgy += (double)go * wy[k] * (double)plane[off];
          }
        }
      gx += lgx / (double)scale;
      gy += lgy / (double)scale;
    }
    dcoords[((size_t)b * 2 + 0) * Q + q] = (float)gx;
    dcoords[((size_t)b * 2 + 1) * Q + q] = (float)gy;
  }
  /* 2. avg_pool2d backward, coarse to fine: each fine cell of a pooled 2x2 block receives grad/4 */
  for (int l = L - 1; l >= 1; --l) {
    const int Hc = Hs[l], Wc = Ws[l], Hf = Hs[l - 1], Wf = Ws[l - 1];
    for (long bq = 0; bq < (long)B * Q; ++bq) {
      const double* dc = dvol_scratch[l] + (size_t)bq * Hc * Wc;
      double* df = dvol_scratch[l - 1] + (size_t)bq * Hf * Wf;
      for (int y = 0; y < Hc; ++y)
        for (int x = 0; x < Wc; ++x) {
          const double g = dc[(size_t)y * Wc + x] * 0.25;
          df[(size_t)(2 * y) * Wf + 2 * x] += g;
          df[(size_t)(2 * y) * Wf + 2 * x + 1] += g;
          df[(size_t)(2 * y + 1) * Wf + 2 * x] += g;
          df[(size_t)(2 * y + 1) * Wf + 2 * x + 1] += g;
        }
    }
  }
  /* 3. contraction backward: dF1[c,q] = sum_p dV[q,p] F2[c,p] / sqrt(C); dF2[c,p] = sum_q dV[q,p] F1[c,q] / sqrt(C) */
  const double inv = 1.0 / (double)sqrtf((float)C);
  for (int b = 0; b < B; ++b) {
    const double* dv = dvol_scratch[0] + (size_t)b * Q * Q;
    const float* f1b = f1 + (size_t)b * C * Q;
    const float* f2b = f2 + (size_t)b * C * Q;
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c) {
      for (int q = 0; q < Q; ++q) {
        double s = 0.0;
        for (int p = 0; p < Q; ++p) s += dv[(size_t)q * Q + p] * (double)f2b[(size_t)c * Q + p];
        df1[((size_t)b * C + c) * Q + q] = (float)(s * inv);
      }
      for (int p = 0; p < Q; ++p) {
        double s = 0.0;
        for (int q = 0; q < Q; ++q) s += dv[(size_t)q * Q + p] * (double)f1b[(size_t)c * Q + q];
        df2[((size_t)b * C + c) * Q + p] = (float)(s * inv);
      }
    }
  }
}
