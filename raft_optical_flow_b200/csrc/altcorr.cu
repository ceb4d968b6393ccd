// K3 / K5: on-the-fly correlation (the alt_cuda_corr extension), forward and backward.
//
// Replaces corr_forward_kernel / corr_backward_kernel (reference alt_cuda_corr/correlation_kernel.cu:18-119,
// 122-256) and, in its fused form, the per-level loop of AlternateCorrBlock.__call__
// (reference core/corr.py:163-198).  fp32 SIMT arithmetic, exactly the reference's arithmetic class.
//
// Forward decomposition
//   CTA  = 32 consecutive query pixels (flattened h,w) of one (batch, slot); slot = coords set n
//          (extension API) or pyramid level (fused API).  4 warps, 8 queries per warp, one at a time.
//   warp = one query: every lane keeps C/32 channels of the query vector in registers; each of the
//          (2r+2)^2 taps is one coalesced C*4-byte read of fmap2 (NHWC), partial dot products are
//          reduced 32 taps at a time with a butterfly transpose-reduce (31 shuffles per 32 taps instead
//          of 160), the tap dots go to shared memory and the (2r+1)^2 bilinear outputs are formed from
//          them (weights dy*dx ... as in correlation_kernel.cu:97-100).
//   store = outputs are staged in shared memory as [channel][query] and written as 128-byte lines.
// Neighbouring queries of a CTA read overlapping taps, which the L1 serves.
#include "rcb_common.cuh"

namespace rcb {

namespace {
constexpr int QT = 32;       // queries per CTA
constexpr int THREADS = 128; // 4 warps
constexpr int MAXV = 4;      // float4 per lane: C <= 512
}

struct AltFwdParams {
  const float* f1;                 // [B, Q, C]
  const float* f2[RCB_MAX_LEVELS]; // per slot: [B, H2, W2, C]
  int H2[RCB_MAX_LEVELS], W2[RCB_MAX_LEVELS];
  const float* coords;             // planar: [B,2,Q];  interleaved: [B,N,Q,2]
  float* out;                      // [B, slots, RD*RD, Q]
  int Q, C, slots;
  float out_scale;
  int planar;                      // 1: fused pyramid API (coords/2^slot), 0: extension API
};

// Butterfly transpose-reduce: on entry every lane holds 32 partial sums v[0..31] (one per tap); on exit
// the return value of lane t is sum over lanes of v[t].
RCB_DEVINL float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 16];
    const float keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 8];
    const float keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 4];
    const float keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? v[i] : v[i + 2];
    const float keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];  // lane t now holds tap index t (bit k of t selected the upper half at stage k)
}

RCB_DEVINL void clamp_floor(float x, float y, int H2, int W2, int R, int& xs, int& ys, float& dx, float& dy) {
  x = fminf(fmaxf(x, -(float)(R + 8)), (float)(W2 + R + 8));
  y = fminf(fmaxf(y, -(float)(R + 8)), (float)(H2 + R + 8));
  const float x0 = floorf(x), y0 = floorf(y);
  dx = x - x0;
  dy = y - y0;
  xs = (int)x0 - R;
  ys = (int)y0 - R;
}

// NV > 0: channel count known at compile time (C = 128 * NV): every tap is two immediate-offset loads and 4 * NV
// FMAs per lane, tap positions and row pointers are resolved at compile time / once per window row.
// NV = 0: any C that is a multiple of 4 (<= 512), runtime bounds.
template <int R, int NV>
__global__ void __launch_bounds__(THREADS) altcorr_fwd_kernel(AltFwdParams p) {
  constexpr int RD = 2 * R + 1, T = 2 * R + 2, NT = T * T;
  constexpr int NG = (NT + 31) / 32;
  __shared__ float s_dot[THREADS / 32][NG * 32];
  __shared__ float s_out[RD * RD][QT + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * QT;
  const int H2 = p.H2[slot], W2 = p.W2[slot];
  const float* __restrict__ f2 = p.f2[slot] + (long long)b * H2 * W2 * p.C;
  const int nv = p.C >> 7;  // float4 per lane (C multiple of 128) -- remainder handled below
  const int C4 = p.C >> 2;

  for (int qi = warp; qi < QT; qi += THREADS / 32) {
    const int q = q0 + qi;
    if (q >= p.Q) break;  // warp-uniform
    float cx, cy;
    if (p.planar) {
      const float inv = 1.0f / (float)(1 << slot);
      cx = __ldg(p.coords + ((long long)b * 2 + 0) * p.Q + q) * inv;
      cy = __ldg(p.coords + ((long long)b * 2 + 1) * p.Q + q) * inv;
    } else {
      const float2 c = __ldg(reinterpret_cast<const float2*>(p.coords) + ((long long)b * p.slots + slot) * p.Q + q);
      cx = c.x;
      cy = c.y;
    }
    int xs, ys;
    float dx, dy;
    clamp_floor(cx, cy, H2, W2, R, xs, ys, dx, dy);

    // query vector: lane holds float4 #(lane + 32*k)
    float4 qv[MAXV];
    const float4* f1q = reinterpret_cast<const float4*>(p.f1 + ((long long)b * p.Q + q) * p.C);
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
      qv[k] = (lane + 32 * k < C4) ? __ldg(f1q + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    (void)nv;

    if constexpr (NV > 0) {
      constexpr int CV = 32 * NV;  // float4 per pixel
      const float4* __restrict__ f2v = reinterpret_cast<const float4*>(f2) + lane;
      // validity of the window's rows / columns, one bit each (warp-uniform)
      uint32_t rows_ok = 0, cols_ok = 0;
#pragma unroll
      for (int i = 0; i < T; ++i) {
        rows_ok |= (uint32_t)(ys + i >= 0 && ys + i < H2) << i;
        cols_ok |= (uint32_t)(xs + i >= 0 && xs + i < W2) << i;
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        float part[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int tap = g * 32 + t;  // compile-time after unrolling
          float s = 0.f;
          if (tap < NT) {
            const int iy = tap / T, ix = tap % T;
            if (((rows_ok >> iy) & (cols_ok >> ix) & 1u) != 0u) {
              const float4* tp = f2v + ((long long)(ys + iy) * W2 + xs) * CV + ix * CV;
#pragma unroll
              for (int k = 0; k < NV; ++k) {
                const float4 v = __ldg(tp + 32 * k);
                s = fmaf(qv[k].x, v.x, s);
                s = fmaf(qv[k].y, v.y, s);
                s = fmaf(qv[k].z, v.z, s);
                s = fmaf(qv[k].w, v.w, s);
              }
            }
          }
          part[t] = s;
        }
        const float d = transpose_reduce32(part, lane);
        s_dot[warp][g * 32 + lane] = d;
      }
    } else
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
      float part[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int tap = g * 32 + t;
        float s = 0.f;
        if (tap < NT) {
          const int iy = tap / T, ix = tap % T;
          const int y = ys + iy, x = xs + ix;
          if (y >= 0 && y < H2 && x >= 0 && x < W2) {  // warp-uniform
            const float4* tp = reinterpret_cast<const float4*>(f2 + ((long long)y * W2 + x) * p.C);
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
              if (lane + 32 * k < C4) {
                const float4 v = __ldg(tp + lane + 32 * k);
                s = fmaf(qv[k].x, v.x, s);
                s = fmaf(qv[k].y, v.y, s);
                s = fmaf(qv[k].z, v.z, s);
                s = fmaf(qv[k].w, v.w, s);
              }
            }
          }
        }
        part[t] = s;
      }
      const float d = transpose_reduce32(part, lane);
      // after the butterfly, lane L holds tap index with bit4..bit0 = (L&16,L&8,L&4,L&2,L&1) -> tap = L
      s_dot[warp][g * 32 + lane] = d;
    }
    __syncwarp();
    // bilinear outputs, channel = iy + RD*ix (correlation_kernel.cu:92-95)
    for (int e = lane; e < RD * RD; e += 32) {
      const int iy = e % RD, ix = e / RD;
      const float* d = s_dot[warp];
      const float v00 = d[iy * T + ix], v01 = d[iy * T + ix + 1];
      const float v10 = d[(iy + 1) * T + ix], v11 = d[(iy + 1) * T + ix + 1];
      const float o = (1.f - dy) * (1.f - dx) * v00 + (1.f - dy) * dx * v01 + dy * (1.f - dx) * v10 + dy * dx * v11;
      s_out[e][qi] = o * p.out_scale;
    }
    __syncwarp();
  }
  __syncthreads();
  float* out = p.out + ((long long)b * p.slots + slot) * RD * RD * p.Q + q0;
  for (int i = tid; i < RD * RD * QT; i += THREADS) {
    const int e = i / QT, qi = i % QT;
    if (q0 + qi < p.Q) out[(long long)e * p.Q + qi] = s_out[e][qi];
  }
}

template <int R>
static int launch_fwd_r(const AltFwdParams& p, int B, cudaStream_t s) {
  dim3 grid((p.Q + QT - 1) / QT, p.slots, B);
  switch (p.C) {
    case 128: altcorr_fwd_kernel<R, 1><<<grid, THREADS, 0, s>>>(p); break;  // RAFT-small
    case 256: altcorr_fwd_kernel<R, 2><<<grid, THREADS, 0, s>>>(p); break;  // RAFT-full
    default: altcorr_fwd_kernel<R, 0><<<grid, THREADS, 0, s>>>(p); break;
  }
  return launch_status();
}

static int launch_fwd(const AltFwdParams& p, int B, int r, cudaStream_t s) {
  if (p.C % 4 != 0 || p.C > 128 * MAXV) return RCB_ERR_UNSUPPORTED;
  switch (r) {
    case 1: return launch_fwd_r<1>(p, B, s);
    case 2: return launch_fwd_r<2>(p, B, s);
    case 3: return launch_fwd_r<3>(p, B, s);
    case 4: return launch_fwd_r<4>(p, B, s);
    default: return RCB_ERR_UNSUPPORTED;
  }
}

int launch_altcorr_forward(const float* f1, const float* f2, const float* coords, float* corr, int B, int N, int H1,
                           int W1, int H2, int W2, int C, int r, cudaStream_t s) {
  // the extension API allows any N; slots ride on gridDim.y, the per-slot table only needs one entry
  int st = RCB_OK;
  for (int n0 = 0; n0 < N && st == RCB_OK; n0 += RCB_MAX_LEVELS) {
    AltFwdParams p{};
    const int ns = (N - n0 < RCB_MAX_LEVELS) ? N - n0 : RCB_MAX_LEVELS;
    p.f1 = f1;
    for (int i = 0; i < RCB_MAX_LEVELS; ++i) { p.f2[i] = f2; p.H2[i] = H2; p.W2[i] = W2; }
    p.Q = H1 * W1; p.C = C; p.slots = ns; p.out_scale = 1.0f; p.planar = 0;
    if (N <= RCB_MAX_LEVELS) {
      p.coords = coords; p.out = corr;
      st = launch_fwd(p, B, r, s);
    } else {
      // chunks of the N axis are not contiguous across batch: launch per batch element
      for (int b = 0; b < B && st == RCB_OK; ++b) {
        const int rd2 = (2 * r + 1) * (2 * r + 1);
        AltFwdParams pb = p;
        pb.f1 = f1 + (long long)b * p.Q * C;
        for (int i = 0; i < RCB_MAX_LEVELS; ++i) pb.f2[i] = f2 + (long long)b * H2 * W2 * C;
        pb.coords = coords + ((long long)b * N + n0) * p.Q * 2;
        pb.out = corr + ((long long)b * N + n0) * rd2 * p.Q;
        st = launch_fwd(pb, 1, r, s);
      }
    }
  }
  return st;
}

int launch_altcorr_pyramid_forward(const float* f1n, const float* const* f2n, const float* coords, float* out,
                                   int B, int C, int H, int W, int levels, int r, float scale, cudaStream_t s) {
  AltFwdParams p{};
  p.f1 = f1n;
  int h = H, w = W;
  for (int l = 0; l < RCB_MAX_LEVELS; ++l) {
    p.f2[l] = l < levels ? f2n[l] : nullptr;
    p.H2[l] = h; p.W2[l] = w;
    h /= 2; w /= 2;
  }
  p.coords = coords; p.out = out; p.Q = H * W; p.C = C; p.slots = levels; p.out_scale = scale; p.planar = 1;
  return launch_fwd(p, B, r, s);
}

// ---------------------------------------------------------------------------------------------
// Prepare: NCHW -> NHWC transpose of fmap1, NCHW -> NHWC pooled feature pyramid of fmap2
// (core/corr.py:157-161 pools, :183-184 permute(0,2,3,1).contiguous() -- here once per frame pair
// instead of once per level per iteration).  Level l is the 2x2 floor-mode mean of level l-1.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int Q) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int q0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, q = q0 + tx;
    tile[i][tx] = (c < C && q < Q) ? in[((long long)b * C + c) * Q + q] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int q = q0 + i, c = c0 + tx;
    if (q < Q && c < C) out[((long long)b * Q + q) * C + c] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256)
pool_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, long long total, int C, int Hi, int Wi,
                 int Ho, int Wo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int x = (int)(t % Wo);
    t /= Wo;
    const int y = (int)(t % Ho);
    const long long b = t / Ho;
    const float* s = in + ((b * Hi + 2 * y) * Wi + 2 * x) * C + c;
    out[i] = ((s[0] + s[C]) + (s[(long long)Wi * C] + s[(long long)Wi * C + C])) * 0.25f;
  }
}

int launch_altcorr_prepare(const float* f1, const float* f2, float* f1n, float* const* f2n, int B, int C, int H,
                           int W, int levels, cudaStream_t s) {
  const int Q = H * W;
  dim3 grid((Q + 31) / 32, (C + 31) / 32, B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, s>>>(f1, f1n, C, Q);
  nchw_to_nhwc_kernel<<<grid, 256, 0, s>>>(f2, f2n[0], C, Q);
  int h = H, w = W;
  for (int l = 1; l < levels; ++l) {
    const int ho = h / 2, wo = w / 2;
    const long long total = (long long)B * ho * wo * C;
    if (total > 0) {
      const long long want = (total + 255) / 256;
      const unsigned g = (unsigned)(want < (long long)kNumSMs * 16 ? want : (long long)kNumSMs * 16);
      pool_nhwc_kernel<<<g, 256, 0, s>>>(f2n[l - 1], f2n[l], total, C, h, w, ho, wo);
    }
    h = ho; w = wo;
  }
  return launch_status();
}

// ---------------------------------------------------------------------------------------------
// Backward (K5).  One warp per (query, coords set).  g(tap) = bilinear-weighted sum of the (up to four)
// corr_grad entries the tap was splatted to (correlation_kernel.cu:204-222);
//   fmap1_grad[q,:]   = sum_tap g * fmap2[tap,:]                 (register accumulation, plain store)
//   fmap2_grad[tap,:] += g * fmap1[q,:]                          (vector red.global.add, taps in bounds)
//   coords_grad[q]    = sum_tap <fmap1[q], fmap2[tap]> * d g-weights / d(dx,dy)   (TRUE_CG only;
//                       the reference leaves zeros, correlation_kernel.cu:307)
// fmap1_grad / fmap2_grad / coords_grad must be zeroed before the launch (done by the launcher).
// ---------------------------------------------------------------------------------------------
template <int R, bool TRUE_CG>
__global__ void __launch_bounds__(THREADS)
altcorr_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ coords,
                   const float* __restrict__ cg, float* __restrict__ g1, float* __restrict__ g2,
                   float* __restrict__ gc, int N, int Q, int H2, int W2, int C) {
  constexpr int RD = 2 * R + 1, T = 2 * R + 2, NT = T * T;
  __shared__ float s_g[THREADS / 32][NT], s_gdx[THREADS / 32][NT], s_gdy[THREADS / 32][NT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int q = blockIdx.x * (THREADS / 32) + warp;
  if (q >= Q) return;
  const int C4 = C >> 2;
  const float4* f1q = reinterpret_cast<const float4*>(f1 + ((long long)b * Q + q) * C);
  float4 qv[MAXV], acc[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    qv[k] = (lane + 32 * k < C4) ? __ldg(f1q + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* f2b = f2 + (long long)b * H2 * W2 * C;
  float* g2b = g2 + (long long)b * H2 * W2 * C;
  for (int n = 0; n < N; ++n) {
    const float2 c = __ldg(reinterpret_cast<const float2*>(coords) + ((long long)b * N + n) * Q + q);
    int xs, ys;
    float dx, dy;
    clamp_floor(c.x, c.y, H2, W2, R, xs, ys, dx, dy);
    const float* gp = cg + ((long long)b * N + n) * RD * RD * Q + q;
    __syncwarp();
    for (int tap = lane; tap < NT; tap += 32) {
      const int iy = tap / T, ix = tap % T;
      float g = 0.f, gdx = 0.f, gdy = 0.f;
      if (iy > 0 && ix > 0) {
        const float go = __ldg(gp + (long long)((iy - 1) + RD * (ix - 1)) * Q);
        g += go * dy * dx; gdx += go * dy; gdy += go * dx;
      }
      if (iy > 0 && ix < RD) {
        const float go = __ldg(gp + (long long)((iy - 1) + RD * ix) * Q);
        g += go * dy * (1.f - dx); gdx -= go * dy; gdy += go * (1.f - dx);
      }
      if (iy < RD && ix > 0) {
        const float go = __ldg(gp + (long long)(iy + RD * (ix - 1)) * Q);
        g += go * (1.f - dy) * dx; gdx += go * (1.f - dy); gdy -= go * dx;
      }
      if (iy < RD && ix < RD) {
        const float go = __ldg(gp + (long long)(iy + RD * ix) * Q);
        g += go * (1.f - dy) * (1.f - dx); gdx -= go * (1.f - dy); gdy -= go * (1.f - dx);
      }
      s_g[warp][tap] = g; s_gdx[warp][tap] = gdx; s_gdy[warp][tap] = gdy;
    }
    __syncwarp();
    float cgx = 0.f, cgy = 0.f;
    for (int tap = 0; tap < NT; ++tap) {
      const int y = ys + tap / T, x = xs + tap % T;
      if (y < 0 || y >= H2 || x < 0 || x >= W2) continue;  // warp-uniform
      const float g = s_g[warp][tap];
      const long long off = ((long long)y * W2 + x) * C;
      const float4* tp = reinterpret_cast<const float4*>(f2b + off);
      float4* gt = reinterpret_cast<float4*>(g2b + off);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (lane + 32 * k < C4) {
          const float4 v = __ldg(tp + lane + 32 * k);
          acc[k].x = fmaf(g, v.x, acc[k].x); acc[k].y = fmaf(g, v.y, acc[k].y);
          acc[k].z = fmaf(g, v.z, acc[k].z); acc[k].w = fmaf(g, v.w, acc[k].w);
          atomicAdd(gt + lane + 32 * k, make_float4(g * qv[k].x, g * qv[k].y, g * qv[k].z, g * qv[k].w));
          if (TRUE_CG) {
            s = fmaf(qv[k].x, v.x, s); s = fmaf(qv[k].y, v.y, s);
            s = fmaf(qv[k].z, v.z, s); s = fmaf(qv[k].w, v.w, s);
          }
        }
      }
      if (TRUE_CG) {
        cgx = fmaf(s, s_gdx[warp][tap], cgx);
        cgy = fmaf(s, s_gdy[warp][tap], cgy);
      }
    }
    if (TRUE_CG) {
      cgx = warp_sum(cgx);
      cgy = warp_sum(cgy);
      if (lane == 0)
        reinterpret_cast<float2*>(gc)[((long long)b * N + n) * Q + q] = make_float2(cgx, cgy);
    }
  }
  float4* g1q = reinterpret_cast<float4*>(g1 + ((long long)b * Q + q) * C);
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (lane + 32 * k < C4) g1q[lane + 32 * k] = acc[k];
}

template <int R>
static int launch_bwd_r(const float* f1, const float* f2, const float* coords, const float* cg, float* g1, float* g2,
                        float* gc, int B, int N, int Q, int H2, int W2, int C, int true_cg, cudaStream_t s) {
  dim3 grid((Q + THREADS / 32 - 1) / (THREADS / 32), 1, B);
  if (true_cg)
    altcorr_bwd_kernel<R, true><<<grid, THREADS, 0, s>>>(f1, f2, coords, cg, g1, g2, gc, N, Q, H2, W2, C);
  else
    altcorr_bwd_kernel<R, false><<<grid, THREADS, 0, s>>>(f1, f2, coords, cg, g1, g2, gc, N, Q, H2, W2, C);
  return launch_status();
}

int launch_altcorr_backward(const float* f1, const float* f2, const float* coords, const float* cg, float* g1,
                            float* g2, float* gc, int B, int N, int H1, int W1, int H2, int W2, int C, int r,
                            int true_cg, cudaStream_t s) {
  if (C % 4 != 0 || C > 128 * MAXV) return RCB_ERR_UNSUPPORTED;
  const int Q = H1 * W1;
  cudaError_t e = cudaMemsetAsync(g2, 0, sizeof(float) * (size_t)B * H2 * W2 * C, s);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(gc, 0, sizeof(float) * (size_t)B * N * Q * 2, s);
  if (e != cudaSuccess) return (int)e;
  switch (r) {
    case 1: return launch_bwd_r<1>(f1, f2, coords, cg, g1, g2, gc, B, N, Q, H2, W2, C, true_cg, s);
    case 2: return launch_bwd_r<2>(f1, f2, coords, cg, g1, g2, gc, B, N, Q, H2, W2, C, true_cg, s);
    case 3: return launch_bwd_r<3>(f1, f2, coords, cg, g1, g2, gc, B, N, Q, H2, W2, C, true_cg, s);
    case 4: return launch_bwd_r<4>(f1, f2, coords, cg, g1, g2, gc, B, N, Q, H2, W2, C, true_cg, s);
    default: return RCB_ERR_UNSUPPORTED;
  }
}

}  // namespace rcb
