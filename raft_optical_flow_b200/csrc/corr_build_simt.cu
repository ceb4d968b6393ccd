// K1 (fp32 SIMT mode) + pooling passes.
//
// RCB_BUILD_FP32_SIMT: the all-pairs contraction of CorrBlock.corr (reference core/corr.py:96-127)
// evaluated in true fp32 on the FMA pipe, exactly the arithmetic class of the reference's cuBLAS
// SGEMM; used as the bit-conservative mode and as the on-device cross-check of the tcgen05 modes.
// Pooled levels are produced by a separate pass per level (core/corr.py:52-54).  The tensor-core
// modes (corr_build_tc.cu) fuse both into one kernel.
#include "rcb_common.cuh"

namespace rcb {

namespace {
constexpr int BM = 128, BN = 128, BK = 8;
constexpr int THREADS = 256;
}  // namespace

// vol[b, q, p] = sum_c f1[b,c,q] * f2[b,c,p] / sqrt(C).  Both operands are "MN-major" ([C][Q] rows), so
// tiles are loaded with coalesced row reads and need no transpose.
__global__ void __launch_bounds__(THREADS)
build_simt_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ v0, int C, int Q,
                  int W, int tiles_x, long long plane_stride, float divisor) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* A = f1 + (long long)b * C * Q;
  const float* Bp = f2 + (long long)b * C * Q;
  const bool vec = (Q & 3) == 0;

  // each thread stages 4 consecutive floats of one k-row for A and for B
  const int lk = tid / 32;         // 0..7
  const int lc = (tid % 32) * 4;   // 0..124
  float4 ra, rb;
  auto load_tile = [&](int k0) {
    const int k = k0 + lk;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (k < C) {
      const float* pa = A + (long long)k * Q + m0 + lc;
      const float* pb = Bp + (long long)k * Q + n0 + lc;
      if (vec && m0 + lc + 3 < Q) ra = __ldg(reinterpret_cast<const float4*>(pa));
      else {
        if (m0 + lc + 0 < Q) ra.x = __ldg(pa + 0);
        if (m0 + lc + 1 < Q) ra.y = __ldg(pa + 1);
        if (m0 + lc + 2 < Q) ra.z = __ldg(pa + 2);
        if (m0 + lc + 3 < Q) ra.w = __ldg(pa + 3);
      }
      if (vec && n0 + lc + 3 < Q) rb = __ldg(reinterpret_cast<const float4*>(pb));
      else {
        if (n0 + lc + 0 < Q) rb.x = __ldg(pb + 0);
        if (n0 + lc + 1 < Q) rb.y = __ldg(pb + 1);
        if (n0 + lc + 2 < Q) rb.z = __ldg(pb + 2);
        if (n0 + lc + 3 < Q) rb.w = __ldg(pb + 3);
      }
    }
  };
  auto store_tile = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lk][lc]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][lk][lc]) = rb;
  };

  const int ty = tid / 16, tx = tid % 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  const int nk = (C + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  const bool chunked = (W & 3) == 0;  // 4 consecutive targets = one 16-byte tile row
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (q >= Q) continue;
    float* plane = v0 + ((long long)b * Q + q) * plane_stride;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int p = n0 + jh * 64 + tx * 4;
      float4 v = make_float4(acc[i][jh * 4 + 0] / divisor, acc[i][jh * 4 + 1] / divisor,
                             acc[i][jh * 4 + 2] / divisor, acc[i][jh * 4 + 3] / divisor);
      if (chunked && p + 3 < Q) {
        *reinterpret_cast<float4*>(plane + tile_off(p / W, p % W, tiles_x)) = v;
      } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int pe = p + e;
          if (pe < Q) plane[tile_off(pe / W, pe % W, tiles_x)] = vv[e];
        }
      }
    }
  }
}

// 2x2 floor-mode mean of every plane of one level (core/corr.py:53).
__global__ void __launch_bounds__(256)
pool_level_kernel(const float* __restrict__ in, float* __restrict__ out, long long total, int Ho, int Wo,
                  int tx_in, long long ps_in, int tx_out, long long ps_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const long long t = i / Wo;
    const int y = (int)(t % Ho);
    const long long plane = t / Ho;
    // (2y, 2x), (2y, 2x+1) share a tile row; row 2y+1 is the next 16-byte row of the same tile
    const float* s = in + plane * ps_in + tile_off(2 * y, 2 * x, tx_in);
    const float2 r0 = *reinterpret_cast<const float2*>(s);
    const float2 r1 = *reinterpret_cast<const float2*>(s + 4);
    out[plane * ps_out + tile_off(y, x, tx_out)] = ((r0.x + r0.y) + (r1.x + r1.y)) * 0.25f;
  }
}

int launch_pool_levels(void* const* pyr, const rcb_pyramid_layout& lay, int B, int H, int W, cudaStream_t s) {
  const long long planes = (long long)B * H * W;
  for (int l = 1; l < lay.levels; ++l) {
    const long long total = planes * lay.H[l] * lay.W[l];
    if (total == 0) continue;
    const long long want = (total + 255) / 256;
    const unsigned grid = (unsigned)(want < (long long)kNumSMs * 16 ? want : (long long)kNumSMs * 16);
    pool_level_kernel<<<grid, 256, 0, s>>>(static_cast<const float*>(pyr[l - 1]), static_cast<float*>(pyr[l]), total,
                                           lay.H[l], lay.W[l], lay.tiles_x[l - 1], lay.plane_stride[l - 1],
                                           lay.tiles_x[l], lay.plane_stride[l]);
  }
  return launch_status();
}

int launch_build_simt(const float* f1, const float* f2, void* const* pyr, const rcb_pyramid_layout& lay, int B,
                      int C, int H, int W, cudaStream_t s) {
  if (lay.dtype != RCB_F32) return RCB_ERR_UNSUPPORTED;
  const int Q = H * W;
  dim3 grid((Q + BN - 1) / BN, (Q + BM - 1) / BM, B);
  build_simt_kernel<<<grid, THREADS, 0, s>>>(f1, f2, static_cast<float*>(pyr[0]), C, Q, W, lay.tiles_x[0],
                                             lay.plane_stride[0], sqrtf((float)C));
  int st = launch_status();
  if (st != RCB_OK) return st;
  return launch_pool_levels(pyr, lay, B, H, W, s);
}

// ---------------------------------------------------------------------------------------------
// K4, second half: avg_pool2d backward folded coarse-to-fine, and the contraction backward.
// ---------------------------------------------------------------------------------------------
// dfine[2y+dy, 2x+dx] += dcoarse[y, x] / 4 ; rows/cols dropped by the floor receive nothing.
__global__ void __launch_bounds__(256)
pool_backward_kernel(const float* __restrict__ dcoarse, float* __restrict__ dfine, long long total, int Hc, int Wc,
                     int tx_c, long long ps_c, int tx_f, long long ps_f) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wc);
    const long long t = i / Wc;
    const int y = (int)(t % Hc);
    const long long plane = t / Hc;
    const float g = dcoarse[plane * ps_c + tile_off(y, x, tx_c)] * 0.25f;
    float* f = dfine + plane * ps_f + tile_off(2 * y, 2 * x, tx_f);
    float2 r0 = *reinterpret_cast<float2*>(f);
    float2 r1 = *reinterpret_cast<float2*>(f + 4);
    r0.x += g; r0.y += g; r1.x += g; r1.y += g;
    *reinterpret_cast<float2*>(f) = r0;
    *reinterpret_cast<float2*>(f + 4) = r1;
  }
}

int launch_pool_backward(float* const* dpyr, const rcb_pyramid_layout& lay, int B, int H, int W, cudaStream_t s) {
  const long long planes = (long long)B * H * W;
  for (int l = lay.levels - 1; l >= 1; --l) {
    const long long total = planes * lay.H[l] * lay.W[l];
    if (total == 0) continue;
    const long long want = (total + 255) / 256;
    const unsigned grid = (unsigned)(want < (long long)kNumSMs * 16 ? want : (long long)kNumSMs * 16);
    pool_backward_kernel<<<grid, 256, 0, s>>>(dpyr[l], dpyr[l - 1], total, lay.H[l], lay.W[l], lay.tiles_x[l],
                                              lay.plane_stride[l], lay.tiles_x[l - 1], lay.plane_stride[l - 1]);
  }
  return launch_status();
}

}  // namespace rcb
