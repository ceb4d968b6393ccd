// Next row of the scope table: RAFT's learned convex upsampling of the 1/8-resolution flow,
// RAFT.upsample_flow (reference core/raft.py:112-142): softmax over the 9 mask logits of every fine pixel, weighted
// sum of the 3x3 neighbourhood of 8*flow (F.unfold, zero padding), permute to [N, 2, 8H, 8W].
//   flow [N, 2, H, W], mask [N, 576, H, W] (channel = k*64 + i*8 + j: neighbour k = ky*3 + kx, sub-pixel (i, j))
//   out[n, c, 8h + i, 8w + j] = sum_k softmax_k(mask[n, k*64 + i*8 + j, h, w]) * 8 * flow[n, c, h + ky - 1, w + kx - 1]
// One fused pass instead of view / softmax / unfold / mul / sum / permute / reshape (seven torch kernels and four
// intermediates of the mask's size): HBM-bound, 2304 + 512 + 8 bytes per coarse pixel.
//   CTA = one coarse row h x 32 columns; thread (tx = column, ty = sub-row i) loops over the 8 sub-columns j: every
//   mask read is a 128-byte line per warp, every thread writes 32 contiguous bytes per flow channel.
// Backward: d mask through the softmax Jacobian, d flow as a gather over the 9 cells a coarse pixel contributes to.
#include "rcb_common.cuh"

namespace rcb {

__global__ void __launch_bounds__(256)
upsample_flow_kernel(const float* __restrict__ flow, const float* __restrict__ mask, float* __restrict__ out, int H,
                     int W) {
  const int w = blockIdx.x * 32 + threadIdx.x, i = threadIdx.y;
  const int h = blockIdx.y, n = blockIdx.z;
  if (w >= W) return;
  const long long HW = (long long)H * W;
  // 3x3 neighbourhood of 8 * flow, zero outside the image
  float nb[2][9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int y = h + k / 3 - 1, x = w + k % 3 - 1;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
    for (int c = 0; c < 2; ++c) nb[c][k] = ok ? 8.0f * __ldg(flow + ((long long)(n * 2 + c) * H + y) * W + x) : 0.f;
  }
  const float* m = mask + (long long)n * 576 * HW + (long long)h * W + w;
  float o[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float l[9], mx = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      l[k] = __ldg(m + (long long)(k * 64 + i * 8 + j) * HW);
      mx = fmaxf(mx, l[k]);
    }
    float s = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float e = __expf(l[k] - mx);
      s += e;
      a0 = fmaf(e, nb[0][k], a0);
      a1 = fmaf(e, nb[1][k], a1);
    }
    const float inv = 1.0f / s;
    o[0][j] = a0 * inv;
    o[1][j] = a1 * inv;
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float4* dst = reinterpret_cast<float4*>(out + (((long long)(n * 2 + c) * 8 * H + 8 * h + i) * 8 * W + 8 * w));
    dst[0] = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
    dst[1] = make_float4(o[c][4], o[c][5], o[c][6], o[c][7]);
  }
}

// d mask: for every fine pixel p_k (g.nb_k - sum_m p_m g.nb_m) with g.nb_k = sum_c gout_c * nb[c][k];
// also accumulates, per coarse pixel and neighbour k, sum_{i,j} p_k * gout_c into `pk_g` [N, 2, 9, H, W] so that the
// flow gradient becomes a 9-tap gather (second kernel) instead of a scatter.
__global__ void __launch_bounds__(256)
upsample_flow_bwd_mask_kernel(const float* __restrict__ flow, const float* __restrict__ mask,
                              const float* __restrict__ gout, float* __restrict__ dmask, float* __restrict__ pk_g,
                              int H, int W) {
  __shared__ float red[8][2 * 9][32];
  const int tx = threadIdx.x, i = threadIdx.y;
  const int w = blockIdx.x * 32 + tx;
  const int h = blockIdx.y, n = blockIdx.z;
  const long long HW = (long long)H * W;
  float acc[2][9];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;
  if (w < W) {
    float nb[2][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int y = h + k / 3 - 1, x = w + k % 3 - 1;
      const bool ok = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
      for (int c = 0; c < 2; ++c) nb[c][k] = ok ? 8.0f * __ldg(flow + ((long long)(n * 2 + c) * H + y) * W + x) : 0.f;
    }
    const float* m = mask + (long long)n * 576 * HW + (long long)h * W + w;
    float* dm = dmask + (long long)n * 576 * HW + (long long)h * W + w;
    float g[2][8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float4* src =
          reinterpret_cast<const float4*>(gout + (((long long)(n * 2 + c) * 8 * H + 8 * h + i) * 8 * W + 8 * w));
      const float4 u0 = __ldg(src), u1 = __ldg(src + 1);
      g[c][0] = u0.x; g[c][1] = u0.y; g[c][2] = u0.z; g[c][3] = u0.w;
      g[c][4] = u1.x; g[c][5] = u1.y; g[c][6] = u1.z; g[c][7] = u1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float l[9], mx = -3.0e38f;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        l[k] = __ldg(m + (long long)(k * 64 + i * 8 + j) * HW);
        mx = fmaxf(mx, l[k]);
      }
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        l[k] = __expf(l[k] - mx);
        s += l[k];
      }
      const float inv = 1.0f / s;
      float dot = 0.f, gn[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        l[k] *= inv;  // p_k
        gn[k] = g[0][j] * nb[0][k] + g[1][j] * nb[1][k];
        dot = fmaf(l[k], gn[k], dot);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        dm[(long long)(k * 64 + i * 8 + j) * HW] = l[k] * (gn[k] - dot);
        acc[0][k] = fmaf(l[k], g[0][j], acc[0][k]);
        acc[1][k] = fmaf(l[k], g[1][j], acc[1][k]);
      }
    }
  }
  // sum over the 8 sub-rows i of this coarse pixel
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) red[i][c * 9 + k][tx] = acc[c][k];
  __syncthreads();
  for (int e = i; e < 18; e += 8) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][e][tx];
    if (w < W) pk_g[((long long)(n * 2 + e / 9) * 9 + e % 9) * HW + (long long)h * W + w] = s;
  }
}

// d flow[n, c, y, x] = 8 * sum_k pk_g[n, c, k, y - (ky - 1), x - (kx - 1)]  (cells whose neighbour k is (y, x))
__global__ void __launch_bounds__(256)
upsample_flow_bwd_flow_kernel(const float* __restrict__ pk_g, float* __restrict__ dflow, int N, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)N * 2 * HW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % W), y = (int)((t / W) % H);
    const long long nc = t / HW;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int yy = y - (k / 3 - 1), xx = x - (k % 3 - 1);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += __ldg(pk_g + (nc * 9 + k) * HW + (long long)yy * W + xx);
    }
    dflow[t] = 8.0f * s;
  }
}

int launch_upsample_flow(const float* flow, const float* mask, float* out, int N, int H, int W, cudaStream_t s) {
  dim3 grid((W + 31) / 32, H, N), block(32, 8);
  upsample_flow_kernel<<<grid, block, 0, s>>>(flow, mask, out, H, W);
  return launch_status();
}

int launch_upsample_flow_backward(const float* flow, const float* mask, const float* gout, float* dflow, float* dmask,
                                  float* workspace, int N, int H, int W, cudaStream_t s) {
  dim3 grid((W + 31) / 32, H, N), block(32, 8);
  upsample_flow_bwd_mask_kernel<<<grid, block, 0, s>>>(flow, mask, gout, dmask, workspace, H, W);
  const long long total = (long long)N * 2 * H * W;
  const unsigned g = (unsigned)((total + 255) / 256 < (long long)kNumSMs * 8 ? (total + 255) / 256 : (long long)kNumSMs * 8);
  upsample_flow_bwd_flow_kernel<<<g, 256, 0, s>>>(workspace, dflow, N, H, W);
  return launch_status();
}

}  // namespace rcb
