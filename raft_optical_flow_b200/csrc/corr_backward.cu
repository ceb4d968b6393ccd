// K4, last step: backward of the all-pairs contraction (the two bmm-backward GEMMs autograd records for
// reference core/corr.py:121,127):
//   dF1[b,c,q] = sum_p dV0[b,q,p] * F2[b,c,p] / sqrt(C)        ("NT": both operands K(p)-contiguous)
//   dF2[b,c,p] = sum_q dV0[b,q,p] * F1[b,c,q] / sqrt(C)        ("NN": B operand N(p)-contiguous)
// fp32 SIMT tiles; dV0 is read in the level-0 pyramid layout (4x4 tiles).
#include "rcb_common.cuh"

namespace rcb {

namespace {
constexpr int GM = 64, GN = 64, GK = 32, GT = 256;
}

// out[m][n] = scale * sum_k A[m][k] * Bop(k, n)
//   A: [M][K] row-major (lda = K)
//   B_NT:  Bop(k,n) = Bm[n * ldb_plane + pad(k)]   (dV0[q=n][p=k])
//   !B_NT: Bop(k,n) = Bm[k * ldb_plane + pad(n)]   (dV0[q=k][p=n])
// pad(p) = tile_off(p / W, p % W) maps a flattened target index to the 4x4-tiled plane.
template <bool B_NT>
__global__ void __launch_bounds__(GT)
contract_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ out, int M, int Nn,
                    int K, long long strideA, long long strideB, long long strideO, long long ldb_plane, int W,
                    int tiles_x, float scale) {
  __shared__ float As[GK][GM + 1];
  __shared__ float Bs[GK][GN + 1];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  A += b * strideA;
  Bm += b * strideB;
  out += b * strideO;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    // A tile: GM x GK, k fastest
    for (int i = tid; i < GM * GK; i += GT) {
      const int kk = i % GK, mm = i / GK;
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? __ldg(A + (long long)m * K + k) : 0.f;
    }
    if (B_NT) {
      for (int i = tid; i < GN * GK; i += GT) {
        const int kk = i % GK, nn = i / GK;
        const int n = n0 + nn, k = k0 + kk;
        float v = 0.f;
        if (n < Nn && k < K) v = __ldg(Bm + (long long)n * ldb_plane + tile_off(k / W, k % W, tiles_x));
        Bs[kk][nn] = v;
      }
    } else {
      for (int i = tid; i < GN * GK; i += GT) {
        const int nn = i % GN, kk = i / GN;
        const int n = n0 + nn, k = k0 + kk;
        float v = 0.f;
        if (n < Nn && k < K) v = __ldg(Bm + (long long)k * ldb_plane + tile_off(n / W, n % W, tiles_x));
        Bs[kk][nn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < Nn) out[(long long)m * Nn + n] = acc[i][j] * scale;
    }
  }
}

int launch_contract_backward(const float* f1, const float* f2, const float* dvol0, float* df1, float* df2, int B,
                             int C, int H, int W, cudaStream_t s) {
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, 1, RCB_F32, &lay);
  if (st != RCB_OK) return st;
  const int Q = H * W;
  const float scale = 1.0f / sqrtf((float)C);
  const long long sF = (long long)C * Q, sV = (long long)Q * lay.plane_stride[0];
  dim3 grid((Q + GN - 1) / GN, (C + GM - 1) / GM, B);
  contract_bwd_kernel<true><<<grid, GT, 0, s>>>(f2, dvol0, df1, C, Q, Q, sF, sV, sF, lay.plane_stride[0], W,
                                                lay.tiles_x[0], scale);
  contract_bwd_kernel<false><<<grid, GT, 0, s>>>(f1, dvol0, df2, C, Q, Q, sF, sV, sF, lay.plane_stride[0], W,
                                                 lay.tiles_x[0], scale);
  return launch_status();
}

}  // namespace rcb
