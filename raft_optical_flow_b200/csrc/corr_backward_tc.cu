// K4, last step on the tensor cores: backward of the all-pairs contraction (the two bmm-backward GEMMs autograd
// records for reference core/corr.py:121,127):
//   dF1[b,c,q] = sum_p dV0[b,q,p] * F2[b,c,p] / sqrt(C)
//   dF2[b,c,p] = sum_q dV0[b,q,p] * F1[b,c,q] / sqrt(C)
// dV0 is the level-0 gradient pyramid (fp32, 4x4-tiled planes, rcb_pyramid_layout); its plane index p' runs in tile
// order and includes the padding of edge tiles, which lookup_backward / pool_backward never touch (it stays zero).
//
// Both are "NT" GEMMs  D[m, n] = sum_k A[m, k] * B[n, k]  with n = channel (<= 256):
//   GEMM 1: m = q,  k = p', A = dV0 as stored,      B = F2 re-ordered into tile order (zeros at the padding)
//   GEMM 2: m = p', k = q,  A = dV0 transposed,     B = F1
// and run as hi/lo bf16 splits (hi*hi + lo*hi + hi*lo, fp32 accumulation in tensor memory) like the forward build:
// error ~5e-6 of max-abs, inside the 2e-4 gradient tolerance with the same headroom as the forward pass.
//   pack kernels   fp32 -> bf16 hi/lo, K-major: dV0 [q][p'] (copy + split), dV0^T [p'][q] (32x32 smem transpose),
//                  F2 -> [c][p'] (tile order), F1 -> [c][q]; row lengths padded to 8 elements (TMA strides are
//                  multiples of 16 bytes).
//   GEMM kernel    one CTA per (128-row tile of m, batch, gemm); 192 threads: warp 0 TMA producer (A hi/lo 128x64,
//                  B hi/lo Nx64 per stage, SWIZZLE_128B), warp 1 tcgen05.mma.kind::f16 issuer (M=128, N=C, K=16,
//                  operands through shared-memory descriptors), warps 2-5 epilogue: tcgen05.ld, scale, store
//                  dF[b, n, .] with lanes = consecutive m (coalesced for GEMM 1; GEMM 2 maps p' back to (y, x) and
//                  skips the padding).
// Replaces the fp32 SIMT tiles of corr_backward.cu (18.8 TFLOP/s at cfg5) when a workspace is supplied.
#include <cstdlib>

#include "rcb_common.cuh"
#include "tcgen05_util.cuh"
#include "tma_util.cuh"

namespace rcb {

namespace bwd {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int THREADS = 192;
constexpr int MAX_N = 256;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB per part

struct Params {
  int M, K, N;        // rows of this GEMM, reduction length, channels (multiple of 16, <= 256)
  int C, Q, H, W;     // output geometry: dF[b][c][q]
  int tiles_x;        // level-0 tiles per tile row (GEMM 2: m = p' -> (y, x))
  int permute_m;      // 1: m is a tile-order plane index (GEMM 2)
  int nparts_batches; // B (operand arrays are [part][B][rows][K])
  float scale;
  float* out;
  int nstage, b_bytes, stage_bytes, bar_off;
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, b = blockIdx.y;
  const uint32_t bar0 = smem_base + p.bar_off;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 64 + 8 * s; };
  const uint32_t acc_full = bar0 + 128;
  const uint32_t tmem_slot = bar0 + 136;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 136);
  const int NS = p.nstage;
  const int kblocks = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<1>(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---- TMA producer: A hi, A lo, B hi, B lo of one 64-wide k-block per stage ----
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(empty(s), ph ^ 1);
      if (elect_one()) {
        const uint32_t st = smem_base + s * p.stage_bytes;
        mbar_expect_tx(full(s), (uint32_t)(2 * A_BYTES + 2 * p.b_bytes));
        for (int part = 0; part < 2; ++part) {
          tma_load_3d(st + part * A_BYTES, &map_a, full(s), kb * BK, mt * BM, part * p.nparts_batches + b);
          tma_load_3d(st + 2 * A_BYTES + part * p.b_bytes, &map_b, full(s), kb * BK, 0, part * p.nparts_batches + b);
        }
      }
      __syncwarp();
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    const uint32_t idesc = tc::make_idesc_mn(BM, p.N);
    const uint64_t desc_base = tc::make_smem_desc(0);
    int s = 0;
    uint32_t ph = 0, acc = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(full(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_base + s * p.stage_bytes;
        const uint64_t a_hi = desc_base | (uint64_t)((st >> 4) & 0x3FFF);
        const uint64_t a_lo = desc_base | (uint64_t)(((st + A_BYTES) >> 4) & 0x3FFF);
        const uint64_t b_hi = desc_base | (uint64_t)(((st + 2 * A_BYTES) >> 4) & 0x3FFF);
        const uint64_t b_lo = desc_base | (uint64_t)(((st + 2 * A_BYTES + p.b_bytes) >> 4) & 0x3FFF);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {  // +32 bytes per K step inside the 128-byte swizzle row
          tc::umma_bf16_ss(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          acc = 1;
          tc::umma_bf16_ss(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
          tc::umma_bf16_ss(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
        }
        tc::umma_commit<1>(empty(s));
      }
      acc = 1;
      __syncwarp();
      if (++s == NS) { s = 0; ph ^= 1; }
    }
    if (elect_one()) tc::umma_commit<1>(acc_full);
    __syncwarp();
  } else {
    // ---- epilogue: lane = row m of the tile, registers = channels ----
    const int lane_q = (warp & 3) * 32;
    const int m = mt * BM + lane_q + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    long long col = -1;  // position inside dF[b][c][.] this row maps to, -1: none
    if (m < p.M) {
      if (!p.permute_m) {
        col = m;
      } else {
        const int tile = m >> 4, r = (m >> 2) & 3, cc = m & 3;
        const int y = (tile / p.tiles_x) * 4 + r, x = (tile % p.tiles_x) * 4 + cc;
        if (y < p.H && x < p.W) col = (long long)y * p.W + x;
      }
    }
    float* out = p.out + (long long)b * p.C * p.Q;
    for (int n0 = 0; n0 < p.N; n0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)lane_q << 16) + n0, v);
      if (col >= 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n0 + i < p.C) out[(long long)(n0 + i) * p.Q + col] = v[i] * p.scale;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc::tmem_dealloc<1>(tmem_base, 256);
  }
}

// ---- operand packing ------------------------------------------------------------------------------------
RCB_DEVINL void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// dV0 [B*Q][P] fp32 -> hi/lo [B*Q][Pp] bf16 (Pp >= P, multiple of 8; the tail is zero)
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                 long long rows, int P, int Pp) {
  const long long n4 = rows * (Pp / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (Pp / 4);
    const int c = (int)(i % (Pp / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < P) v = *reinterpret_cast<const float4*>(in + r * P + c);  // P is a multiple of 16
    __nv_bfloat16 h[4], l[4];
    split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi + r * Pp + c) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(lo + r * Pp + c) = *reinterpret_cast<uint2*>(l);
  }
}

// dV0 [B][Q][P] fp32 -> transposed hi/lo [B][P][Qp] bf16 (zero tail)
__global__ void __launch_bounds__(256)
pack_transpose_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                      int Q, int P, int Qp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, q0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = in + (long long)b * Q * P;
  for (int j = ty; j < 32; j += 8) {
    const int q = q0 + j, pp = p0 + tx;
    tile[j][tx] = (q < Q && pp < P) ? __ldg(src + (long long)q * P + pp) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, q = q0 + tx;
    if (pp < P && q < Qp) {
      __nv_bfloat16 h, l;
      split_bf16(tile[tx][j], h, l);  // zero beyond Q
      hi[((long long)b * P + pp) * Qp + q] = h;
      lo[((long long)b * P + pp) * Qp + q] = l;
    }
  }
}

// F [B][C][Q] fp32 -> hi/lo [B][Np][Kp] bf16; tile_order: k = tile-order plane index (zeros at the padding)
__global__ void __launch_bounds__(256)
pack_fmap_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int C,
                 int Np, int Q, int Kp, int H, int W, int tiles_x, int tile_order, int Kvalid) {
  const long long n = (long long)gridDim.y * Np * Kp;  // gridDim.y = B
  const int b = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)Np * Kp;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / Kp), k = (int)(i % Kp);
    float v = 0.f;
    if (c < C && k < Kvalid) {
      if (tile_order) {
        const int tile = k >> 4, r = (k >> 2) & 3, cc = k & 3;
        const int y = (tile / tiles_x) * 4 + r, x = (tile % tiles_x) * 4 + cc;
        if (y < H && x < W) v = __ldg(in + ((long long)b * C + c) * Q + (long long)y * W + x);
      } else {
        v = __ldg(in + ((long long)b * C + c) * Q + k);
      }
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[(long long)b * Np * Kp + i] = h;
    lo[(long long)b * Np * Kp + i] = l;
  }
  (void)n;
}

}  // namespace bwd

struct BwdSizes {
  int Q, P, Pp, Qp, Np;
  size_t a1, a2, b1, b2;  // bytes of ONE part of each packed operand
};
static BwdSizes bwd_sizes(int B, int C, int H, int W, const rcb_pyramid_layout& lay) {
  BwdSizes z;
  z.Q = H * W;
  z.P = (int)lay.plane_stride[0];
  z.Pp = (z.P + 7) / 8 * 8;
  z.Qp = (z.Q + 7) / 8 * 8;
  z.Np = (C + 15) / 16 * 16;
  z.a1 = (size_t)B * z.Q * z.Pp * 2;  // all multiples of 16 bytes (Pp, Qp are multiples of 8)
  z.a2 = (size_t)B * z.P * z.Qp * 2;
  z.b1 = (size_t)B * z.Np * z.Pp * 2;
  z.b2 = (size_t)B * z.Np * z.Qp * 2;
  return z;
}

size_t contract_backward_tc_workspace_bytes(int B, int C, int H, int W) {
  rcb_pyramid_layout lay;
  if (fill_layout(B, H, W, 1, RCB_F32, &lay) != RCB_OK || C > bwd::MAX_N) return 0;
  const BwdSizes z = bwd_sizes(B, C, H, W, lay);
  return 2 * (z.a1 + z.a2 + z.b1 + z.b2);
}

int launch_contract_backward_tc(const float* f1, const float* f2, const float* dvol0, float* df1, float* df2, int B,
                                int C, int H, int W, void* ws, size_t ws_bytes, cudaStream_t s) {
  using namespace bwd;
  rcb_pyramid_layout lay;
  int st = fill_layout(B, H, W, 1, RCB_F32, &lay);
  if (st != RCB_OK) return st;
  if (C > MAX_N) return RCB_ERR_UNSUPPORTED;
  if (!encode_fn()) return RCB_ERR_NO_DEVICE;
  const BwdSizes z = bwd_sizes(B, C, H, W, lay);
  if (!ws || ws_bytes < 2 * (z.a1 + z.a2 + z.b1 + z.b2) || (reinterpret_cast<uintptr_t>(ws) & 255))
    return RCB_ERR_WORKSPACE;
  unsigned char* w = static_cast<unsigned char*>(ws);
  __nv_bfloat16* a1 = reinterpret_cast<__nv_bfloat16*>(w);            w += 2 * z.a1;  // [part][B][Q][Pp]
  __nv_bfloat16* a2 = reinterpret_cast<__nv_bfloat16*>(w);            w += 2 * z.a2;  // [part][B][P][Qp]
  __nv_bfloat16* b1 = reinterpret_cast<__nv_bfloat16*>(w);            w += 2 * z.b1;  // F2: [part][B][Np][Pp]
  __nv_bfloat16* b2 = reinterpret_cast<__nv_bfloat16*>(w);                            // F1: [part][B][Np][Qp]
  auto part1 = [](__nv_bfloat16* p0, size_t bytes) { return reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(p0) + bytes); };

  // 1. pack
  pack_rows_kernel<<<kNumSMs * 8, 256, 0, s>>>(dvol0, a1, part1(a1, z.a1), (long long)B * z.Q, z.P, z.Pp);
  {
    dim3 grid((z.P + 31) / 32, (z.Qp + 31) / 32, B);
    pack_transpose_kernel<<<grid, 256, 0, s>>>(dvol0, a2, part1(a2, z.a2), z.Q, z.P, z.Qp);
  }
  {
    dim3 grid(kNumSMs, B);
    pack_fmap_kernel<<<grid, 256, 0, s>>>(f2, b1, part1(b1, z.b1), C, z.Np, z.Q, z.Pp, H, W, lay.tiles_x[0], 1, z.P);
    pack_fmap_kernel<<<grid, 256, 0, s>>>(f1, b2, part1(b2, z.b2), C, z.Np, z.Q, z.Qp, H, W, lay.tiles_x[0], 0, z.Q);
  }
  st = launch_status();
  if (st != RCB_OK) return st;

  // 2. the two GEMMs
  const int b_bytes = z.Np * BK * 2;
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;
  int nstage = (227 * 1024 - 1024) / stage_bytes;
  if (nstage > 4) nstage = 4;
  if (nstage < 2) return RCB_ERR_UNSUPPORTED;
  const int smem_total = nstage * stage_bytes + 1024;
  // per launch, not once per process: the attribute belongs to the current device (nn.DataParallel drives several)
  const cudaError_t attr =
      cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (attr != cudaSuccess) return (int)attr;
  for (int g = 0; g < 2; ++g) {
    const int M = g == 0 ? z.Q : z.P, K = g == 0 ? z.P : z.Q, Kp = g == 0 ? z.Pp : z.Qp;
    const __nv_bfloat16* A = g == 0 ? a1 : a2;
    const __nv_bfloat16* Bm = g == 0 ? b1 : b2;
    const size_t a_part = g == 0 ? z.a1 : z.a2, b_part = g == 0 ? z.b1 : z.b2;
    // the hi and lo parts of an operand are stored back to back: one [2 * B] batch dimension
    CUtensorMap map_a, map_b;
    if (a_part != (size_t)B * M * Kp * 2 || b_part != (size_t)B * z.Np * Kp * 2) return RCB_ERR_WORKSPACE;
    {
      cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)M, (cuuint64_t)2 * B};
      cuuint64_t str[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)M * Kp * 2};
      cuuint32_t box[3] = {BK, BM, 1};
      if (!encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
        return RCB_ERR_INVALID_ARGUMENT;
    }
    {
      cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)z.Np, (cuuint64_t)2 * B};
      cuuint64_t str[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)z.Np * Kp * 2};
      cuuint32_t box[3] = {BK, (cuuint32_t)z.Np, 1};
      if (!encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, Bm, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
        return RCB_ERR_INVALID_ARGUMENT;
    }
    Params p{};
    p.M = M; p.K = K; p.N = z.Np;
    p.C = C; p.Q = z.Q; p.H = H; p.W = W;
    p.tiles_x = lay.tiles_x[0];
    p.permute_m = g;
    p.nparts_batches = B;
    p.scale = 1.0f / sqrtf((float)C);
    p.out = g == 0 ? df1 : df2;
    p.nstage = nstage; p.b_bytes = b_bytes; p.stage_bytes = stage_bytes; p.bar_off = nstage * stage_bytes;
    dim3 grid((M + BM - 1) / BM, B);
    gemm_nt_kernel<<<grid, THREADS, smem_total, s>>>(map_a, map_b, p);
    st = launch_status();
    if (st != RCB_OK) return st;
  }
  return RCB_OK;
}

}  // namespace rcb
