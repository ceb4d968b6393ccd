// Lookup fused with its only consumer: the first layer of the motion encoder,
//   cor = relu(convc1(corr))            reference core/update.py:136,154 (small) and 182,202 (basic),
// a 1x1 convolution over the L*(2r+1)^2 correlation channels of CorrBlock.__call__ (core/corr.py:56-94).
// The 73 MB correlation tensor of every GRU iteration is never written to or re-read from HBM: the window
// samples of 128 queries go straight into the A operand of a tensor-core GEMM
//   out[b, n, q] = act( sum_k corr[b, k, q] * Wt[n, k] + bias[n] ),   n < Cout <= 256.
//
//   CTA     = 128 consecutive queries of one batch element, all levels; 576 threads:
//   warps 0-15 gather + resample exactly like lookup_tma_kernel (one TMA box per (query, level) into the warp's own
//              slots, 4 lanes per query, a warp owns 8 queries through all levels), but the samples are rounded to
//              fp16 and stored into the shared-memory A operand of the level instead of global memory; one mbarrier
//              per level tells the MMA warp that its K range is complete.  Afterwards the same warps are the
//              epilogue: tcgen05.ld (lane = query), bias, ReLU, 128-byte coalesced stores of out[b, n, q..q+31].
//   warp 16    producer of the packed weights (B operand): one bulk copy per level.
//   warp 17    tcgen05.mma.kind::f16 issuer (M = 128, N = Cout, K = 16), fp32 accumulator in tensor memory; the
//              MMAs of level l run while the other warps resample level l + 1.
// Both operands use the un-swizzled K-major core-matrix layout ([K/8][rows][8 halfs], 16-byte rows): the K extent of
// a level (96 or 64) is then free of the 64-element swizzle atom, the A operand needs two level buffers (48 KB)
// instead of the whole K range (96 KB), and the 8 queries of a warp write 128 contiguous bytes per word.  What is
// saved goes to the gather ring: 16 warps x 8 KB in flight, which is what bounds this kernel (HBM latency).
// K layout (private to this kernel; pack_convc1_kernel permutes the weights to match): inside a level entry
// b*RP + a holds window sample (dx = a - r, dy = b - r), i.e. reference channel l*(2r+1)^2 + a*(2r+1) + b; RP = 2r+2
// pads a window row to whole 32-bit words, KL rounds the level to a multiple of 16.  Padding entries are written as
// zeros on the A side and are zero in the packed weights.
// Arithmetic: fp16 operands (11-bit significands, the class of the TF32 convolution cuDNN runs for the reference
// by default), fp32 accumulation.  fp32 pyramids only.
#include <cstdlib>

#include "lookup_common.cuh"
#include "rcb_common.cuh"
#include "tcgen05_util.cuh"
#include "tma_util.cuh"

namespace rcb {
namespace lconv {

template <int R>
struct Cfg {
  using G = TmaCfg<R>;
  static constexpr int RD = G::RD;
  static constexpr int RP = RD + 1;                        // window row padded to an even number of halfs
  static constexpr int KL = (RD * RP + 15) / 16 * 16;      // K entries per level: 96 (r = 4), 64 (r = 3)
  static constexpr int BM = 128;
  static constexpr int MATH_WARPS = 8;                     // two groups of 8 queries each
  static constexpr int EPI_WARPS = 4;                       // one per tensor-memory lane quarter
  static constexpr int THREADS = 32 * (MATH_WARPS + EPI_WARPS + 2);
  static constexpr int SLOT_BYTES = G::SLOT_BYTES;
  static constexpr int WARP_RING = 8 * SLOT_BYTES;
  static constexpr int MAX_N = 256;
  static constexpr int A_LBO = BM * 16;                    // bytes between K chunks of 8: [K/8][128 rows][16 B]
  static constexpr int A_BUF_BYTES = (KL / 8) * A_LBO;     // one level
  static constexpr int B_BYTES = (KL / 8) * MAX_N * 16;    // one level of weights: [K/8][N][16 B]
  static constexpr int OFF_B = 2 * A_BUF_BYTES;
  static constexpr int OFF_RING = OFF_B + B_BYTES;
  static constexpr int OFF_BIAS = OFF_RING + 2 * MATH_WARPS * WARP_RING;
  static constexpr int OFF_BAR = OFF_BIAS + MAX_N * 4;
  // barriers: gather[2 per warp], level_done[4], a_free[2], acc_full[2], d_free[2], b_full[2], b_empty[2], then the
  // tensor-memory slot
  static constexpr int NBAR = 2 * MATH_WARPS + RCB_MAX_LEVELS + 10;
  static constexpr int SMEM_BYTES = OFF_BAR + 8 * NBAR + 16;
  static constexpr int SMEM_ALLOC = SMEM_BYTES + 128;      // the base is rounded up to 128 bytes (TMA destinations)
};

// K-major operand without swizzle: 8-row x 16-byte core matrices; lbo = bytes between the two K chunks of one
// MMA, sbo = bytes between 8-row groups
RCB_DEVINL uint64_t make_desc_interleaved(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46);
}

RCB_DEVINL void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int R>
__global__ void __launch_bounds__(Cfg<R>::THREADS, 1)
lookup_conv_kernel(const __grid_constant__ LookupMaps maps, PyramidDev pyr, const float* __restrict__ coords,
                   const __half* __restrict__ wpack, const float* __restrict__ bias, float* __restrict__ out, int Q,
                   int L, int N, int relu, int tiles_q, int ntiles, int dbg) {
  using C = Cfg<R>;
  using G = typename C::G;
  constexpr int RD = C::RD, RP = C::RP, KL = C::KL, ROWS = G::ROWS, NMIN = G::NMIN, NMAX = G::NMAX;
  constexpr int NBMAX = G::NBMAX, MW = C::MATH_WARPS;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  unsigned char* smem = smem_raw + (base - raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = base + C::OFF_BAR;
  auto gbar = [&](int w) { return bar0 + 8 * w; };
  auto level_done = [&](int l) { return bar0 + 8 * (2 * MW + l); };
  auto a_free = [&](int i) { return bar0 + 8 * (2 * MW + RCB_MAX_LEVELS + i); };
  auto acc_full = [&](int d) { return bar0 + 8 * (2 * MW + RCB_MAX_LEVELS + 2 + d); };
  auto d_free = [&](int d) { return bar0 + 8 * (2 * MW + RCB_MAX_LEVELS + 4 + d); };
  auto b_full = [&](int i) { return bar0 + 8 * (2 * MW + RCB_MAX_LEVELS + 6 + i); };
  auto b_empty = [&](int i) { return bar0 + 8 * (2 * MW + RCB_MAX_LEVELS + 8 + i); };
  constexpr int EW0 = MW, PW = MW + C::EPI_WARPS, TW = MW + C::EPI_WARPS + 1;  // first epilogue / producer / MMA warp
  const uint32_t tmem_slot = bar0 + 8 * C::NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * C::NBAR);
  float* sbias = reinterpret_cast<float*>(smem + C::OFF_BIAS);

  if (tid == 0) {
    for (int w = 0; w < 2 * MW; ++w) mbar_init(gbar(w), 8);
    for (int l = 0; l < RCB_MAX_LEVELS; ++l) mbar_init(level_done(l), MW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_free(i), 1);
      mbar_init(acc_full(i), 1);
      mbar_init(d_free(i), C::EPI_WARPS);
      mbar_init(b_full(i), 1);
      mbar_init(b_empty(i), 1);
    }
    fence_barrier_init();
  }
  if (tid < N) sbias[tid] = bias ? __ldg(bias + tid) : 0.f;
  if (warp == TW) tc::tmem_alloc<1>(tmem_slot, 512);  // two accumulators: tile i and tile i + 1
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int tile0 = blockIdx.x, tstep = gridDim.x;

  if (warp == PW) {
    // ---- weight producer: half a level's [KL/16][N][8] fp16 block per bulk copy, two buffers, so the weights of a
    // level are resident before its samples are and the MMAs start the moment the level completes ----
    const uint32_t bytes = (uint32_t)(KL / 2 * N * 2);
    int c = 0;
    for (int tile = tile0; tile < ntiles; tile += tstep) {
      for (int l2 = 0; l2 < 2 * L; ++l2, ++c) {
        mbar_wait(b_empty(c & 1), (uint32_t)(((c >> 1) & 1) ^ 1));
        if (elect_one()) {
          mbar_expect_tx(b_full(c & 1), bytes);
          bulk_load(base + C::OFF_B + (c & 1) * (C::B_BYTES / 2), wpack + (long long)l2 * (KL / 2) * N, bytes,
                    b_full(c & 1));
        }
        __syncwarp();
      }
    }
  } else if (warp == TW) {
    // ---- MMA issuer: KL/16 K steps per level as soon as the level's samples are in place ----
    const uint32_t idesc = tc::make_idesc_f16_mn(C::BM, N);
    int g = 0, i = 0, c = 0;
    for (int tile = tile0; tile < ntiles; tile += tstep, ++i) {
      const int d = i & 1;
      if (i >= 2) mbar_wait(d_free(d), (uint32_t)(((i >> 1) - 1) & 1));  // the epilogue of tile i - 2 has read it
      for (int l = 0; l < L; ++l, ++g) {
        mbar_wait(level_done(l), (uint32_t)(i & 1));
#pragma unroll
        for (int hf = 0; hf < 2; ++hf, ++c) {
          mbar_wait(b_full(c & 1), (uint32_t)((c >> 1) & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = base + (g & 1) * C::A_BUF_BYTES + hf * (KL / 16) * C::A_LBO;
            const uint32_t b_addr = base + C::OFF_B + (c & 1) * (C::B_BYTES / 2);
#pragma unroll
            for (int t = 0; t < KL / 32; ++t) {
              if (dbg & 1) break;  // timing experiment: no MMAs
              const uint64_t adesc = make_desc_interleaved(a_addr + t * 2 * C::A_LBO, C::A_LBO, 128);
              const uint64_t bdesc = make_desc_interleaved(b_addr + t * 2 * N * 16, (uint32_t)N * 16, 128);
              tc::umma_bf16_ss(tmem_base + 256 * d, adesc, bdesc, idesc, (l > 0 || hf > 0 || t > 0) ? 1u : 0u);
            }
            tc::umma_commit<1>(b_empty(c & 1));
            if (hf == 1) {
              tc::umma_commit<1>(a_free(g & 1));
              if (l == L - 1) tc::umma_commit<1>(acc_full(d));
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= EW0) {
    // ---- epilogue: lane = query, registers = output channels; bias, ReLU, 128-byte coalesced stores ----
    const int quarter = warp & 3;
    int i = 0;
    for (int tile = tile0; tile < ntiles; tile += tstep, ++i) {
      const int d = i & 1;
      const int bb = tile / tiles_q, qe = (tile % tiles_q) * C::BM + quarter * 32 + lane;
      mbar_wait(acc_full(d), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      float* o = out + (long long)bb * N * Q + qe;
      for (int n0 = 0; n0 < N; n0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + 256 * d + ((uint32_t)(quarter * 32) << 16) + n0, v);
        if (n0 + 32 >= N) {  // the accumulator may be overwritten by the MMAs of tile i + 2
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_free(d));
        }
        if (qe < Q && !(dbg & 8)) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            if (n0 + k < N) {
              float x = v[k] + sbias[n0 + k];
              if (relu) x = fmaxf(x, 0.f);
              o[(long long)(n0 + k) * Q] = x;
            }
          }
        }
      }
    }
  } else {
    // ---- gather + resample into the level's A operand ----
    // A warp owns two groups of 8 queries (h = 0, 1), each with its own slots and mbarrier: while it resamples one
    // group, the gather of the other is in flight, and every finished chunk immediately issues the gather of the
    // same group's next level (or of the next tile's first level).
    const int ql = lane >> 2, sub = lane & 3;
    const int b0 = (RD * sub) >> 2, nb = ((RD * (sub + 1)) >> 2) - b0;  // output rows [b0, b0 + nb), nb <= NBMAX

    struct Pending {
      LevelCoord lc;
      int Hl, Wl;
      bool ok;
    };
    auto load_coords = [&](int tile, int h, float& cx, float& cy) {
      const int bb = tile / tiles_q, q = (tile % tiles_q) * C::BM + (warp + MW * h) * 8 + ql;
      cx = cy = -1.0e6f;
      if (tile < ntiles && q < Q) {
        cx = __ldg(coords + (long long)(bb * 2 + 0) * Q + q);
        cy = __ldg(coords + (long long)(bb * 2 + 1) * Q + q);
      }
    };
    auto issue = [&](int tile, int l, int h, float cx, float cy) {
      Pending p;
      const int bb = tile / tiles_q, q = (tile % tiles_q) * C::BM + (warp + MW * h) * 8 + ql;
      p.ok = q < Q;
      p.Hl = l == 0 ? pyr.H[0] : l == 1 ? pyr.H[1] : l == 2 ? pyr.H[2] : pyr.H[3];
      p.Wl = l == 0 ? pyr.W[0] : l == 1 ? pyr.W[1] : l == 2 ? pyr.W[2] : pyr.W[3];
      p.lc = level_coord<R>(cx, cy, l, p.Hl, p.Wl);
      if (sub == 0) {
        const uint32_t bar = gbar(2 * warp + h);
        if (p.ok && !(dbg & 2)) {
          const int nx = ((p.lc.xs & 3) + ROWS + 3) >> 2, ny = ((p.lc.ys & 3) + ROWS + 3) >> 2;
          mbar_expect_tx(bar, (uint32_t)(nx * ny * 64));
          tma_load_3d(base + C::OFF_RING + (2 * warp + h) * C::WARP_RING + ql * C::SLOT_BYTES,
                      &maps.m[l * 4 + (ny - NMIN) * 2 + (nx - NMIN)], bar, (p.lc.xs >> 2) * 16, p.lc.ys >> 2,
                      bb * Q + q);
        } else {
          mbar_arrive(bar);
        }
      }
      return p;
    };
    // resamples the gathered windows of group h and writes them into rows of A buffer `abuf`
    auto consume = [&](const Pending& cur, int h, int abuf) {
      if (!cur.ok || (dbg & 4)) return;
      const float4* slot =
          reinterpret_cast<const float4*>(smem + C::OFF_RING + (2 * warp + h) * C::WARP_RING + ql * C::SLOT_BYTES);
      const int m = (warp + MW * h) * 8 + ql;  // row of the A tile
      const LevelCoord lc = cur.lc;
      const int Hl = cur.Hl, Wl = cur.Wl;
      const int ph = lc.xs & 3, py = lc.ys & 3;
      const int nx = (ph + ROWS + 3) >> 2;
      const float fx = lc.fx, fy = lc.fy, gx = 1.0f - lc.fx, gy = 1.0f - lc.fy;
      const bool ragged_w = (Wl & 3) != 0;
      unsigned char* arow = smem + abuf * C::A_BUF_BYTES + m * 16;
      auto a_store = [&](int kbyte, uint32_t v) {  // kbyte: byte offset inside the level's K row, multiple of 4
        *reinterpret_cast<uint32_t*>(arow + (kbyte >> 4) * C::A_LBO + (kbyte & 15)) = v;
      };
      float hp[RD];
#pragma unroll
      for (int jj = 0; jj <= NBMAX; ++jj) {
        if (jj > nb) break;
        const int j = b0 + jj;  // window row
        const int ya = py + j;  // row inside the fetched box
        const bool row_ok = lc.ys + j < Hl;
        const float4* rowp = slot + ((ya >> 2) * nx) * 4 + (ya & 3);
        float w[4 * NMAX];
#pragma unroll
        for (int k = 0; k < NMAX; ++k) {
          float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && k < nx) u = rowp[k * 4];
          w[4 * k + 0] = u.x; w[4 * k + 1] = u.y; w[4 * k + 2] = u.z; w[4 * k + 3] = u.w;
        }
        float v1[ROWS + 2];
#pragma unroll
        for (int k = 0; k < ROWS + 2; ++k) v1[k] = (ph & 1) ? w[k + 1] : w[k];
        float t[ROWS];
#pragma unroll
        for (int k = 0; k < ROWS; ++k) t[k] = (ph & 2) ? v1[k + 2] : v1[k];
        if (ragged_w) {
#pragma unroll
          for (int k = 0; k < ROWS; ++k)
            if (lc.xs + k >= Wl) t[k] = 0.f;
        }
        float hh[RD];
#pragma unroll
        for (int a = 0; a < RD; ++a) hh[a] = gx * t[a] + fx * t[a + 1];
        if (jj > 0) {
          const int kb = (j - 1) * RP * 2;
#pragma unroll
          for (int a2 = 0; a2 < RP / 2; ++a2) {
            const float o0 = gy * hp[2 * a2] + fy * hh[2 * a2];
            const float o1 = (2 * a2 + 1 < RD) ? gy * hp[(2 * a2 + 1) % RD] + fy * hh[(2 * a2 + 1) % RD] : 0.f;
            const __half2 hv = __floats2half2_rn(o0, o1);
            a_store(kb + 4 * a2, *reinterpret_cast<const uint32_t*>(&hv));
          }
        }
#pragma unroll
        for (int a = 0; a < RD; ++a) hp[a] = hh[a];
      }
      if (sub == 3) {  // the level's trailing K padding must be finite: zeros
#pragma unroll
        for (int k = 0; k < (KL - RD * RP) / 2; ++k) a_store(RD * RP * 2 + 4 * k, 0u);
      }
    };
    float cx0, cy0, cx1, cy1, cxn0, cyn0, cxn1, cyn1;
    load_coords(tile0, 0, cx0, cy0);
    load_coords(tile0, 1, cx1, cy1);
    Pending pend0 = issue(tile0, 0, 0, cx0, cy0);
    Pending pend1 = issue(tile0, 0, 1, cx1, cy1);
    int cnt = 0, g = 0;
    for (int tile = tile0; tile < ntiles; tile += tstep) {
      const int next_tile = tile + tstep;
      load_coords(next_tile, 0, cxn0, cyn0);
      load_coords(next_tile, 1, cxn1, cyn1);
      for (int l = 0; l < L; ++l, ++g, ++cnt) {
        if (g >= 2) mbar_wait(a_free(g & 1), (uint32_t)(((g >> 1) - 1) & 1));  // the MMAs two levels back have read it
        // group 0
        mbar_wait(gbar(2 * warp), (uint32_t)(cnt & 1));
        consume(pend0, 0, g & 1);
        __syncwarp();  // every lane is done with the slots before the next gather lands in them
        if (l + 1 < L) {
          pend0 = issue(tile, l + 1, 0, cx0, cy0);
        } else if (next_tile < ntiles) {
          pend0 = issue(next_tile, 0, 0, cxn0, cyn0);
        }
        // group 1
        mbar_wait(gbar(2 * warp + 1), (uint32_t)(cnt & 1));
        consume(pend1, 1, g & 1);
        fence_proxy_async_smem();  // A writes (both groups) -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive(level_done(l));
        if (l + 1 < L) {
          pend1 = issue(tile, l + 1, 1, cx1, cy1);
        } else if (next_tile < ntiles) {
          pend1 = issue(next_tile, 0, 1, cxn1, cyn1);
        }
      }
      cx0 = cxn0; cy0 = cyn0; cx1 = cxn1; cy1 = cyn1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TW) {
    tc_fence_after();
    tc::tmem_dealloc<1>(tmem_base, 512);
  }
}

// weight[n][l*RD*RD + a*RD + b] (fp32, the Conv2d weight of convc1 viewed [Cout, Cin]) -> wp[l][k/8][n][k%8] fp16
// with k = b*RP + a inside the level, zero padded
__global__ void __launch_bounds__(256)
pack_convc1_kernel(const float* __restrict__ w, __half* __restrict__ wp, int cout, int L, int RD, int RP, int KL) {
  const long long n_el = (long long)L * KL * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    const int n = (int)((i >> 3) % cout);
    const int k8 = (int)((i >> 3) / cout % (KL / 8));
    const int l = (int)(i / ((long long)KL * cout));
    const int r = k8 * 8 + e;
    const int bb = r / RP, a = r % RP;
    float v = 0.f;
    if (bb < RD && a < RD) v = __ldg(w + (long long)n * (L * RD * RD) + l * RD * RD + a * RD + bb);
    wp[i] = __float2half_rn(v);
  }
}

inline bool geometry(int radius, int* rd, int* rp, int* kl) {
  if (radius != 3 && radius != 4) return false;
  *rd = 2 * radius + 1;
  *rp = *rd + 1;
  *kl = (*rd * *rp + 15) / 16 * 16;
  return true;
}

template <int R>
static int launch_r(const LookupPlan& plan, const PyramidDev& pd, const float* coords, const __half* wpack,
                    const float* bias, float* out, int cout, int relu, cudaStream_t s) {
  using C = Cfg<R>;
  const int Q = plan.H * plan.W;
  cudaError_t e = cudaFuncSetAttribute(lookup_conv_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_ALLOC);
  if (e != cudaSuccess) return (int)e;
  const int tiles_q = (Q + C::BM - 1) / C::BM, ntiles = tiles_q * plan.B;
  static const int dbg = getenv("RCB_LCONV_DEBUG") ? atoi(getenv("RCB_LCONV_DEBUG")) : 0;  // timing experiments
  const int grid = ntiles < kNumSMs ? ntiles : kNumSMs;  // persistent: one CTA per SM walks tiles grid apart
  lookup_conv_kernel<R><<<grid, C::THREADS, C::SMEM_ALLOC, s>>>(plan.maps, pd, coords, wpack, bias, out, Q,
                                                                 plan.lay.levels, cout, relu, tiles_q, ntiles, dbg);
  return launch_status();
}

}  // namespace lconv

size_t convc1_pack_bytes(int cout, int levels, int radius) {
  int rd, rp, kl;
  if (!lconv::geometry(radius, &rd, &rp, &kl) || cout < 16 || cout > 256 || cout % 16 || levels < 1 ||
      levels > RCB_MAX_LEVELS)
    return 0;
  return (size_t)levels * kl * cout * sizeof(__half);
}

int launch_convc1_pack(const float* weight, void* wpack, int cout, int levels, int radius, cudaStream_t s) {
  int rd, rp, kl;
  if (!weight || !wpack) return RCB_ERR_INVALID_ARGUMENT;
  if (convc1_pack_bytes(cout, levels, radius) == 0 || !lconv::geometry(radius, &rd, &rp, &kl)) return RCB_ERR_UNSUPPORTED;
  const long long n_el = (long long)levels * kl * cout;
  lconv::pack_convc1_kernel<<<(int)((n_el + 255) / 256), 256, 0, s>>>(weight, static_cast<__half*>(wpack), cout, levels,
                                                                     rd, rp, kl);
  return launch_status();
}

int launch_lookup_convc1(const void* plan_, const float* coords, const void* wpack, const float* bias, float* out,
                         int cout, int relu, cudaStream_t s) {
  const LookupPlan* plan = static_cast<const LookupPlan*>(plan_);
  if (!plan || plan->magic != kPlanMagic || !coords || !wpack || !out) return RCB_ERR_INVALID_ARGUMENT;
  if (plan->lay.dtype != RCB_F32) return RCB_ERR_UNSUPPORTED;
  int rd, rp, kl;
  if (!lconv::geometry(plan->radius, &rd, &rp, &kl) || cout < 16 || cout > 256 || cout % 16) return RCB_ERR_UNSUPPORTED;
  if (((uintptr_t)wpack & 15) != 0) return RCB_ERR_INVALID_ARGUMENT;
  const PyramidDev pd = make_pyramid_dev(plan->ptr, plan->lay);
  const __half* wp = static_cast<const __half*>(wpack);
  if (plan->radius == 3) return lconv::launch_r<3>(*plan, pd, coords, wp, bias, out, cout, relu, s);
  return lconv::launch_r<4>(*plan, pd, coords, wp, bias, out, cout, relu, s);
}

}  // namespace rcb
