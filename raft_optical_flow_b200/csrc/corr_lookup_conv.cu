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
  static constexpr int MATH_WARPS = 16;                    // 8 queries each
  static constexpr int THREADS = 32 * (MATH_WARPS + 2);
  static constexpr int SLOT_BYTES = G::SLOT_BYTES;
  static constexpr int WARP_RING = 8 * SLOT_BYTES;
  static constexpr int MAX_N = 256;
  static constexpr int A_LBO = BM * 16;                    // bytes between K chunks of 8: [K/8][128 rows][16 B]
  static constexpr int A_BUF_BYTES = (KL / 8) * A_LBO;     // one level
  static constexpr int B_BYTES = (KL / 8) * MAX_N * 16;    // one level of weights: [K/8][N][16 B]
  static constexpr int OFF_B = 2 * A_BUF_BYTES;
  static constexpr int OFF_RING = OFF_B + B_BYTES;
  static constexpr int OFF_BIAS = OFF_RING + MATH_WARPS * WARP_RING;
  static constexpr int OFF_BAR = OFF_BIAS + MAX_N * 4;
  // barriers: gather[16], level_done[4], a_free[2], b_full, b_empty, acc_full, then the tensor-memory slot
  static constexpr int NBAR = MATH_WARPS + RCB_MAX_LEVELS + 2 + 3;
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
                   int L, int N, int relu) {
  using C = Cfg<R>;
  using G = typename C::G;
  constexpr int RD = C::RD, RP = C::RP, KL = C::KL, ROWS = G::ROWS, NMIN = G::NMIN, NMAX = G::NMAX;
  constexpr int NBMAX = G::NBMAX, MW = C::MATH_WARPS;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  unsigned char* smem = smem_raw + (base - raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * C::BM;
  const uint32_t bar0 = base + C::OFF_BAR;
  auto gbar = [&](int w) { return bar0 + 8 * w; };
  auto level_done = [&](int l) { return bar0 + 8 * (MW + l); };
  auto a_free = [&](int i) { return bar0 + 8 * (MW + RCB_MAX_LEVELS + i); };
  const uint32_t b_full = bar0 + 8 * (MW + RCB_MAX_LEVELS + 2);
  const uint32_t b_empty = b_full + 8;
  const uint32_t acc_full = b_full + 16;
  const uint32_t tmem_slot = b_full + 24;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * C::NBAR);
  float* sbias = reinterpret_cast<float*>(smem + C::OFF_BIAS);

  if (tid == 0) {
    for (int w = 0; w < MW; ++w) mbar_init(gbar(w), 8);
    for (int l = 0; l < RCB_MAX_LEVELS; ++l) mbar_init(level_done(l), MW);
    mbar_init(a_free(0), 1);
    mbar_init(a_free(1), 1);
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (tid < N) sbias[tid] = bias ? __ldg(bias + tid) : 0.f;
  if (warp == MW + 1) tc::tmem_alloc<1>(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == MW) {
    // ---- weight producer: the level's [KL/8][N][8] fp16 block in one bulk copy ----
    const uint32_t bytes = (uint32_t)(KL * N * 2);
    for (int l = 0; l < L; ++l) {
      mbar_wait(b_empty, (uint32_t)((l & 1) ^ 1));
      if (elect_one()) {
        mbar_expect_tx(b_full, bytes);
        bulk_load(base + C::OFF_B, wpack + (long long)l * KL * N, bytes, b_full);
      }
      __syncwarp();
    }
  } else if (warp == MW + 1) {
    // ---- MMA issuer: KL/16 K steps per level as soon as the level's samples and weights are in place ----
    const uint32_t idesc = tc::make_idesc_f16_mn(C::BM, N);
    uint32_t acc = 0;
    for (int l = 0; l < L; ++l) {
      mbar_wait(level_done(l), 0);
      mbar_wait(b_full, (uint32_t)(l & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = base + (l & 1) * C::A_BUF_BYTES;
        const uint32_t b_addr = base + C::OFF_B;
#pragma unroll
        for (int t = 0; t < KL / 16; ++t) {
          const uint64_t adesc = make_desc_interleaved(a_addr + t * 2 * C::A_LBO, C::A_LBO, 128);
          const uint64_t bdesc = make_desc_interleaved(b_addr + t * 2 * N * 16, (uint32_t)N * 16, 128);
          tc::umma_bf16_ss(tmem_base, adesc, bdesc, idesc, acc);
          acc = 1;
        }
        tc::umma_commit<1>(b_empty);
        tc::umma_commit<1>(a_free(l & 1));
      }
      acc = 1;
      __syncwarp();
    }
    if (elect_one()) tc::umma_commit<1>(acc_full);
    __syncwarp();
  } else {
    // ---- gather + resample into the level's A operand ----
    const int ql = lane >> 2, sub = lane & 3;
    const uint32_t bar = gbar(warp);
    const float4* slot = reinterpret_cast<const float4*>(smem + C::OFF_RING + warp * C::WARP_RING + ql * C::SLOT_BYTES);
    const uint32_t slot_addr = base + C::OFF_RING + warp * C::WARP_RING + ql * C::SLOT_BYTES;
    const int m = warp * 8 + ql;  // row of the A tile
    const int q = q0 + m;
    const bool q_ok = q < Q;
    float cx = -1.0e6f, cy = -1.0e6f;
    if (q_ok) {
      cx = __ldg(coords + (long long)(b * 2 + 0) * Q + q);
      cy = __ldg(coords + (long long)(b * 2 + 1) * Q + q);
    }
    const int b0 = (RD * sub) >> 2, nb = ((RD * (sub + 1)) >> 2) - b0;  // output rows [b0, b0 + nb), nb <= NBMAX
    for (int l = 0; l < L; ++l) {
      const int Hl = l == 0 ? pyr.H[0] : l == 1 ? pyr.H[1] : l == 2 ? pyr.H[2] : pyr.H[3];
      const int Wl = l == 0 ? pyr.W[0] : l == 1 ? pyr.W[1] : l == 2 ? pyr.W[2] : pyr.W[3];
      const LevelCoord lc = level_coord<R>(cx, cy, l, Hl, Wl);
      const int ph = lc.xs & 3, py = lc.ys & 3;
      const int nx = (ph + ROWS + 3) >> 2, ny = (py + ROWS + 3) >> 2;
      if (sub == 0) {
        if (q_ok) {
          mbar_expect_tx(bar, (uint32_t)(nx * ny * 64));
          tma_load_3d(slot_addr, &maps.m[l * 4 + (ny - NMIN) * 2 + (nx - NMIN)], bar, (lc.xs >> 2) * 16, lc.ys >> 2,
                      b * Q + q);
        } else {
          mbar_arrive(bar);
        }
      }
      if (l >= 2) mbar_wait(a_free(l & 1), 0);  // the MMAs of level l - 2 have read this buffer
      mbar_wait(bar, (uint32_t)(l & 1));
      if (q_ok) {
        const float fx = lc.fx, fy = lc.fy, gx = 1.0f - lc.fx, gy = 1.0f - lc.fy;
        const bool ragged_w = (Wl & 3) != 0;
        unsigned char* arow = smem + (l & 1) * C::A_BUF_BYTES + m * 16;
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
          for (int i = 0; i < ROWS + 2; ++i) v1[i] = (ph & 1) ? w[i + 1] : w[i];
          float t[ROWS];
#pragma unroll
          for (int i = 0; i < ROWS; ++i) t[i] = (ph & 2) ? v1[i + 2] : v1[i];
          if (ragged_w) {
#pragma unroll
            for (int i = 0; i < ROWS; ++i)
              if (lc.xs + i >= Wl) t[i] = 0.f;
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
          for (int i = 0; i < (KL - RD * RP) / 2; ++i) a_store(RD * RP * 2 + 4 * i, 0u);
        }
      }
      fence_proxy_async_smem();  // A writes -> visible to the tensor core's reads
      __syncwarp();              // and every lane is done with the slots before the next gather lands in them
      if (lane == 0) mbar_arrive(level_done(l));
    }

    // ---- epilogue: lane = query, registers = output channels ----
    const int quarter = warp & 3, cblk = warp >> 2;
    const int qe = q0 + quarter * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* o = out + (long long)b * N * Q + qe;
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      const int n0 = cblk * 64 + cb * 32;
      if (n0 >= N) break;
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + n0, v);
      if (qe < Q) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (n0 + i < N) {
            float x = v[i] + sbias[n0 + i];
            if (relu) x = fmaxf(x, 0.f);
            o[(long long)(n0 + i) * Q] = x;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MW + 1) {
    tc_fence_after();
    tc::tmem_dealloc<1>(tmem_base, 256);
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
  dim3 grid((Q + C::BM - 1) / C::BM, plan.B);
  lookup_conv_kernel<R><<<grid, C::THREADS, C::SMEM_ALLOC, s>>>(plan.maps, pd, coords, wpack, bias, out, Q,
                                                                 plan.lay.levels, cout, relu);
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
