// Lookup fused with its only consumer: the first layer of the motion encoder,
//   cor = relu(convc1(corr))            reference core/update.py:136,154 (small) and 182,202 (basic),
// a 1x1 convolution over the L*(2r+1)^2 correlation channels of CorrBlock.__call__ (core/corr.py:56-94).
// The 73 MB correlation tensor of every GRU iteration is never written to or re-read from HBM: the window
// samples of 128 queries go straight into the A operand of a tensor-core GEMM
//   out[b, n, q] = act( sum_k corr[b, k, q] * Wt[n, k] + bias[n] ),   n < Cout <= 256.
//
// The GEMM is issued transposed, out^T[n, q] = W[n, k] * S[q, k]^T, so that the WEIGHTS are the tensor-memory (A)
// operand: they are loaded into tensor memory once per CTA and stay there for the whole persistent kernel, and the
// shared memory a streamed weight operand would need goes to the gather ring instead.
//   CTA     = persistent, one per SM; a tile = 64 consecutive queries of one batch element, all levels; 672 threads:
//   warps 0-15 gather + resample exactly like lookup_tma_kernel (one TMA box per (query, level), 4 lanes per query),
//              but the samples are rounded to fp16 and stored into the level's shared-memory B operand
//              S[64 queries][KL] instead of global memory.  The 8 queries of a group form one chunk stream
//              (tile, level) over three slot sets; two warps alternate on it, and whoever finishes chunk c issues the
//              gather of chunk c + 3, so three gathers per group (~190 KB per SM) stay in flight across levels and
//              tiles.  One mbarrier per level tells the MMA warp that S is complete.
//   warps 16-19 epilogue: tcgen05.ld (lane = output channel, registers = 32 consecutive queries), bias, ReLU, a
//              SWIZZLE_64B staging tile (32 channels x 16 queries, conflict-free 16-byte stores) and one TMA store per
//              tile of out[b, n0..n0+31, q..q+15], which also clips the ragged edges; direct stores when the row
//              pitch is not a multiple of 16 bytes.
//   warp 20    tcgen05.mma.kind::f16 issuer (A from tensor memory, M = 128 channels per block, N = 64 queries,
//              K = 16), fp32 accumulator in tensor memory; the MMAs of level l run while the other warps resample
//              level l + 1.  Tensor memory: Cout/128 blocks x (K/2 weight columns + 64 accumulator columns) <= 512.
//   All warps of a lane quarter share the one-time copy of the packed weights into tensor memory (tcgen05.st).
// S uses the un-swizzled K-major core-matrix layout ([K/8][64 rows][8 halfs]): the K extent of a level (96 or 64)
// is then free of the 64-element swizzle atom and the 8 queries of a warp write 128 contiguous bytes per word.
// K layout (private to this kernel; pack_convc1_kernel permutes the weights to match): level l owns entries
// [l*KL, (l+1)*KL); inside a level entry b*RP + a holds window sample (dx = a - r, dy = b - r), i.e. reference channel
// l*(2r+1)^2 + a*(2r+1) + b; RP = 2r+2 pads a window row to whole 32-bit words, KL rounds the level to a multiple
// of 16.  Padding entries are written as zeros on the S side and are zero in the packed weights.
// Arithmetic: fp16 operands (11-bit significands, the class of the TF32 convolution cuDNN runs for the reference
// by default), fp32 accumulation.  fp32 pyramids only.
#include <cstdlib>
#include <cstring>

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
  static constexpr int BQ = 64;                            // queries per tile = N of the MMA
  static constexpr int GROUPS = BQ / 8;                    // query groups of 8
  static constexpr int MATH_WARPS = 2 * GROUPS;            // two warps alternate on the chunks of a group
  static constexpr int EPI_WARPS = 4;                      // one per tensor-memory lane quarter
  static constexpr int THREADS = 32 * (MATH_WARPS + EPI_WARPS + 1);
  static constexpr int NSLOT = 3;                          // gathers in flight per query group
  static constexpr int SLOT_BYTES = G::SLOT_BYTES;
  static constexpr int WARP_RING = 8 * SLOT_BYTES;
  static constexpr int S_LBO = BQ * 16;                    // bytes between K chunks of 8: [K/8][64 rows][16 B]
  static constexpr int S_BUF_BYTES = (KL / 8) * S_LBO;     // one level
  static constexpr int OFF_RING = 2 * S_BUF_BYTES;
  static constexpr int STAGE_BYTES = 32 * 64;              // epilogue staging per warp: 32 channels x 16 queries
  static constexpr int OFF_STAGE = OFF_RING + GROUPS * NSLOT * WARP_RING;
  static constexpr int OFF_BAR = OFF_STAGE + EPI_WARPS * STAGE_BYTES;
  // barriers: gather[NSLOT per group], level_done[4], s_free[2], acc_full, d_free, w_full, then the tensor-memory slot
  static constexpr int NBAR = GROUPS * NSLOT + RCB_MAX_LEVELS + 5;
  static constexpr int SMEM_BYTES = OFF_BAR + 8 * NBAR + 16;
  static constexpr int SMEM_ALLOC = SMEM_BYTES + 128;      // the base is rounded up to 128 bytes (TMA destinations)
};

// K-major operand without swizzle: 8-row x 16-byte core matrices; lbo = bytes between the two K chunks of one
// MMA, sbo = bytes between 8-row groups
RCB_DEVINL uint64_t make_desc_interleaved(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46);
}

template <int R, bool PROF>
__global__ void __launch_bounds__(Cfg<R>::THREADS, 1)
lookup_conv_kernel(const __grid_constant__ LookupMaps maps, const __grid_constant__ CUtensorMap omap, PyramidDev pyr,
                   const float* __restrict__ coords, const __half* __restrict__ wpack, const float* __restrict__ bias,
                   float* __restrict__ out, int Q, int L, int N, int relu, int tiles_q, int ntiles,
                   unsigned tiles_q_magic, int tmem_cols, int use_tma_store, int dbg_, unsigned long long* prof_) {
  using C = Cfg<R>;
  using G = typename C::G;
  constexpr int RD = C::RD, RP = C::RP, KL = C::KL, ROWS = G::ROWS, NMIN = G::NMIN, NMAX = G::NMAX;
  constexpr int NBMAX = G::NBMAX, MW = C::MATH_WARPS, NG = C::GROUPS, NS = C::NSLOT, BQ = C::BQ;
  // the timing switches and cycle counters exist only in the PROF instantiation (tools/time_lookup_conv.py)
  const int dbg = PROF ? dbg_ : 0;
  unsigned long long* const prof = PROF ? prof_ : nullptr;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  unsigned char* smem = smem_raw + (base - raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = base + C::OFF_BAR;
  auto gbar = [&](int i) { return bar0 + 8 * i; };
  auto level_done = [&](int l) { return bar0 + 8 * (NG * NS + l); };
  auto s_free = [&](int i) { return bar0 + 8 * (NG * NS + RCB_MAX_LEVELS + i); };
  const uint32_t acc_full = bar0 + 8 * (NG * NS + RCB_MAX_LEVELS + 2);
  const uint32_t d_free = acc_full + 8;
  const uint32_t w_full = acc_full + 16;
  const uint32_t tmem_slot = bar0 + 8 * C::NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * C::NBAR);
  constexpr int EW0 = MW, TW = MW + C::EPI_WARPS;  // first epilogue warp, MMA warp

  if (tid == 0) {
    for (int i = 0; i < NG * NS; ++i) mbar_init(gbar(i), 8);
    for (int l = 0; l < RCB_MAX_LEVELS; ++l) mbar_init(level_done(l), NG);
    mbar_init(s_free(0), 1);
    mbar_init(s_free(1), 1);
    mbar_init(acc_full, 1);
    mbar_init(d_free, C::EPI_WARPS);
    mbar_init(w_full, MW + C::EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == TW) tc::tmem_alloc<1>(tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int tile0 = blockIdx.x, tstep = gridDim.x;
  // tile -> (batch element, first query) without integer division: tiles_q_magic = 2^32 / tiles_q + 1
  auto tile_batch = [&](int tile) { return (int)__umulhi((unsigned)tile, tiles_q_magic); };
  auto tile_q0 = [&](int tile) { return (tile - tile_batch(tile) * tiles_q) * BQ; };
  const int nmb = (N + 127) >> 7;        // 128-channel blocks
  const int KW = (L * KL) >> 1;          // 32-bit weight columns per block
  const uint32_t d_col0 = (uint32_t)(nmb * KW);

  // weights -> tensor memory, once per CTA: lane = output channel; the five warps of a lane quarter (four resampling
  // warps and one epilogue warp) share the 32-column chunks of its rows
  auto load_weights = [&]() {
    const int quarter = warp & 3, part = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int nchunk = (KW + 31) >> 5;
    for (int idx = part; idx < nmb * nchunk; idx += (MW + C::EPI_WARPS) / 4) {
      const int mb = idx / nchunk, ch = idx % nchunk;
      // packed as [chunk][k][128 rows][16 B]: the 32 lanes of one load read 512 contiguous bytes
      const uint4* src = reinterpret_cast<const uint4*>(wpack) + (long long)idx * 8 * 128 + quarter * 32 + lane;
      uint32_t r[32];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ch * 32 + 4 * k < KW) v = __ldg(src + k * 128);
        r[4 * k + 0] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
      }
      if (ch * 32 + 32 <= KW) {
        tc::tmem_st32(tmem_base + lane_base + mb * KW + ch * 32, r);
      } else {  // last partial chunk of a row: 16 columns
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
            ::"r"(tmem_base + lane_base + mb * KW + ch * 32), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
              "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
              "r"(r[14]), "r"(r[15])
            : "memory");
      }
    }
    tc::tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_full);
  };

  if (warp == TW) {
    // ---- MMA issuer: per level and channel block KL/16 K steps, as soon as the level's samples are in place ----
    const uint32_t idesc = tc::make_idesc_f16_mn(128, BQ);
    long long w_lv = 0, w_df = 0, w_wf = 0;
    long long* pw = prof ? &w_lv : nullptr;
    const long long clk0 = clock64();
    mbar_wait_t(w_full, 0, prof ? &w_wf : nullptr);
    tc_fence_after();
    int g = 0, i = 0;
    for (int tile = tile0; tile < ntiles; tile += tstep, ++i) {
      if (i >= 1) mbar_wait_t(d_free, (uint32_t)((i - 1) & 1), prof ? &w_df : nullptr);  // previous tile's epilogue has read it
      for (int l = 0; l < L; ++l, ++g) {
        mbar_wait_t(level_done(l), (uint32_t)(i & 1), pw);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t s_addr = base + (g & 1) * C::S_BUF_BYTES;
          for (int mb = 0; mb < nmb; ++mb) {
#pragma unroll
            for (int t = 0; t < KL / 16; ++t) {
              if (dbg & 1) break;  // timing experiment: no MMAs
              const uint64_t sdesc = make_desc_interleaved(s_addr + t * 2 * C::S_LBO, C::S_LBO, 128);
              tc::umma_bf16_ts<1>(tmem_base + d_col0 + mb * BQ, tmem_base + mb * KW + l * (KL / 2) + t * 8, sdesc, idesc,
                                  (l > 0 || t > 0) ? 1u : 0u);
            }
          }
          tc::umma_commit<1>(s_free(g & 1));
          if (l == L - 1) tc::umma_commit<1>(acc_full);
        }
        __syncwarp();
      }
    }
    if (prof && lane == 0) {
      atomicAdd(prof + 5, (unsigned long long)(clock64() - clk0));
      atomicAdd(prof + 6, (unsigned long long)w_lv);
      atomicAdd(prof + 7, (unsigned long long)w_df);
      atomicAdd(prof + 8, (unsigned long long)w_wf);
    }
  } else if (warp >= EW0) {
    // ---- weights -> tensor memory (once), then the epilogue of every tile ----
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    long long w_acc = 0;
    const long long clk0 = clock64();
    load_weights();
    const long long clk1 = clock64();
    int i = 0;
    for (int tile = tile0; tile < ntiles; tile += tstep, ++i) {
      const int bb = tile_batch(tile), qt = tile_q0(tile);
      mbar_wait_t(acc_full, (uint32_t)(i & 1), prof ? &w_acc : nullptr);
      tc_fence_after();
      for (int mb = 0; mb < nmb; ++mb) {
        const int n = mb * 128 + quarter * 32 + lane;  // this lane's output channel
        const float bv = (bias && n < N) ? __ldg(bias + n) : 0.f;
        float* o = out + ((long long)bb * N + n) * Q + qt;
#pragma unroll
        for (int c = 0; c < BQ / 32; ++c) {
          float v[32];
          tc::tmem_ld32(tmem_base + lane_base + d_col0 + mb * BQ + c * 32, v);
          if (mb == nmb - 1 && c == BQ / 32 - 1) {  // the accumulator may be overwritten by the next tile's MMAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_free);
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            v[k] += bv;
            if (relu) v[k] = fmaxf(v[k], 0.f);
          }
          const int q = qt + c * 32;
          if (dbg & 8) continue;
          if (use_tma_store) {
            // 32 channels x 16 queries at a time through a SWIZZLE_64B staging tile and one TMA store, which also
            // clips queries >= Q and channels >= N
            unsigned char* stage = smem + C::OFF_STAGE + (warp - EW0) * C::STAGE_BYTES;
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              if (lane == 0) tma_store_wait_read<0>();  // the previous store has read the staging tile
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_float4(v[hq * 16 + 4 * j], v[hq * 16 + 4 * j + 1], v[hq * 16 + 4 * j + 2], v[hq * 16 + 4 * j + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0 && mb * 128 + quarter * 32 < N) {
                tma_store_3d(&omap, base + C::OFF_STAGE + (warp - EW0) * C::STAGE_BYTES, q + hq * 16,
                             mb * 128 + quarter * 32, bb);
                tma_store_commit();
              }
            }
          } else if (n < N) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (q + k < Q) o[c * 32 + k] = v[k];
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
    if (prof && lane == 0) {
      atomicAdd(prof + 9, (unsigned long long)(clock64() - clk0));
      atomicAdd(prof + 10, (unsigned long long)w_acc);
      atomicAdd(prof + 12, (unsigned long long)(clk1 - clk0));
    }
  } else {
    // ---- gather + resample into the level's S operand ----
    // A group of 8 queries has one chunk stream (tile, level) and three slot sets: chunk c lives in slot set c % 3.
    // Two warps alternate on the stream (warp parity = chunk parity); whoever finishes chunk c issues the gather of
    // chunk c + 3 into the slot set it has just emptied, so three gathers per group stay in flight.
    const int ql = lane >> 2, sub = lane & 3;
    const int grp = warp >> 1, par = warp & 1;
    const int m = grp * 8 + ql;  // query row inside the tile
    const int b0 = (RD * sub) >> 2, nb = ((RD * (sub + 1)) >> 2) - b0;  // output rows [b0, b0 + nb), nb <= NBMAX
    const int my_tiles = tile0 < ntiles ? (ntiles - tile0 + tstep - 1) / tstep : 0;
    const int nch = my_tiles * L;  // chunks of the group's stream; chunk c is also the CTA's global level counter

    struct Pending {
      LevelCoord lc;
      int Hl, Wl;
      bool ok;
    };
    auto load_coords = [&](int tile, float& cx, float& cy) {
      const int bb = tile_batch(tile), q = tile_q0(tile) + m;
      cx = cy = -1.0e6f;
      if (tile < ntiles && q < Q) {
        cx = __ldg(coords + (long long)(bb * 2 + 0) * Q + q);
        cy = __ldg(coords + (long long)(bb * 2 + 1) * Q + q);
      }
    };
    auto describe = [&](int tile, int l, float cx, float cy) {
      Pending p;
      p.ok = tile_q0(tile) + m < Q;
      p.Hl = l == 0 ? pyr.H[0] : l == 1 ? pyr.H[1] : l == 2 ? pyr.H[2] : pyr.H[3];
      p.Wl = l == 0 ? pyr.W[0] : l == 1 ? pyr.W[1] : l == 2 ? pyr.W[2] : pyr.W[3];
      p.lc = level_coord<R>(cx, cy, l, p.Hl, p.Wl);
      return p;
    };
    auto issue = [&](int tile, int l, int slot, float cx, float cy) {
      const Pending p = describe(tile, l, cx, cy);
      if (sub == 0) {
        const int bb = tile_batch(tile), q = tile_q0(tile) + m;
        const uint32_t bar = gbar(grp * NS + slot);
        if (p.ok && !(dbg & 2)) {
          const int nx = ((p.lc.xs & 3) + ROWS + 3) >> 2, ny = ((p.lc.ys & 3) + ROWS + 3) >> 2;
          mbar_expect_tx(bar, (uint32_t)(nx * ny * 64));
          tma_load_3d(base + C::OFF_RING + (grp * NS + slot) * C::WARP_RING + ql * C::SLOT_BYTES,
                      &maps.m[l * 4 + (ny - NMIN) * 2 + (nx - NMIN)], bar, (p.lc.xs >> 2) * 16, p.lc.ys >> 2,
                      bb * Q + q);
        } else {
          mbar_arrive(bar);
        }
      }
    };
    // resamples the gathered windows in slot set `slot` and writes them into rows of S buffer `sbuf`
    auto consume = [&](const Pending& cur, int slot, int sbuf) {
      if (!cur.ok || (dbg & 4)) return;
      const float4* win =
          reinterpret_cast<const float4*>(smem + C::OFF_RING + (grp * NS + slot) * C::WARP_RING + ql * C::SLOT_BYTES);
      const LevelCoord lc = cur.lc;
      const int Hl = cur.Hl, Wl = cur.Wl;
      const int ph = lc.xs & 3, py = lc.ys & 3;
      const int nx = (ph + ROWS + 3) >> 2;
      const float fx = lc.fx, fy = lc.fy, gx = 1.0f - lc.fx, gy = 1.0f - lc.fy;
      const bool ragged_w = (Wl & 3) != 0;
      unsigned char* srow = smem + sbuf * C::S_BUF_BYTES + m * 16;
      auto s_store = [&](int kbyte, uint32_t v) {  // kbyte: byte offset inside the level's K row, multiple of 4
        *reinterpret_cast<uint32_t*>(srow + (kbyte >> 4) * C::S_LBO + (kbyte & 15)) = v;
      };
      float hp[RD];
#pragma unroll
      for (int jj = 0; jj <= NBMAX; ++jj) {
        if (jj > nb) break;
        const int j = b0 + jj;  // window row
        const int ya = py + j;  // row inside the fetched box
        const bool row_ok = lc.ys + j < Hl;
        const float4* rowp = win + ((ya >> 2) * nx) * 4 + (ya & 3);
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
            s_store(kb + 4 * a2, *reinterpret_cast<const uint32_t*>(&hv));
          }
        }
#pragma unroll
        for (int a = 0; a < RD; ++a) hp[a] = hh[a];
      }
      if (sub == 3) {  // the level's trailing K padding must be finite: zeros
#pragma unroll
        for (int k = 0; k < (KL - RD * RP) / 2; ++k) s_store(RD * RP * 2 + 4 * k, 0u);
      }
    };

    // Cursors over the group's stream.  The consume cursor visits this warp's chunks (c = par, par + 2, ...); the
    // issue cursor runs three chunks ahead of it.  Both keep the coordinates of their tile and prefetch the next.
    long long w_g = 0, w_s = 0, t_c = 0, t_f = 0, t_i = 0;
    const long long clk0 = clock64();
    if (par == 0) {  // prologue: chunks 0, 1, 2 of the stream
      int pt = tile0, pl = 0;
      float px, py_;
      load_coords(pt, px, py_);
      for (int c = 0; c < NS && c < nch; ++c) {
        issue(pt, pl, c, px, py_);
        if (++pl == L) { pl = 0; pt += tstep; load_coords(pt, px, py_); }
      }
    }
    const long long clk_a = clock64();
    load_weights();  // while the first gathers are in flight
    const long long clk_b = clock64();
    int ct = tile0, cl = par, it = tile0, il = par + NS;  // (tile, level) of chunk c and of chunk c + 3
    while (cl >= L) { cl -= L; ct += tstep; }
    while (il >= L) { il -= L; it += tstep; }
    float ccx, ccy, cnx, cny, icx, icy, inx, iny;
    load_coords(ct, ccx, ccy);
    load_coords(ct + tstep, cnx, cny);
    load_coords(it, icx, icy);
    load_coords(it + tstep, inx, iny);
    const long long clk_loop = clock64();
    for (int c = par; c < nch; c += 2) {
      const int slot = c % NS;
      const Pending cur = describe(ct, cl, ccx, ccy);
      if (c >= 2) mbar_wait_t(s_free(c & 1), (uint32_t)(((c >> 1) - 1) & 1), prof ? &w_s : nullptr);  // MMAs two levels back
      mbar_wait_t(gbar(grp * NS + slot), (uint32_t)((c / NS) & 1), prof ? &w_g : nullptr);
      const long long tc0 = prof ? clock64() : 0;
      consume(cur, slot, c & 1);
      const long long tc1 = prof ? clock64() : 0;
      if (prof) t_c += tc1 - tc0;
      fence_proxy_async_smem();  // S writes -> visible to the tensor core's reads
      __syncwarp();              // and every lane is done with the slots before the next gather lands in them
      if (lane == 0) mbar_arrive(level_done(cl));
      const long long tc2 = prof ? clock64() : 0;
      if (c + NS < nch) issue(it, il, slot, icx, icy);
      if (prof) { t_f += tc2 - tc1; t_i += clock64() - tc2; }
      // advance both cursors by two chunks
      cl += 2;
      if (cl >= L) { cl -= L; ct += tstep; ccx = cnx; ccy = cny; load_coords(ct + tstep, cnx, cny); }
      il += 2;
      if (il >= L) { il -= L; it += tstep; icx = inx; icy = iny; load_coords(it + tstep, inx, iny); }
    }
    if (prof && lane == 0) {
      atomicAdd(prof + 0, (unsigned long long)(clock64() - clk0));
      atomicAdd(prof + 1, (unsigned long long)w_g);
      atomicAdd(prof + 2, (unsigned long long)w_s);
      atomicAdd(prof + 3, (unsigned long long)t_c);
      atomicAdd(prof + 4, (unsigned long long)nch);
      atomicAdd(prof + 13, (unsigned long long)t_f);
      atomicAdd(prof + 14, (unsigned long long)t_i);
      atomicAdd(prof + 15, (unsigned long long)(clk_loop - clk0));
      atomicAdd(prof + 16, (unsigned long long)(clk_a - clk0));
      atomicAdd(prof + 17, (unsigned long long)(clk_b - clk_a));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TW) {
    tc_fence_after();
    tc::tmem_dealloc<1>(tmem_base, (uint32_t)tmem_cols);
  }
}

// weight[n][l*RD*RD + a*RD + b] (fp32, the Conv2d weight of convc1 viewed [Cout, Cin]) -> fp16 in the order the
// kernel copies it into tensor memory: [128-channel block][32-column chunk][k = 0..7][row 0..127][8 halfs], where
// the 8 halfs are K entries ch*64 + k*8 + 0..7 of channel block*128 + row and K entry l*KL + b*RP + a is the window
// sample (a, b) of level l; rows >= Cout, K padding and the tail of a last partial chunk are zero
__global__ void __launch_bounds__(256)
pack_convc1_kernel(const float* __restrict__ w, __half* __restrict__ wp, int cout, int nmb, int nchunk, int L, int RD,
                   int RP, int KL) {
  const int K = L * KL;
  const long long n_el = (long long)nmb * nchunk * 8 * 128 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7), rr = (int)((i >> 3) & 127), k8 = (int)((i >> 10) & 7);
    const int chunk = (int)(i >> 13);
    const int mb = chunk / nchunk, ch = chunk % nchunk;
    const int n = mb * 128 + rr, kk = ch * 64 + k8 * 8 + e;
    const int l = kk / KL, r = kk % KL;
    const int bb = r / RP, a = r % RP;
    float v = 0.f;
    if (n < cout && kk < K && bb < RD && a < RD) v = __ldg(w + (long long)n * (L * RD * RD) + l * RD * RD + a * RD + bb);
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
  const int Q = plan.H * plan.W, L = plan.lay.levels;
  cudaError_t e = cudaFuncSetAttribute(lookup_conv_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_ALLOC);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(lookup_conv_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_ALLOC);
  if (e != cudaSuccess) return (int)e;
  const int nmb = (cout + 127) / 128;
  const int need = nmb * (L * C::KL / 2 + C::BQ);  // weight + accumulator columns
  int tmem_cols = 32;
  while (tmem_cols < need) tmem_cols *= 2;
  if (tmem_cols > 512) return RCB_ERR_UNSUPPORTED;
  static const int dbg = debug_env_int("RCB_LCONV_DEBUG", 0);  // timing experiments (RCB_DEBUG builds only)
  // output as a [B][N][Q] tensor for the epilogue's TMA stores (needs 16-byte row pitch)
  CUtensorMap omap;
  int use_tma_store = (Q % 4 == 0) && encode_fn() != nullptr;
  if (use_tma_store) {
    cuuint64_t dims[3] = {(cuuint64_t)Q, (cuuint64_t)cout, (cuuint64_t)plan.B};
    cuuint64_t str[2] = {(cuuint64_t)Q * 4, (cuuint64_t)Q * 4 * cout};
    cuuint32_t box[3] = {16, 32, 1};
    if (!encode(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) use_tma_store = 0;
  }
  if (!use_tma_store) memset(&omap, 0, sizeof(omap));
  // RCB_DEBUG builds only: device uint64 counters supplied by tools/time_lookup_conv.py
  unsigned long long* prof = debug_env_ptr("RCB_LCONV_PROF_PTR");
  const int tiles_q = (Q + C::BQ - 1) / C::BQ, ntiles = tiles_q * plan.B;
  const unsigned magic = (unsigned)(0x100000000ull / (unsigned)tiles_q) + 1u;  // exact for tile * tiles_q < 2^32
  if ((unsigned long long)ntiles * tiles_q >= 0x100000000ull) return RCB_ERR_UNSUPPORTED;
  const int grid = ntiles < kNumSMs ? ntiles : kNumSMs;  // persistent: one CTA per SM walks tiles grid apart
  if (dbg || prof)
    lookup_conv_kernel<R, true><<<grid, C::THREADS, C::SMEM_ALLOC, s>>>(plan.maps, omap, pd, coords, wpack, bias, out, Q,
                                                                       L, cout, relu, tiles_q, ntiles, magic, tmem_cols,
                                                                       use_tma_store, dbg, prof);
  else
    lookup_conv_kernel<R, false><<<grid, C::THREADS, C::SMEM_ALLOC, s>>>(plan.maps, omap, pd, coords, wpack, bias, out,
                                                                        Q, L, cout, relu, tiles_q, ntiles, magic, tmem_cols,
                                                                        use_tma_store, 0, nullptr);
  return launch_status();
}

}  // namespace lconv

size_t convc1_pack_bytes(int cout, int levels, int radius) {
  int rd, rp, kl;
  if (!lconv::geometry(radius, &rd, &rp, &kl) || cout < 16 || cout > 256 || cout % 16 || levels < 2 ||
      levels > RCB_MAX_LEVELS)
    return 0;
  const int nmb = (cout + 127) / 128, nchunk = (levels * kl / 2 + 31) / 32;  // 32-column chunks of a row
  return (size_t)nmb * nchunk * 8 * 128 * 16;
}

int launch_convc1_pack(const float* weight, void* wpack, int cout, int levels, int radius, cudaStream_t s) {
  int rd, rp, kl;
  if (!weight || !wpack) return RCB_ERR_INVALID_ARGUMENT;
  if (convc1_pack_bytes(cout, levels, radius) == 0 || !lconv::geometry(radius, &rd, &rp, &kl)) return RCB_ERR_UNSUPPORTED;
  const int nmb = (cout + 127) / 128, nchunk = (levels * kl / 2 + 31) / 32;
  const long long n_el = (long long)nmb * nchunk * 8 * 128 * 8;
  lconv::pack_convc1_kernel<<<(int)((n_el + 255) / 256), 256, 0, s>>>(weight, static_cast<__half*>(wpack), cout, nmb,
                                                                     nchunk, levels, rd, rp, kl);
  return launch_status();
}

int launch_lookup_convc1(const void* plan_, const float* coords, const void* wpack, const float* bias, float* out,
                         int cout, int relu, cudaStream_t s) {
  const LookupPlan* plan = static_cast<const LookupPlan*>(plan_);
  if (!plan || plan->magic != kPlanMagic || !coords || !wpack || !out) return RCB_ERR_INVALID_ARGUMENT;
  if (plan->lay.dtype != RCB_F32) return RCB_ERR_UNSUPPORTED;
  int rd, rp, kl;
  if (!lconv::geometry(plan->radius, &rd, &rp, &kl) || cout < 16 || cout > 256 || cout % 16) return RCB_ERR_UNSUPPORTED;
  if (plan->lay.levels < 2) return RCB_ERR_UNSUPPORTED;  // the per-level barriers need two levels to alternate
  if (((uintptr_t)wpack & 15) != 0 || ((uintptr_t)out & 15) != 0) return RCB_ERR_INVALID_ARGUMENT;
  const PyramidDev pd = make_pyramid_dev(plan->ptr, plan->lay);
  const __half* wp = static_cast<const __half*>(wpack);
  if (plan->radius == 3) return lconv::launch_r<3>(*plan, pd, coords, wp, bias, out, cout, relu, s);
  return lconv::launch_r<4>(*plan, pd, coords, wp, bias, out, cout, relu, s);
}

}  // namespace rcb
