// K1, tensor-core modes: all-pairs correlation volume + fused 4-level pooled pyramid in ONE kernel.
//
// Replaces CorrBlock.corr (reference core/corr.py:96-127: matmul + / sqrt(C)) and the avg_pool2d loop of
// CorrBlock.__init__ (core/corr.py:52-54).  See DESIGN.md "K1" for the roofline discussion: with an fp32
// pyramid the kernel is bound by the 2.2 GB of stores, and in the 3-product fp32-parity mode the tensor pipe
// needs about as long, so the design goal is to keep both busy at the same time.
//
//   pack kernel   fp32 NCHW feature maps -> bf16 K-major [part][B][Q][Kp] (part 0 = hi, part 1 = lo = x - hi),
//                 i.e. a transpose + split.  RCB_BUILD_BF16X3 evaluates  hi*hi + hi*lo + lo*hi  (error ~5e-6 of
//                 max-abs); RCB_BUILD_BF16 uses the hi parts only.
//   main kernel   persistent, warp-specialised, 320 threads per CTA, one CTA per SM; template NCTA:
//     NCTA = 1    (default) single-CTA MMAs, M = 128 queries x N = 128 targets.
//     NCTA = 2    the two CTAs of a cluster pair up on ONE tcgen05.mma.cta_group::2 of M = 256 x N = 128: each CTA
//                 keeps its own 128 query rows (A operand, accumulator) and supplies HALF of every B tile, so per
//                 flop the pair moves half the B bytes through L2, the TMA unit and shared memory and the MMAs
//                 issue at full rate (60 vs 83 cycles each).  Measured slower end to end (the epilogue's stores
//                 bound the kernel and the pair couples two SMs' epilogues: 693 vs 643 us at cfg2), so it is
//                 opt-in (RCB_TC_NCTA=2); both forms run the same parity tests.
//     warp 0      TMA producer: B tiles (an 8 x 16 PATCH of target pixels x 64 channels; a 4-D box of the
//                 [part*B, H, W, Kp] tensor, out-of-image rows/cols zero-filled; NCTA = 2: 4 patch rows per CTA)
//                 stream through an mbarrier ring, two boxes per stage.
//     warp 1      MMA issuer (the pair's leader CTA only): kind::f16, K = 16, the A operand comes from TENSOR MEMORY
//                 (TS form), B through SWIZZLE_128B K-major smem descriptors, fp32 accumulators in TMEM.
//     warps 2-9   epilogue, two warps per TMEM lane quarter (32 queries), split by output level: warps 2-5
//                 tcgen05.ld the accumulator (lane = query, 64 columns = 4 patch rows x 16 targets = one band of
//                 four 4x4 pyramid tiles = 256 contiguous bytes per query), scale by 1/sqrt(C) and send level 0
//                 through SWIZZLE_128B staging boxes and 4-D TMA stores of [32 queries][128 B] (the store clips
//                 tile columns / tile rows / queries outside the tensor).  Warps 6-9 read the same accumulator
//                 and form levels 1-3 as 2x2 means in registers -- the patch is 8x8-aligned, so all three pooled
//                 levels are patch-local; floor-mode dropping of odd rows/cols needs no code: a level-k cell
//                 computed from a dropped row/col is itself outside H_k x W_k, i.e. tile padding or a clipped
//                 tile.  Level 1 is TMA-stored, levels 2/3 (6% / 1.5% of the bytes) are written with 32/8-byte
//                 global stores.  Warps 6-9 also place the A operand in tensor memory (global -> registers ->
//                 tcgen05.st) at the start of every unit.
//   work split    a unit = all patches of one (batch, query tile[s]); units go round-robin over the clusters, which
//                 therefore sweep the same batch's patches in step (B tiles are found in L2); the units left over
//                 after the full rounds are cut along the patch sequence so the last round is short, not idle.
//   measured      (tools/time_build.py, tools/probe/tma_store_probe.cu, tools/power_probe.sh) the store path is the
//                 bound: scattered 128-byte rows reach 3.5-5 TB/s through TMA stores, lane-scattered or octet-
//                 coalesced st.global variants of the epilogue were slower, and in the 3-product mode the chip
//                 sits at its 1 kW power cap (~1.55 GHz).
#include <cstdlib>

#include "rcb_common.cuh"
#include "tcgen05_util.cuh"
#include "tma_util.cuh"

namespace rcb {

namespace tc {

constexpr int BM = 128;          // queries per CTA tile (TMEM lanes)
constexpr int PH = 8, PW = 16;   // target patch: 8 rows x 16 cols
constexpr int BN = PH * PW;      // 128 targets per tile (TMEM columns)
constexpr int BK = 64;           // bf16 channels per smem stage row (128 bytes, SWIZZLE_128B)
constexpr int UMMA_K = 16;
constexpr int MAX_KB = 4;        // Kp <= 256
constexpr int MAX_STAGE = 12;    // B ring depth is chosen at launch from the shared memory available
constexpr int MAX_ACC = 4;       // TMEM accumulator buffers (128 columns each); count chosen at launch
constexpr int B_TILE_BYTES = BN * BK * 2;  // 16 KB: one B box of the whole patch (half of it per CTA when NCTA = 2)
constexpr int BOXES_PER_STAGE = 2;         // a ring stage = 2 boxes on one mbarrier: (kb,hi)+(kb,lo) or (kb)+(kb+1)
constexpr int STG_BYTES = 4096;            // one staged store box per warp
constexpr int NUM_EPI_WARPS = 4;           // per epilogue role (level 0 / pooled levels)
constexpr int THREADS = 32 * (2 + 2 * NUM_EPI_WARPS);

// Dynamic shared memory (all tile buffers 1024-byte aligned for the swizzle atoms):
//   [b_off, +nstage * stage bytes)  B ring
//   [stg_off, +8 * nstg * 4 KB)     epilogue staging rings, nstg boxes per warp
//   [bar_off, +1 KB)                mbarriers + TMEM base slot
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int BAR_BYTES = 1024;

struct Params {
  int B, C, H, W, Q;
  int kblocks;      // Kp / 64
  int parts;        // 1 (bf16) or 2 (bf16x3)
  int levels;
  int mtiles;       // ceil(Q / 128)
  int mgroups;      // ceil(mtiles / NCTA): query-tile groups, one tile per CTA of the cluster
  int pcols, prows; // patch grid
  int npatch;       // pcols * prows
  int full_rounds;  // rounds in which every cluster sweeps all patches of one (batch, query-tile group) unit
  int tail_pieces, tail_split, tail_len;  // left-over units: cut into tail_split ranges of tail_len patches
  float scale;      // 1 / sqrt(C)
  int nstage, b_off, stg_off, bar_off;  // shared memory carve-up (bytes)
  int nstg;                             // staging boxes per epilogue warp
  int f16;                              // pyramid stored as fp16 (tiles of 4 rows x 8 columns), else fp32 (4 x 4)
  int nacc, acc_col0;                   // TMEM: accumulator count, first accumulator column
  int Kp;                               // padded channel count
  const uint32_t* a_pack;               // packed bf16 A operand, [part][B][Q][Kp/2] 32-bit words
  unsigned long long* prof;  // debug: per-CTA cycle counters (16 per CTA), null in production
  int debug_skip;   // debug bitmask: 1 skip L0 TMA store issue, 2 skip L1 store, 4 skip L2/L3, 8 skip staging writes,
                    // 16 skip B loads, 32 skip MMAs, 64 skip TMEM loads, 128/256 plain arrives instead of commits
  float* pyr[RCB_MAX_LEVELS];           // (fp16 pyramids: the same pointers, reinterpreted)
  int Hl[RCB_MAX_LEVELS], Wl[RCB_MAX_LEVELS], tx[RCB_MAX_LEVELS];  // level sizes, tiles per tile row
  long long ps[RCB_MAX_LEVELS];
};

// kind::f16 instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, M = 128 * NCTA, N = 128
__host__ __device__ constexpr uint32_t make_idesc(int ncta) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)((BM * ncta) >> 4) << 24);
}

// ---- operand packing -----------------------------------------------------------------------
// in  [2 maps][B][C][Q] fp32 (two separate base pointers);  out [map][part][B][Q][Kp] bf16, zero padded to Kp.
// CTA = 64 channels x 64 queries: coalesced 8-byte loads along q into a padded fp32 tile, then every thread turns
// 8 channels of one query into one 16-byte store per part (8 lanes cover the 128 bytes of a query's k-block).
__global__ void __launch_bounds__(256)
pack_operands_kernel(const float* __restrict__ f1, const float* __restrict__ f2, __nv_bfloat16* __restrict__ out,
                     int B, int C, int Q, int Kp, int parts) {
  __shared__ float tile[64][65];  // odd pitch: the column reads below are at most 2-way bank-conflicted
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int map = blockIdx.z / B, b = blockIdx.z % B;
  const float* in = (map == 0 ? f1 : f2) + (long long)b * C * Q;
  const int q0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const bool even_q = (Q & 1) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = warp * 8 + i, cc = c0 + c, q = q0 + 2 * lane;
    float2 v = make_float2(0.f, 0.f);
    if (cc < C) {
      const float* src = in + (long long)cc * Q + q;
      if (even_q && q + 1 < Q) {
        v = __ldg(reinterpret_cast<const float2*>(src));  // rows start 8-byte aligned when Q is even
      } else {
        if (q < Q) v.x = __ldg(src);
        if (q + 1 < Q) v.y = __ldg(src + 1);
      }
    }
    tile[c][2 * lane] = v.x;
    tile[c][2 * lane + 1] = v.y;
  }
  __syncthreads();
  const long long part_stride = (long long)B * Q * Kp;
  __nv_bfloat16* o = out + (long long)map * parts * part_stride + (long long)b * Q * Kp;
  const int cg = lane & 7;  // group of 8 channels
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ql = warp * 8 + i * 4 + (lane >> 3), q = q0 + ql;
    if (q >= Q) continue;
    __nv_bfloat162 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = tile[8 * cg + 2 * j][ql], x1 = tile[8 * cg + 2 * j + 1][ql];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      hi[j] = __halves2bfloat162(h0, h1);
      lo[j] = __halves2bfloat162(__float2bfloat16_rn(x0 - __bfloat162float(h0)),
                                 __float2bfloat16_rn(x1 - __bfloat162float(h1)));
    }
    __nv_bfloat16* dst = o + (long long)q * Kp + c0 + 8 * cg;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    if (parts > 1) *reinterpret_cast<uint4*>(dst + part_stride) = *reinterpret_cast<const uint4*>(lo);
  }
}

// 8 floats -> 8 fp16 (round to nearest even) in one 16-byte vector
RCB_DEVINL uint4 pack_half8(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  uint4 u;
  __half2 h;
  h = __floats2half2_rn(a0, a1); u.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2half2_rn(a2, a3); u.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2half2_rn(a4, a5); u.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2half2_rn(a6, a7); u.w = *reinterpret_cast<uint32_t*>(&h);
  return u;
}
// element offset (in halfs) of (y, x) in a plane of 4 x 8 fp16 tiles
RCB_DEVINL long long tile_off_h(int y, int x, int tiles_x) {
  return ((long long)((y >> 2) * tiles_x + (x >> 3)) << 5) + ((y & 3) << 3) + (x & 7);
}

// ---- main kernel -------------------------------------------------------------------------------
// One piece of a cluster's work: a patch range of one (batch, query tile) unit.
struct Segment {
  int b, mt, p_begin, p_end;
};

template <int NCTA>
__global__ void __launch_bounds__(THREADS, 1)
build_tc_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_l0,
                const __grid_constant__ CUtensorMap map_l1, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = NCTA == 2 ? cluster_ctarank() : 0u;  // rank inside the MMA pair; rank 0 leads
  const bool leader = crank == 0;
  constexpr int STAGE_BYTES = BOXES_PER_STAGE * B_TILE_BYTES / NCTA;  // per CTA
  constexpr int BOX_BYTES = B_TILE_BYTES / NCTA;

  // barriers (8 bytes each).  With NCTA = 2 the barriers the MMA issuer waits on (a_full, b_full, acc_empty) are
  // the LEADER's copies and collect arrivals from both CTAs; the ones it signals (a_empty, b_empty, acc_full)
  // exist in both CTAs and are signalled by one multicast commit.
  const uint32_t bar0 = smem_base + p.bar_off;
  const int NSTAGE = p.nstage;
  const uint32_t a_full = bar0, a_empty = bar0 + 8;
  auto b_full = [&](int s) { return bar0 + 16 + 8 * s; };
  auto b_empty = [&](int s) { return bar0 + 16 + 8 * MAX_STAGE + 8 * s; };
  auto acc_full = [&](int s) { return bar0 + 16 + 16 * MAX_STAGE + 8 * s; };
  auto acc_empty = [&](int s) { return bar0 + 16 + 16 * MAX_STAGE + 8 * MAX_ACC + 8 * s; };
  const uint32_t tmem_slot = bar0 + 16 + 16 * MAX_STAGE + 16 * MAX_ACC;
  const int NACC = p.nacc;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 16 + 16 * MAX_STAGE + 16 * MAX_ACC);

  if (threadIdx.x == 0) {
    mbar_init(a_full, NUM_EPI_WARPS * NCTA);  // the pooled-levels warps of every CTA of the pair load A
    mbar_init(a_empty, 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(b_full(s), 1);  // the leader's producer arrives with the byte count of BOTH CTAs' boxes
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 2 * NUM_EPI_WARPS * NCTA);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<NCTA>(tmem_slot, 512);
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // This cluster's work: segment i < full_rounds is the whole patch sweep of unit i * nclus + cid (unit = batch,
  // query-tile group); the units left over after the full rounds are cut into `tail_split` patch ranges each so that
  // every cluster gets at most one piece.  Clusters of one round sweep the same batch's patches in step, so a B tile
  // is fetched from DRAM once and found in L2 by the other clusters.
  const int nclus = gridDim.x / NCTA, cid = blockIdx.x / NCTA;
  const int nseg = p.full_rounds + (cid < p.tail_pieces ? 1 : 0);
  auto segment = [&](int i) {
    Segment sg;
    int u;
    if (i < p.full_rounds) {
      u = i * nclus + cid;
      sg.p_begin = 0;
      sg.p_end = p.npatch;
    } else {
      u = p.full_rounds * nclus + cid / p.tail_split;
      sg.p_begin = (cid % p.tail_split) * p.tail_len;
      sg.p_end = min(p.npatch, sg.p_begin + p.tail_len);
    }
    sg.b = u / p.mgroups;
    sg.mt = (u % p.mgroups) * NCTA + (int)crank;  // may be >= mtiles for the pair's second CTA: all-padding tile
    return sg;
  };

  if (warp == 0) {
    // =============================== TMA producer (whole warp runs the loop, one elected lane issues) ====
    int s = 0;           // B ring slot
    uint32_t ph = 0;     // its phase
    long long w_b = 0;
    long long* pw_b = p.prof ? &w_b : nullptr;
    const long long clk0 = clock64();
    for (int si = 0; si < nseg; ++si) {
      const Segment sg = segment(si);
      for (int pi = sg.p_begin; pi < sg.p_end; ++pi) {
        const int py = pi / p.pcols, px = pi % p.pcols;
        const int nb = p.kblocks * p.parts;  // boxes of this tile in (kb, part) order
        for (int j = 0; j < nb; j += BOXES_PER_STAGE) {
          mbar_wait_t(b_empty(s), ph ^ 1, pw_b);
          if (elect_one()) {
            const int nbox = min(BOXES_PER_STAGE, nb - j);
            if (p.debug_skip & 16) {
              if (leader) mbar_arrive(b_full(s));
            } else {
              // The peer's boxes are counted on the leader's barrier too; it cannot run a ring lap ahead because
              // its b_empty is only signalled after the leader's MMAs consumed this phase.
              if (leader) mbar_expect_tx(b_full(s), (uint32_t)(nbox * B_TILE_BYTES));
              for (int i = 0; i < nbox; ++i) {
                const int jj = j + i;
                const int kb = p.parts == 2 ? jj >> 1 : jj, part = p.parts == 2 ? jj & 1 : 0;
                tma_load_b<NCTA>(smem_base + p.b_off + s * STAGE_BYTES + i * BOX_BYTES, &map_b, b_full(s), kb * BK,
                                 px * PW, py * PH + (int)crank * (PH / NCTA), part * p.B + sg.b);
              }
            }
          }
          __syncwarp();
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
    if (p.prof && lane == 0) {
      p.prof[blockIdx.x * 16 + 0] = clock64() - clk0;
      p.prof[blockIdx.x * 16 + 1] = w_b;
      p.prof[blockIdx.x * 16 + 2] = 0;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA; whole warp runs the loop, one lane issues) ===
    if (leader) {
      constexpr uint32_t idesc = make_idesc(NCTA);
      const uint64_t desc_base = make_smem_desc(0);  // everything but the start address
      int s = 0, buf = 0;
      uint32_t ph = 0, aph = 0, nunit = 0;
      long long w_bf = 0, w_acc = 0, w_af = 0;
      long long* pw_bf = p.prof ? &w_bf : nullptr;
      long long* pw_acc = p.prof ? &w_acc : nullptr;
      long long* pw_af = p.prof ? &w_af : nullptr;
      const long long clk0 = clock64();
      for (int si = 0; si < nseg; ++si, ++nunit) {
        const Segment sg = segment(si);
        mbar_wait_t(a_full, nunit & 1, pw_af);
        tc_fence_after();
        for (int pi = sg.p_begin; pi < sg.p_end; ++pi) {
          mbar_wait_t(acc_empty(buf), aph ^ 1, pw_acc);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + p.acc_col0 + buf * BN;
          uint32_t acc = 0;
          const int nb = p.kblocks * p.parts;
          for (int j = 0; j < nb; j += BOXES_PER_STAGE) {
            mbar_wait_t(b_full(s), ph, pw_bf);
            tc_fence_after();
            if (elect_one()) {
              const int nbox = min(BOXES_PER_STAGE, nb - j);
              if (!(p.debug_skip & 32)) {
                for (int i = 0; i < nbox; ++i) {
                  const int jj = j + i;
                  const int kb = p.parts == 2 ? jj >> 1 : jj, part = p.parts == 2 ? jj & 1 : 0;
                  const uint64_t bdesc =
                      desc_base | (uint64_t)(((smem_base + p.b_off + s * STAGE_BYTES + i * BOX_BYTES) >> 4) & 0x3FFF);
                  const uint32_t ta_hi = tmem_base + kb * (BK / 2);  // A columns of this k-block
                  const uint32_t ta_lo = tmem_base + (p.Kp >> 1) + kb * (BK / 2);
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k) {  // +32 bytes (smem) / +8 columns (TMEM) per K step
                    umma_bf16_ts<NCTA>(d_tmem, ta_hi + 8 * k, bdesc + 2 * k, idesc, acc);
                    acc = 1;
                  }
                  if (part == 0 && p.parts > 1) {  // B_hi also meets A_lo
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                      umma_bf16_ts<NCTA>(d_tmem, ta_lo + 8 * k, bdesc + 2 * k, idesc, 1u);
                  }
                }
              }
              if (p.debug_skip & 128) {
                mbar_arrive(b_empty(s));
                if (NCTA == 2) mbar_arrive_cluster(b_empty(s), 1);
              } else {
                umma_commit<NCTA>(b_empty(s));  // frees the stage (in both CTAs) once these MMAs have read it
              }
            }
            acc = 1;
            __syncwarp();
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
          if (elect_one()) {
            if (p.debug_skip & 256) {
              mbar_arrive(acc_full(buf));
              if (NCTA == 2) mbar_arrive_cluster(acc_full(buf), 1);
            } else {
              umma_commit<NCTA>(acc_full(buf));
            }
          }
          __syncwarp();
          if (++buf == NACC) { buf = 0; aph ^= 1; }
        }
        if (elect_one()) umma_commit<NCTA>(a_empty);
        __syncwarp();
      }
      if (p.prof && lane == 0) {
        p.prof[blockIdx.x * 16 + 3] = clock64() - clk0;
        p.prof[blockIdx.x * 16 + 4] = w_bf;
        p.prof[blockIdx.x * 16 + 5] = w_acc;
        p.prof[blockIdx.x * 16 + 6] = w_af;
      }
    }
  } else {
    // =============================== epilogue (+ A loading) ===============================
    // Two warps per TMEM lane quarter (warp w works on quarter w % 4 = 32 queries), split by OUTPUT LEVEL so that
    // every global write keeps its full width (128-byte rows; 64-byte rows and 8-byte stores measured 2-4x more
    // expensive per byte):
    //   warps 2-5  level 0: tcgen05.ld, scale, four staged boxes [32 queries][2 tiles = 128 B] per tile
    //   warps 6-9  levels 1-3: tcgen05.ld, scale, 2x2 means in registers, one staged level-1 box and the few
    //              level-2/3 values; they also place the A operand in tensor memory at the start of every unit.
    const bool pooled_role = warp >= 2 + NUM_EPI_WARPS;
    const int ew = warp - 2;             // staging ring slot
    const int lane_q = (warp & 3) * 32;  // TMEM lane quarter this warp may access
    int buf = 0;
    uint32_t aph = 0;
    const int NSTG = p.nstg;
    unsigned char* stg = smem + p.stg_off + ew * NSTG * STG_BYTES;
    int sbuf = 0;  // staging ring position
    // the bulk group that last used a staging buffer is NSTG groups old: wait until at most NSTG - 1 are unread
    auto wait_stg = [&]() {
      if (lane == 0) {
        switch (NSTG) {
          case 2: tma_store_wait_read<1>(); break;
          case 3: tma_store_wait_read<2>(); break;
          case 4: tma_store_wait_read<3>(); break;
          default: tma_store_wait_read<5>(); break;
        }
      }
      __syncwarp();
    };
    auto release_acc = [&]() {  // all TMEM reads of this accumulator are done: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 1) mbar_arrive(acc_empty(buf));
        else mbar_arrive_cluster(acc_empty(buf), 0);
      }
      if (++buf == NACC) { buf = 0; aph ^= 1; }
    };
    const int words = p.Kp >> 1;  // 32-bit words per packed A row and part (multiple of 32)
    uint32_t tile = 0, nunit = 0;
    long long w_full = 0, w_st = 0, w_ld = 0;
    long long* pw_full = p.prof ? &w_full : nullptr;
    const long long clk0 = clock64();
    for (int si = 0; si < nseg; ++si, ++nunit) {
      const Segment sg = segment(si);
      const int q_w = sg.mt * BM + lane_q;  // first query of this warp
      const int q = q_w + lane;
      const bool q_ok = q < p.Q;
      const long long bq = (long long)sg.b * p.Q + q;
      if (pooled_role) {
        // ---- A operand of this unit: global -> registers -> tensor memory.  Lane i copies the packed bf16 row of
        // its query (Kp/2 32-bit words per part) into columns [part*Kp/2, ...) of its TMEM lane.
        for (int part = 0; part < p.parts; ++part) {
          // all loads of a part are in flight before the first store (128 registers at Kp = 256); the loads of
          // part 0 are issued before waiting for the previous unit to release A
          const uint4* row = reinterpret_cast<const uint4*>(
              p.a_pack + (((long long)part * p.B + sg.b) * p.Q + (q_ok ? q : 0)) * words);
          uint32_t r[MAX_KB][32];
#pragma unroll
          for (int ch = 0; ch < MAX_KB; ++ch) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 v = make_uint4(0u, 0u, 0u, 0u);
              if (q_ok && ch * 32 < words) v = __ldg(row + ch * 8 + i);
              r[ch][4 * i + 0] = v.x; r[ch][4 * i + 1] = v.y; r[ch][4 * i + 2] = v.z; r[ch][4 * i + 3] = v.w;
            }
          }
          if (part == 0 && nunit > 0) mbar_wait(a_empty, (nunit - 1) & 1);  // previous unit's MMAs are done with A
          if (part == 0) tc_fence_after();
#pragma unroll
          for (int ch = 0; ch < MAX_KB; ++ch)
            if (ch * 32 < words) tmem_st32(tmem_base + ((uint32_t)lane_q << 16) + part * words + ch * 32, r[ch]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 1) mbar_arrive(a_full);
          else mbar_arrive_cluster(a_full, 0);
        }
      }
      for (int pi = sg.p_begin; pi < sg.p_end; ++pi, ++tile) {
        const int py = pi / p.pcols, px = pi % p.pcols;
        const int y0 = py * PH, x0 = px * PW;
        mbar_wait_t(acc_full(buf), aph, pw_full);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)lane_q << 16) + p.acc_col0 + buf * BN;
        if (!pooled_role) {
          // ---------------- level 0 ----------------
#pragma unroll
          for (int band = 0; band < 2; ++band) {  // 4 patch rows = one row of 4x4 tiles
            float va[32], vb[32];                 // rows 4*band + {0,1} and + {2,3}, 16 columns each
            if (p.debug_skip & 64) {
#pragma unroll
              for (int i = 0; i < 32; ++i) va[i] = vb[i] = 0.f;
            } else {
              const long long c0 = p.prof ? clock64() : 0;
              tmem_ld32(taddr + band * 64, va);
              tmem_ld32(taddr + band * 64 + 32, vb);
              if (p.prof) w_ld += clock64() - c0;
            }
            if (band == 1) release_acc();  // both bands are in registers
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              va[i] *= p.scale;
              vb[i] *= p.scale;
            }
            const int yy = y0 + 4 * band;
            if (p.f16) {
              // fp16 pyramid: the band is 2 tiles of 4 x 8 halfs = 128 contiguous bytes per query = ONE box
              if (yy < p.H && q_w < p.Q) {  // warp-uniform (x0 < W always holds)
                unsigned char* sb = stg + sbuf * STG_BYTES;
                {
                  const long long c0 = p.prof ? clock64() : 0;
                  wait_stg();
                  if (p.prof) w_st += clock64() - c0;
                }
                if (!(p.debug_skip & 8)) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) {  // chunk c = tile (c >> 2), tile row (c & 3): 8 halfs
                    const int r = c & 3;
                    const float* src = (r < 2) ? va : vb;
                    const int o = (r & 1) * 16 + 8 * (c >> 2);
                    *reinterpret_cast<uint4*>(sb + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                        pack_half8(src[o], src[o + 1], src[o + 2], src[o + 3], src[o + 4], src[o + 5], src[o + 6], src[o + 7]);
                  }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !(p.debug_skip & 1)) {
                  tma_store_4d(&map_l0, smem_u32(sb), (x0 >> 3) * 16, (yy >> 2), q_w, sg.b);
                  tma_store_commit();
                }
                if (++sbuf == NSTG) sbuf = 0;
              }
            } else if (yy < p.H && q_w < p.Q) {  // warp-uniform
              // The band's 4 tiles are 256 contiguous bytes per query; they leave as two 128-byte halves
              // (2 tiles each) through a SWIZZLE_128B staging box [32 queries][128 B]: conflict-free st.shared.
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                if (x0 + 8 * half >= p.W) break;  // warp-uniform: these tiles do not exist
                unsigned char* sb = stg + sbuf * STG_BYTES;
                {
                  const long long c0 = p.prof ? clock64() : 0;
                  wait_stg();  // the store that used this buffer has been read out
                  if (p.prof) w_st += clock64() - c0;
                }
                if (!(p.debug_skip & 8)) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) {  // chunk c = tile (c >> 2) of this half, tile row (c & 3)
                    const int col = 8 * half + 4 * (c >> 2);
                    const int r = c & 3;
                    const float* src = (r < 2) ? va : vb;
                    const int o = (r & 1) * 16 + col;
                    *reinterpret_cast<float4*>(sb + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                        make_float4(src[o], src[o + 1], src[o + 2], src[o + 3]);
                  }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !(p.debug_skip & 1)) {
                  tma_store_4d(&map_l0, smem_u32(sb), ((x0 >> 2) + 2 * half) * 16, (yy >> 2), q_w, sg.b);
                  tma_store_commit();
                }
                if (++sbuf == NSTG) sbuf = 0;
              }
            }
          }
        } else {
          // ---------------- levels 1-3: 2x2 means of the same accumulator values ----------------
          float l1[4][8];
#pragma unroll
          for (int band = 0; band < 2; ++band) {
            float va[32], vb[32];
            if (p.debug_skip & 64) {
#pragma unroll
              for (int i = 0; i < 32; ++i) va[i] = vb[i] = 0.f;
            } else {
              tmem_ld32(taddr + band * 64, va);
              tmem_ld32(taddr + band * 64 + 32, vb);
            }
            // scaled first: level 1 is the mean of the STORED level-0 values, bit for bit
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              va[i] *= p.scale;
              vb[i] *= p.scale;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              l1[2 * band][j] = ((va[2 * j] + va[2 * j + 1]) + (va[16 + 2 * j] + va[16 + 2 * j + 1])) * 0.25f;
              l1[2 * band + 1][j] = ((vb[2 * j] + vb[2 * j + 1]) + (vb[16 + 2 * j] + vb[16 + 2 * j + 1])) * 0.25f;
            }
          }
          release_acc();
          if (p.f16) {
            // fp16 pyramid.  The means are formed from the fp32 accumulator values and rounded once.
            if (p.levels > 1) {
              const int y1 = y0 >> 1, x1 = x0 >> 1;  // 4 rows x 8 cols = ONE tile = 64 contiguous bytes per query
              if (y1 < p.Hl[1] && x1 < p.Wl[1] && q_w < p.Q && !(p.debug_skip & 2)) {
                unsigned char* sb = stg + sbuf * STG_BYTES;
                wait_stg();
                // SWIZZLE_64B box [32 queries][64 B]: 16-byte chunk r of row `lane` sits at chunk r ^ ((lane >> 1) & 3)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                  *reinterpret_cast<uint4*>(sb + lane * 64 + ((r ^ ((lane >> 1) & 3)) << 4)) =
                      pack_half8(l1[r][0], l1[r][1], l1[r][2], l1[r][3], l1[r][4], l1[r][5], l1[r][6], l1[r][7]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_4d(&map_l1, smem_u32(sb), (x1 >> 3) * 16, (y1 >> 2), q_w, sg.b);
                  tma_store_commit();
                }
                if (++sbuf == NSTG) sbuf = 0;
              }
            }
            if (p.levels > 2 && q_ok && !(p.debug_skip & 4)) {
              float l2[2][4];
#pragma unroll
              for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  l2[r][j] = ((l1[2 * r][2 * j] + l1[2 * r][2 * j + 1]) + (l1[2 * r + 1][2 * j] + l1[2 * r + 1][2 * j + 1])) * 0.25f;
              __half* base2 = reinterpret_cast<__half*>(p.pyr[2]);
              const int y2 = y0 >> 2, x2 = x0 >> 2;  // y2 even, x2 a multiple of 4: two 8-byte half rows of one tile
              if (y2 < p.Hl[2] && x2 < p.Wl[2]) {
                __half* t2 = base2 + bq * p.ps[2] + tile_off_h(y2, x2, p.tx[2]);
                const __half2 a0 = __floats2half2_rn(l2[0][0], l2[0][1]), a1 = __floats2half2_rn(l2[0][2], l2[0][3]);
                const __half2 b0 = __floats2half2_rn(l2[1][0], l2[1][1]), b1 = __floats2half2_rn(l2[1][2], l2[1][3]);
                *reinterpret_cast<uint2*>(t2) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1));
                *reinterpret_cast<uint2*>(t2 + 8) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1));
              }
              if (p.levels > 3) {
                const float a = ((l2[0][0] + l2[0][1]) + (l2[1][0] + l2[1][1])) * 0.25f;
                const float c = ((l2[0][2] + l2[0][3]) + (l2[1][2] + l2[1][3])) * 0.25f;
                const int y3 = y0 >> 3, x3 = x0 >> 3;  // x3 even: both values sit in one tile row
                if (y3 < p.Hl[3] && x3 < p.Wl[3])
                  *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p.pyr[3]) + bq * p.ps[3] +
                                              tile_off_h(y3, x3, p.tx[3])) = __floats2half2_rn(a, c);
              }
            }
            continue;
          }
          if (p.levels > 1) {
            const int y1 = y0 >> 1, x1 = x0 >> 1;  // 4 rows x 8 cols = 2 tiles = 128 contiguous bytes per query
            if (y1 < p.Hl[1] && x1 < p.Wl[1] && q_w < p.Q && !(p.debug_skip & 2)) {
              unsigned char* sb = stg + sbuf * STG_BYTES;
              wait_stg();
#pragma unroll
              for (int c = 0; c < 8; ++c) {  // chunk c = tile (c >> 2), tile row (c & 3)
                const int r = c & 3, col = 4 * (c >> 2);
                *reinterpret_cast<float4*>(sb + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                    make_float4(l1[r][col], l1[r][col + 1], l1[r][col + 2], l1[r][col + 3]);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&map_l1, smem_u32(sb), (x1 >> 2) * 16, (y1 >> 2), q_w, sg.b);
                tma_store_commit();
              }
              if (++sbuf == NSTG) sbuf = 0;
            }
          }
          if (p.levels > 2 && q_ok && !(p.debug_skip & 4)) {
            float l2[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                l2[r][j] = ((l1[2 * r][2 * j] + l1[2 * r][2 * j + 1]) + (l1[2 * r + 1][2 * j] + l1[2 * r + 1][2 * j + 1])) * 0.25f;
            // level 2: 2 rows x 4 cols = two adjacent 16-byte rows of one tile (y2 is even, x2 a multiple of 4)
            const int y2 = y0 >> 2, x2 = x0 >> 2;
            if (y2 < p.Hl[2] && x2 < p.Wl[2]) {
              float* t2 = p.pyr[2] + bq * p.ps[2] + tile_off(y2, x2, p.tx[2]);
              *reinterpret_cast<float4*>(t2) = make_float4(l2[0][0], l2[0][1], l2[0][2], l2[0][3]);
              *reinterpret_cast<float4*>(t2 + 4) = make_float4(l2[1][0], l2[1][1], l2[1][2], l2[1][3]);
            }
            if (p.levels > 3) {
              const float a = ((l2[0][0] + l2[0][1]) + (l2[1][0] + l2[1][1])) * 0.25f;
              const float c = ((l2[0][2] + l2[0][3]) + (l2[1][2] + l2[1][3])) * 0.25f;
              const int y3 = y0 >> 3, x3 = x0 >> 3;  // x3 is even: both values sit in one tile row
              if (y3 < p.Hl[3] && x3 < p.Wl[3])
                *reinterpret_cast<float2*>(p.pyr[3] + bq * p.ps[3] + tile_off(y3, x3, p.tx[3])) = make_float2(a, c);
            }
          }
        }
      }
    }
    if (p.prof && lane == 0 && ew == 2) {  // warp 4 = TMEM lanes 0-31, level-0 role
      p.prof[blockIdx.x * 16 + 7] = clock64() - clk0;
      p.prof[blockIdx.x * 16 + 8] = w_full;
      p.prof[blockIdx.x * 16 + 9] = w_st;
      p.prof[blockIdx.x * 16 + 10] = w_ld;
      p.prof[blockIdx.x * 16 + 11] = tile;
    }
    if (p.prof && lane == 0 && ew == 6) {  // warp 8 = TMEM lanes 0-31, pooled-levels role
      p.prof[blockIdx.x * 16 + 12] = clock64() - clk0;
      p.prof[blockIdx.x * 16 + 13] = w_full;
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared and tensor memory until the end
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<NCTA>(tmem_base, 512);
  }
}

}  // namespace tc

static int padded_k(int C) { return (C + tc::BK - 1) / tc::BK * tc::BK; }

size_t build_tc_workspace_bytes(int B, int C, int H, int W, int mode) {
  const int parts = mode == RCB_BUILD_BF16X3 ? 2 : 1;
  return (size_t)2 * parts * B * H * W * padded_k(C) * sizeof(__nv_bfloat16);
}

template <int NCTA>
static int launch_main(const CUtensorMap& map_b, const CUtensorMap& map_l0, const CUtensorMap& map_l1, tc::Params& p,
                       int grid, int smem_total, cudaStream_t s) {
  using namespace tc;
  cudaError_t e = cudaFuncSetAttribute(build_tc_kernel<NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = (size_t)smem_total;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, build_tc_kernel<NCTA>, map_b, map_l0, map_l1, p);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

int launch_build_tc(const float* f1, const float* f2, void* const* pyr, const rcb_pyramid_layout& lay, int B,
                    int C, int H, int W, int mode, void* ws, size_t ws_bytes, cudaStream_t s) {
  using namespace tc;
  const bool f16 = lay.dtype == RCB_F16;
  const int esize = f16 ? 2 : 4;
  const int Kp = padded_k(C);
  if (Kp > MAX_KB * BK) return RCB_ERR_UNSUPPORTED;  // A must stay resident in tensor memory (C <= 256)
  if (!encode_fn()) return RCB_ERR_NO_DEVICE;
  const int parts = mode == RCB_BUILD_BF16X3 ? 2 : 1;
  const size_t need = build_tc_workspace_bytes(B, C, H, W, mode);
  if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 127)) return RCB_ERR_WORKSPACE;
  const int Q = H * W;
  // CTAs per MMA.  Both forms are verified by the tests; on B200 the kernel is bound by its stores under the 1 kW
  // power cap, where coupling two SMs costs more than the halved B traffic saves (cfg2: 643 us vs 693 us), so
  // single-CTA MMAs are the default.  RCB_TC_NCTA=2 selects the pairs.
  static const int ncta_env = [] { const char* e = getenv("RCB_TC_NCTA"); return e ? atoi(e) : 1; }();
  const int ncta = ncta_env == 2 ? 2 : 1;

  // 1. pack: fp32 NCHW -> bf16 hi/lo, K-major
  __nv_bfloat16* packed = static_cast<__nv_bfloat16*>(ws);
  {
    dim3 grid((Q + 63) / 64, Kp / 64, 2 * B);
    pack_operands_kernel<<<grid, 256, 0, s>>>(f1, f2, packed, B, C, Q, Kp, parts);
    int st = launch_status();
    if (st != RCB_OK) return st;
  }
  const __nv_bfloat16* a_pack = packed;
  const __nv_bfloat16* b_pack = packed + (size_t)parts * B * Q * Kp;

  // 2. tensor maps
  CUtensorMap map_b, map_l0, map_l1;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Kp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)parts * B};
    cuuint64_t str[3] = {(cuuint64_t)Kp * 2, (cuuint64_t)W * Kp * 2, (cuuint64_t)Q * Kp * 2};
    cuuint32_t box[4] = {BK, PW, (cuuint32_t)(PH / ncta), 1};  // each CTA of a pair loads half of the patch rows
    if (!encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, b_pack, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return RCB_ERR_INVALID_ARGUMENT;
  }
  // stores: a level is viewed as [B][Q][tile rows][tiles_x * 16 floats]; one box = 2 tiles (128 B) x 32 queries
  for (int l = 0; l < 2; ++l) {
    CUtensorMap* m = l == 0 ? &map_l0 : &map_l1;
    if (l >= lay.levels) {
      *m = map_l0;
      break;
    }
    // in 4-byte words: a tile is 16 words for both element types
    cuuint64_t dims[4] = {(cuuint64_t)lay.tiles_x[l] * 16, (cuuint64_t)lay.tiles_y[l], (cuuint64_t)Q, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)lay.tiles_x[l] * 64, (cuuint64_t)lay.plane_stride[l] * esize,
                         (cuuint64_t)Q * lay.plane_stride[l] * esize};
    const bool one_tile = f16 && l == 1;  // fp16 level 1: one 64-byte tile per query and patch
    cuuint32_t box[4] = {(cuuint32_t)(one_tile ? 16 : 32), 1, 32, 1};
    if (!encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, pyr[l], dims, str, box,
                one_tile ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B))
      return RCB_ERR_INVALID_ARGUMENT;
  }

  // 3. work decomposition
  Params p{};
  p.B = B; p.C = C; p.H = H; p.W = W; p.Q = Q;
  p.kblocks = Kp / BK;
  p.parts = parts;
  p.levels = lay.levels;
  p.mtiles = (Q + BM - 1) / BM;
  p.mgroups = (p.mtiles + ncta - 1) / ncta;
  p.pcols = (W + PW - 1) / PW;
  p.prows = (H + PH - 1) / PH;
  p.npatch = p.pcols * p.prows;
  p.scale = 1.0f / sqrtf((float)C);
  const char* prof = getenv("RCB_TC_PROF_PTR");  // debug: device buffer of 16 x 148 uint64 supplied by tools/time_build.py
  p.prof = prof ? reinterpret_cast<unsigned long long*>(strtoull(prof, nullptr, 0)) : nullptr;
  const char* skip = getenv("RCB_TC_DEBUG_SKIP");
  p.debug_skip = skip ? atoi(skip) : 0;
  for (int l = 0; l < RCB_MAX_LEVELS; ++l) {
    p.pyr[l] = l < lay.levels ? static_cast<float*>(pyr[l]) : nullptr;
    p.Hl[l] = lay.H[l]; p.Wl[l] = lay.W[l]; p.tx[l] = lay.tiles_x[l]; p.ps[l] = lay.plane_stride[l];
  }
  p.Kp = Kp;
  p.f16 = f16 ? 1 : 0;
  p.a_pack = reinterpret_cast<const uint32_t*>(a_pack);
  p.acc_col0 = (parts * (Kp / 2) + 127) / 128 * 128;  // A occupies the first parts*Kp/2 TMEM columns
  p.nacc = (512 - p.acc_col0) / BN < MAX_ACC ? (512 - p.acc_col0) / BN : MAX_ACC;

  int nstg = 2;
  if (const char* e = getenv("RCB_TC_NSTG")) nstg = atoi(e);
  if (nstg != 2 && nstg != 3 && nstg != 4 && nstg != 6) nstg = 2;
  p.nstg = nstg;
  const int stg_total = 2 * NUM_EPI_WARPS * nstg * STG_BYTES;
  const int stage_bytes = BOXES_PER_STAGE * B_TILE_BYTES / ncta;
  int nstage = (SMEM_BUDGET - BAR_BYTES - stg_total) / stage_bytes;
  if (nstage > MAX_STAGE) nstage = MAX_STAGE;
  if (const char* ns = getenv("RCB_TC_NSTAGE")) nstage = atoi(ns) < nstage ? atoi(ns) : nstage;
  if (nstage < 2) return RCB_ERR_UNSUPPORTED;
  p.nstage = nstage;
  p.b_off = 0;
  p.stg_off = p.b_off + nstage * stage_bytes;
  p.bar_off = p.stg_off + stg_total;
  const int smem_total = p.bar_off + BAR_BYTES;

  // one CTA per SM.  Units (batch, query-tile group) go round-robin over the clusters; what is left after the full
  // rounds is cut along the patch sequence so that the last round is short instead of mostly idle.
  const int units = B * p.mgroups;
  int clusters = kNumSMs / ncta;
  if (clusters > units * p.npatch) clusters = units * p.npatch;
  p.full_rounds = units / clusters;
  const int left = units % clusters;
  p.tail_split = left ? clusters / left : 1;
  if (p.tail_split > p.npatch) p.tail_split = p.npatch;
  p.tail_len = (p.npatch + p.tail_split - 1) / p.tail_split;
  p.tail_split = (p.npatch + p.tail_len - 1) / p.tail_len;
  p.tail_pieces = left * p.tail_split;
  const int grid = clusters * ncta;
  return ncta == 2 ? launch_main<2>(map_b, map_l0, map_l1, p, grid, smem_total, s)
                   : launch_main<1>(map_b, map_l0, map_l1, p, grid, smem_total, s);
}

}  // namespace rcb
