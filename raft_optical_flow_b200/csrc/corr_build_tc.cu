// K1, tensor-core modes: all-pairs correlation volume + fused 4-level pooled pyramid in ONE kernel.
//
// Replaces CorrBlock.corr (reference core/corr.py:96-127: matmul + / sqrt(C)) and the avg_pool2d loop of
// CorrBlock.__init__ (core/corr.py:52-54).  See DESIGN.md "K1" for the roofline discussion: with an fp32 pyramid the
// kernel is bound by its 2.2 GB of stores (the memory system takes 4.0-5.5 TB/s of this pattern,
// tools/probe/tma_store_probe2.cu), so the design goal is to keep the tensor pipe and the epilogue BELOW that bound.
//
//   pack kernel   fp32 NCHW feature maps -> K-major tensor-core operands (a transpose + split), per build mode:
//     RCB_BUILD_F16F8    x = hi + lo, hi = fp16(x).  hi*hi runs as ONE full-rate kind::f16 pass; the two cross terms
//                        hi*lo + lo*hi only need ~4 significant bits (they are 2^-12 of the product), so they run in
//                        8-bit e4m3 on kind::f8f6f4 (K = 32 per instruction): e4m3(hi) x e4m3(lo * 2^12).  The 2^12 is
//                        undone for free by the first kind::f16 MMA of a tile, which scales the accumulator it adds to
//                        (tcgen05.mma ... scale-input-d = 12).  32 instead of 48 MMAs per tile (2 pass-equivalents of
//                        tensor work instead of 3), error ~1e-5 of max-abs (tools/split_precision_probe.py; bound 1e-4).
//     RCB_BUILD_BF16X3   bf16 hi/lo, hi*hi + hi*lo + lo*hi, three full-rate passes, error ~5e-6 of max-abs; no range
//                        restriction (fp16 overflows beyond 65504).
//     RCB_BUILD_BF16     hi parts only (fast mode, ~2.6e-3).
//   main kernel   persistent, warp-specialised, 352 threads per CTA, one CTA per SM:
//     warp 0      producer: B tiles (an 8 x 16 PATCH of target pixels x 128 bytes of K = a 16 KB box) stream through an
//                 mbarrier ring, one box per stage.  The pack kernel writes the target operand as ready-made TILE
//                 IMAGES -- per (batch, patch) the boxes of the tile's MMA schedule back to back, rows already in the
//                 SWIZZLE_128B order the MMA descriptors expect, out-of-image targets as zeros -- so a box is ONE
//                 linear cp.async.bulk of 16 contiguous KB.  (Round 1 fetched each box as a 4-D tensor-map box of 128
//                 rows x 128 B: the TMA unit then spends ~3.3 cycles per row, 1024 rows per tile, and the B stream
//                 alone took 290 us of the kernel, profiles/r2b_build_anatomy_band_split.txt.)
//     warps 1,10  MMA issuers, alternate tiles: M = 128 queries x N = 128 targets per instruction, the A operand
//                 (queries) lives in TENSOR MEMORY for a whole unit (TS form), B through SWIZZLE_128B K-major smem
//                 descriptors, fp32 accumulators in TMEM (two, one per issuer).  The per-tile MMA schedule is a small
//                 table (Params::box).  Two issuers because the issuing side, not the tensor pipe, paced the MMAs
//                 (see the comment at the issuer loop).
//     warps 2-9   epilogue, two warps per TMEM lane quarter (32 queries), split by patch BAND (4 of the 8 patch rows =
//                 one row of four 4x4 pyramid tiles = 256 contiguous bytes per query): a warp tcgen05.ld's its band
//                 (lane = query, 64 columns), scales by 1/sqrt(C), stages it as two SWIZZLE_128B boxes
//                 [32 queries][128 B], forms its two rows of the level-1 block and its row of the level-2 block in
//                 registers (the patch is 8x8-aligned, so all three pooled levels are patch-local; floor-mode dropping
//                 of odd rows/cols needs no code: a level-k cell computed from a dropped row/col is itself outside
//                 H_k x W_k, i.e. tile padding or a clipped tile), adds them to the pair's shared level-1 box, executes
//                 ONE fence.proxy.async, meets its partner at a named barrier and issues its TMA stores (band 0 also
//                 the level-1 box).  Band 1 then writes levels 2 and 3 for both: the two level-2 rows of a patch are
//                 32 contiguous bytes of one tile, i.e. a FULL sector per query (band 0's row comes over through
//                 shared memory) -- written as two 16-byte halves by two warps they cost ~30 us more.  Round 1 split
//                 the warps by output LEVEL instead: the four level-0 warps then spent ~4000 cycles per tile in four
//                 fence / wait / issue rounds and were, with the MMAs, the critical path (role counters in
//                 profiles/r2a_build_anatomy_round1_kernel.txt) while the pooled-level warps idled and re-read the
//                 accumulator.  Levels 2/3 through TMA boxes of 32 / 16-byte rows were measured too: cheaper when the
//                 stores run alone, slower in the full kernel (more traffic through the TMA unit that also feeds B).
//     both        epilogue warps of a quarter also place the A operand in tensor memory (global -> registers ->
//                 tcgen05.st) at the start of every unit, half of the columns each.
//   work split    a unit = all patches of one (batch, 128-query tile); units go round-robin over the CTAs, which
//                 therefore sweep the same batch's patches in step (B tiles are found in L2); the units left over
//                 after the full rounds are cut along the patch sequence so the last round is short, not idle.
//                 Patches are swept in vertical pairs, (0,0) (1,0) (0,1) (1,1) ...: the 2 x 2 patches that share a
//                 128-byte line of level 2 then follow each other closely and the line is completed while it is still
//                 in L2 (row-major order writes it in pieces a whole patch row apart: +20 us).
//   what bounds it  the store path: this pattern (128-byte rows, one per query plane, 28 KB apart) is absorbed at
//                 4.0-5.5 TB/s depending on the GPU of the pool (tools/probe/tma_store_probe2.cu; a plain fill reaches
//                 7.3), the epilogue + stores alone take ~500 us at cfg2, and MMAs and B loads add 60-90 us of
//                 interference on top (profiles/r2c..r2l; measured-and-rejected variants are listed in DESIGN.md).
#include "rcb_common.cuh"
#include "tcgen05_util.cuh"
#include "tma_util.cuh"

#include <cuda_fp8.h>

// Build-time variant (A/B timing inside one GPU session: python -m raft_optical_flow_b200.build --variant NAME -D...)
#ifndef RCB_PAIR_ORDER
#define RCB_PAIR_ORDER 1     // patch sweep in vertical pairs: (0,0) (1,0) (0,1) (1,1) ... (0: row-major)
#endif

namespace rcb {

namespace tc {

constexpr int BM = 128;          // queries per CTA tile (TMEM lanes)
constexpr int PH = 8, PW = 16;   // target patch: 8 rows x 16 cols
constexpr int BN = PH * PW;      // 128 targets per tile (TMEM columns)
constexpr int BOX_K_BYTES = 128; // K extent of one B box in bytes (SWIZZLE_128B row): 64 x 16-bit or 128 x 8-bit
constexpr int BOX_BYTES = BN * BOX_K_BYTES;  // 16 KB
constexpr int BOXES_PER_STAGE = 1;
constexpr int STAGE_BYTES = BOXES_PER_STAGE * BOX_BYTES;
constexpr int MAX_STAGE = 12;    // B ring depth is chosen at launch from the shared memory available
constexpr int MAX_ACC = 4;       // TMEM accumulator buffers (128 columns each); count chosen at launch
constexpr int MAX_BOX = 8;       // B boxes per tile (C <= 256)
constexpr int NUM_EPI_WARPS = 8;
#ifndef RCB_NUM_ISSUERS
#define RCB_NUM_ISSUERS 2        // MMA-issuing warps: tiles alternate between them (1: a single issuer)
#endif
constexpr int NUM_ISSUERS = RCB_NUM_ISSUERS;
constexpr int THREADS = 32 * (1 + NUM_ISSUERS + NUM_EPI_WARPS);  // producer, issuer A, 8 epilogue warps, [issuer B]
constexpr int STG0_BYTES = 8192;  // per epilogue warp: the band's two level-0 boxes
constexpr int STG1_BYTES = 4096;  // per lane quarter: the level-1 box both warps of the pair fill
constexpr int XCH_BYTES = 512;    // per lane quarter: band 0's level-2 row for its partner (16 bytes per query)

// Dynamic shared memory (all tile buffers 1024-byte aligned for the swizzle atoms):
//   [b_off, +nstage * 16 KB)   B ring -- as deep as the rest allows (9 boxes): the kernel is latency-bound on the B
//                              stream while the store path is saturated (4 boxes: 900 us, 7: 665 us at cfg2)
//   [stg0_off, +8 * 8 KB)      level-0 staging, one band per epilogue warp
//   [stg1_off, +4 * 4 KB)      level-1 staging, one box per lane quarter
//   [xch_off, +4 * 512 B)      level-2 rows handed from band 0 to band 1
//   [bar_off, +1 KB)           mbarriers + TMEM base slot
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int BAR_BYTES = 1024;

enum MmaKind { KIND_F16 = 0, KIND_F8 = 1 };

// One B box of a tile (box j of the tile image) and the MMAs it feeds (four K steps of 32 bytes each).
struct BoxDesc {
  int a_col0;   // TMEM column of the A operand this box meets
  int a_col1;   // a second A operand for the same box (bf16x3: B_hi also meets A_lo), -1 = none
  int kind;     // MmaKind
  uint32_t idesc;
  int scaled;   // 1: the first MMA of this box rescales the accumulator by 2^-12 (f16f8: first kind::f16 box)
};

struct Params {
  int B, C, H, W, Q;
  int levels;
  int mtiles;       // ceil(Q / 128)
  int pcols, prows; // patch grid
  int npatch;       // pcols * prows
  int full_rounds;  // rounds in which every CTA sweeps all patches of one (batch, query tile) unit
  int tail_pieces, tail_split, tail_len;  // left-over units: cut into tail_split ranges of tail_len patches
  float scale;      // 1 / sqrt(C)
  int nstage, b_off, stg0_off, stg1_off, xch_off, bar_off;  // shared memory carve-up (bytes)
  int f16;                              // pyramid stored as fp16 (tiles of 4 rows x 8 columns), else fp32 (4 x 4)
  int nacc, acc_col0;                   // TMEM: accumulator count, first accumulator column
  int a_words;                          // 32-bit words of one packed A row = TMEM columns of the A operand
  const uint32_t* a_pack;               // packed A operand image, [B][Q][a_words]
  const unsigned char* b_img;           // packed B tile images, [B][npatch][nbox][128 targets][128 B], pre-swizzled
  int nbox;
  BoxDesc box[MAX_BOX];
  float* pyr[RCB_MAX_LEVELS];           // (fp16 pyramids: the same pointers, reinterpreted)
  int Hl[RCB_MAX_LEVELS], Wl[RCB_MAX_LEVELS], tx[RCB_MAX_LEVELS];  // level sizes, tiles per tile row
  long long ps[RCB_MAX_LEVELS];
#ifdef RCB_DEBUG
  unsigned long long* prof;  // per-CTA cycle counters (16 per CTA)
  int debug_skip;   // bitmask: 1 skip level-0 TMA stores, 2 skip the level-1 store, 4 skip level 2/3, 8 skip staging writes,
                    // 16 skip B loads, 32 skip MMAs, 64 skip TMEM loads, 128 skip the whole epilogue body
#endif
};

#ifdef RCB_DEBUG
#define RCB_SKIP(p, bit) (((p).debug_skip & (bit)) != 0)
// epilogue phase timers: PH_MARK(i) adds the cycles since the previous mark to phase i
#define PH_DECL long long ph_t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long ph_last = 0
#define PH_START() do { if (p.prof) ph_last = clock64(); } while (0)
#define PH_MARK(i) do { if (p.prof) { const long long now_ = clock64(); ph_t[i] += now_ - ph_last; ph_last = now_; } } while (0)
#else
#define RCB_SKIP(p, bit) false
#define PH_DECL
#define PH_START()
#define PH_MARK(i)
#endif

// kind::f16 / kind::f8f6f4 instruction descriptor: fp32 accumulate, both operands K-major, M = 128, N = 128;
// fmt: 0 = fp16 (kind::f16) or e4m3 (kind::f8f6f4), 1 = bf16
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T, A: lane = row, one 32-bit column = 32 bits of K (two 16-bit or four 8-bit elements)
RCB_DEVINL void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same, but the accumulator is multiplied by 2^-12 before the product is added (scale-input-d)
RCB_DEVINL void umma_f16_ts_scaled12(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 12;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc)
      : "memory");
}
// 8-bit operands (e4m3 x e4m3), K = 32 per instruction, twice the rate of kind::f16 per K element
RCB_DEVINL void umma_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
RCB_DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- operand packing -----------------------------------------------------------------------
// Where the operands go.  Queries (fmap1, the A operand): one row of a_row bytes per pixel = the image of the
// pixel's TMEM lane, [hi16 | lo16] (bf16x3) or [hi16 | e4m3(hi) | e4m3(lo * 2^12)] (f16f8).  Targets (fmap2, the B
// operand): tile images -- for every (batch, patch) the tile's boxes back to back, a box = 128 targets (row n =
// patch row * 16 + patch column) x 128 bytes of K with the 16-byte chunks of row n XOR-swizzled by n & 7 (the
// SWIZZLE_128B shared-memory layout), so the main kernel fetches a stage with one linear bulk copy.
struct PackParams {
  unsigned char* a_img;
  long long a_row;               // bytes per query row
  int a_lo16, a_hi8, a_lo8;      // byte offsets inside a row (a_lo16 < 0: no 16-bit lo part)
  unsigned char* b_img;
  int nbox, pcols, npatch, Hp, Wp;  // boxes per tile; patch grid; padded target grid (multiples of 8 x 16)
  int box16_hi[4], box16_lo[4];  // box index of the 16-bit k-block kb (64 channels each); lo: bf16x3 only
  int box8_hi[2], box8_lo[2];    // box index of the 8-bit k-block k8 (128 channels each); f16f8 only
  int K16, K8;                   // padded channel counts of the two operand widths
};

// in  [2 maps][B][C][Q] fp32 (two separate base pointers), zero padded beyond C and outside the image.
// CTA = 64 channels x 64 pixels (queries: pixels of the image; targets: pixels of the PADDED grid): coalesced
// loads along the row into a padded fp32 tile, then every thread turns 8 channels of one pixel into one 16-byte
// store per 16-bit part (8 lanes cover the 128 bytes of a pixel's k-block) and one 8-byte store per 8-bit part.
template <bool F8>
__global__ void __launch_bounds__(256)
pack_operands_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const __grid_constant__ PackParams d,
                     int B, int C, int H, int W) {
  __shared__ float tile[64][65];  // odd pitch: the column reads below are at most 2-way bank-conflicted
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int map = blockIdx.z / B, b = blockIdx.z % B;
  const int Q = H * W;
  const int npix = map == 0 ? Q : d.Hp * d.Wp;  // pixel space of this map
  const int q0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  if (q0 >= npix) return;
  const float* in = (map == 0 ? f1 : f2) + (long long)b * C * Q;
  // source offsets of this lane's two pixels (-1: outside the image)
  int s0, s1;
  {
    const int i0 = q0 + 2 * lane, i1 = i0 + 1;
    if (map == 0) {
      s0 = i0 < Q ? i0 : -1;
      s1 = i1 < Q ? i1 : -1;
    } else {
      const int y0 = i0 / d.Wp, x0 = i0 - y0 * d.Wp, y1 = i1 / d.Wp, x1 = i1 - y1 * d.Wp;
      s0 = (y0 < H && x0 < W) ? y0 * W + x0 : -1;
      s1 = (y1 < H && x1 < W) ? y1 * W + x1 : -1;
    }
  }
  const bool pair = s0 >= 0 && s1 == s0 + 1 && (s0 & 1) == 0 && (Q & 1) == 0;  // one aligned 8-byte load
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = warp * 8 + i, cc = c0 + c;
    float2 v = make_float2(0.f, 0.f);
    if (cc < C) {
      const float* src = in + (long long)cc * Q;
      if (pair) {
        v = __ldg(reinterpret_cast<const float2*>(src + s0));
      } else {
        if (s0 >= 0) v.x = __ldg(src + s0);
        if (s1 >= 0) v.y = __ldg(src + s1);
      }
    }
    tile[c][2 * lane] = v.x;
    tile[c][2 * lane + 1] = v.y;
  }
  __syncthreads();
  const int cg = lane & 7;  // group of 8 channels
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ql = warp * 8 + i * 4 + (lane >> 3), q = q0 + ql;
    if (q >= npix) continue;
    const int ch = c0 + 8 * cg;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = tile[8 * cg + j][ql];
    // destinations: queries -> a row of the A image; targets -> row n of the patch's boxes
    unsigned char *d16_hi = nullptr, *d16_lo = nullptr, *d8_hi = nullptr, *d8_lo = nullptr;
    if (map == 0) {
      unsigned char* row = d.a_img + ((long long)b * Q + q) * d.a_row;
      if (ch < d.K16) {
        d16_hi = row + ch * 2;
        if (d.a_lo16 >= 0) d16_lo = row + d.a_lo16 + ch * 2;
      }
      if (F8) {
        d8_hi = row + d.a_hi8 + ch;
        d8_lo = row + d.a_lo8 + ch;
      }
    } else {
      const int yp = q / d.Wp, xp = q - yp * d.Wp;
      const int n = (yp & 7) * 16 + (xp & 15);
      unsigned char* tb = d.b_img + (((long long)b * d.npatch + (yp >> 3) * d.pcols + (xp >> 4)) * d.nbox) * 16384 +
                          n * 128;
      if (ch < d.K16) {
        const int kb = ch >> 6, off = ((((ch & 63) >> 3) ^ (n & 7)) << 4);
        d16_hi = tb + d.box16_hi[kb] * 16384 + off;
        if (d.box16_lo[kb] >= 0) d16_lo = tb + d.box16_lo[kb] * 16384 + off;
      }
      if (F8) {
        const int k8 = ch >> 7, off = ((((ch & 127) >> 4) ^ (n & 7)) << 4) + (ch & 15);
        d8_hi = tb + d.box8_hi[k8] * 16384 + off;
        d8_lo = tb + d.box8_lo[k8] * 16384 + off;
      }
    }
    if (!F8) {
      __nv_bfloat162 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * j]), h1 = __float2bfloat16_rn(x[2 * j + 1]);
        hi[j] = __halves2bfloat162(h0, h1);
        lo[j] = __halves2bfloat162(__float2bfloat16_rn(x[2 * j] - __bfloat162float(h0)),
                                   __float2bfloat16_rn(x[2 * j + 1] - __bfloat162float(h1)));
      }
      if (d16_hi) *reinterpret_cast<uint4*>(d16_hi) = *reinterpret_cast<const uint4*>(hi);
      if (d16_lo) *reinterpret_cast<uint4*>(d16_lo) = *reinterpret_cast<const uint4*>(lo);
    } else {
      // hi = fp16(x) (the full-rate pass), and for the cross terms e4m3(hi) and e4m3((x - hi) * 2^12):
      // x - hi is exact in fp32 and at most half an fp16 ulp of x, so the scaled lo part never exceeds |x| in
      // magnitude (no overflow where hi fits) and keeps ~4 significant bits, all that 2^-12-weighted terms need
      __half2 hi[4];
      uint32_t h8[2], l8[2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
        const float2 hf = __half22float2(hi[j]);
        const uint32_t a = (uint32_t)__nv_cvt_float2_to_fp8x2(hf, __NV_SATFINITE, __NV_E4M3);
        const uint32_t l = (uint32_t)__nv_cvt_float2_to_fp8x2(
            make_float2((x[2 * j] - hf.x) * 4096.0f, (x[2 * j + 1] - hf.y) * 4096.0f), __NV_SATFINITE, __NV_E4M3);
        if (j & 1) {
          h8[j >> 1] |= a << 16;
          l8[j >> 1] |= l << 16;
        } else {
          h8[j >> 1] = a;
          l8[j >> 1] = l;
        }
      }
      if (d16_hi) *reinterpret_cast<uint4*>(d16_hi) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint2*>(d8_hi) = make_uint2(h8[0], h8[1]);
      *reinterpret_cast<uint2*>(d8_lo) = make_uint2(l8[0], l8[1]);
    }
  }
}

// shared::cta <- global linear bulk copy, completion counted in bytes on an mbarrier
RCB_DEVINL void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
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
// One piece of a CTA's work: a patch range of one (batch, query tile) unit.
struct Segment {
  int b, mt, p_begin, p_end;
};

// Position i of the patch sweep -> patch (py, px).  Vertical pairs: the four patches (2 x 2) that share a 128-byte
// line of level 2 are swept within four consecutive tiles, so the line is completed while it is still in L2 instead
// of being written in four pieces a whole patch row apart (each piece evicted on its own by the 2 GB stream).
RCB_DEVINL void sweep_to_patch(int i, int prows, int pcols, int& py, int& px) {
#if RCB_PAIR_ORDER
  const int full = (prows >> 1) * 2 * pcols;  // sweep positions inside complete row pairs
  if (i < full) {
    const int rp = i / (2 * pcols), w = i - rp * 2 * pcols;
    px = w >> 1;
    py = 2 * rp + (w & 1);
  } else {
    py = prows - 1;
    px = i - full;
  }
#else
  py = i / pcols;
  px = i - py * pcols;
#endif
}

__global__ void __launch_bounds__(THREADS, 1)
build_tc_kernel(const __grid_constant__ CUtensorMap map_l0, const __grid_constant__ CUtensorMap map_l1,
                const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // barriers (8 bytes each)
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
    mbar_init(a_full, NUM_EPI_WARPS);  // every epilogue warp loads a share of A
    mbar_init(a_empty, NUM_ISSUERS);  // every issuer commits once per unit
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // This CTA's work: segment i < full_rounds is the whole patch sweep of unit i * nctas + cid (unit = batch, query
  // tile); the units left over after the full rounds are cut into `tail_split` patch ranges each so that every CTA
  // gets at most one piece.  CTAs of one round sweep the same batch's patches in step, so a B tile is fetched from
  // DRAM once and found in L2 by the other CTAs.
  const int nctas = gridDim.x, cid = blockIdx.x;
  const int nseg = p.full_rounds + (cid < p.tail_pieces ? 1 : 0);
  auto segment = [&](int i) {
    Segment sg;
    int u;
    if (i < p.full_rounds) {
      u = i * nctas + cid;
      sg.p_begin = 0;
      sg.p_end = p.npatch;
    } else {
      u = p.full_rounds * nctas + cid / p.tail_split;
      sg.p_begin = (cid % p.tail_split) * p.tail_len;
      sg.p_end = min(p.npatch, sg.p_begin + p.tail_len);
    }
    sg.b = u / p.mtiles;
    sg.mt = u % p.mtiles;
    return sg;
  };
#ifdef RCB_DEBUG
  const long long clk0 = clock64();
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  long long* pw0 = p.prof ? &w0 : nullptr;
  long long* pw1 = p.prof ? &w1 : nullptr;
  long long* pw2 = p.prof ? &w2 : nullptr;
#else
  long long* const pw0 = nullptr;
  long long* const pw1 = nullptr;
  long long* const pw2 = nullptr;
#endif

  if (warp == 0) {
    // =============================== TMA producer (whole warp runs the loop, one elected lane issues) ====
    int s = 0;           // B ring slot
    uint32_t ph = 0;     // its phase
    for (int si = 0; si < nseg; ++si) {
      const Segment sg = segment(si);
      for (int pi = sg.p_begin; pi < sg.p_end; ++pi) {
        int py, px;
        sweep_to_patch(pi, p.prows, p.pcols, py, px);
        const unsigned char* tile_img = p.b_img + ((long long)sg.b * p.npatch + py * p.pcols + px) * p.nbox * BOX_BYTES;
        for (int j = 0; j < p.nbox; j += BOXES_PER_STAGE) {
          mbar_wait_t(b_empty(s), ph ^ 1, pw0);
          if (elect_one()) {
            const int nbox = min(BOXES_PER_STAGE, p.nbox - j);
            if (RCB_SKIP(p, 16)) {
              mbar_arrive(b_full(s));
            } else {  // the stage's boxes are contiguous in the tile image: one linear copy
              mbar_expect_tx(b_full(s), (uint32_t)(nbox * BOX_BYTES));
              bulk_load(smem_base + p.b_off + s * STAGE_BYTES, tile_img + (long long)j * BOX_BYTES,
                        (uint32_t)(nbox * BOX_BYTES), b_full(s));
            }
          }
          __syncwarp();
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
#ifdef RCB_DEBUG
    if (p.prof && lane == 0) {
      p.prof[blockIdx.x * 16 + 0] = clock64() - clk0;
      p.prof[blockIdx.x * 16 + 1] = w0;
    }
#endif
  } else if (warp == 1 || warp >= 2 + NUM_EPI_WARPS) {
    // =============================== MMA issuers (whole warp runs the loop, one lane issues) ===
    // The issuing side costs ~200 cycles per box (barrier wait, tcgen05 fence, elect, commit) and ~500 per tile on top
    // of ~60 cycles per MMA, and tcgen05.mma issue is effectively synchronous (the time of the loop is the SUM,
    // profiles/r2s_mma_pipeline_only.txt), so with 4 MMAs per 16 KB box the tensor pipe idled 40 % of the time under
    // ONE issuer.  Two issuer warps take alternate tiles (with two accumulators each owns one): the overhead of one
    // overlaps the MMAs of the other.  Both follow the same global box sequence through the ring; a stage is freed
    // and an accumulator published by the commit of the thread whose MMAs used it; the A operand is released to the
    // next unit when BOTH have committed (a_empty counts NUM_ISSUERS).
    const int me = warp == 1 ? 0 : 1;  // issuer index
    const uint64_t desc_base = make_smem_desc(0);  // everything but the start address
    int s = 0, buf = 0;
    uint32_t ph = 0, aph = 0, nunit = 0, gtile = 0;
    for (int si = 0; si < nseg; ++si, ++nunit) {
      const Segment sg = segment(si);
      mbar_wait_t(a_full, nunit & 1, pw2);
      tc_fence_after();
      for (int pi = sg.p_begin; pi < sg.p_end; ++pi, ++gtile) {
        if ((int)(gtile % NUM_ISSUERS) != me) {
          // The other issuer's tile: step over its accumulator and its boxes -- but OBSERVE every phase of the stage
          // barriers on the way.  A parity wait is only unambiguous one phase ahead of the last phase the waiter has
          // seen; stepping over a lap would let this warp take the completed lap before it for its own box.
          for (int j = 0; j < p.nbox; ++j) {
            mbar_wait(b_full(s), ph);
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
          if (++buf == NACC) { buf = 0; aph ^= 1; }  // (NACC is a multiple of NUM_ISSUERS: its phases are never skipped)
          continue;
        }
        mbar_wait_t(acc_empty(buf), aph ^ 1, pw1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + p.acc_col0 + buf * BN;
        uint32_t acc = 0;
        for (int j = 0; j < p.nbox; ++j) {
          mbar_wait_t(b_full(s), ph, pw0);
          tc_fence_after();
          if (elect_one()) {
            if (!RCB_SKIP(p, 32)) {
              const BoxDesc& bx = p.box[j];
              const uint64_t bdesc = desc_base | (uint64_t)(((smem_base + p.b_off + s * STAGE_BYTES) >> 4) & 0x3FFF);
              const uint32_t ta0 = tmem_base + bx.a_col0;
              if (bx.kind == KIND_F8) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // +32 bytes (smem) / +8 columns (TMEM) per K step
                  umma_f8_ts(d_tmem, ta0 + 8 * k, bdesc + 2 * k, bx.idesc, acc);
                  acc = 1;
                }
              } else {
                if (bx.scaled) umma_f16_ts_scaled12(d_tmem, ta0, bdesc, bx.idesc);  // acc = 2^-12 acc + a * b
                else umma_f16_ts(d_tmem, ta0, bdesc, bx.idesc, acc);
                acc = 1;
#pragma unroll
                for (int k = 1; k < 4; ++k) umma_f16_ts(d_tmem, ta0 + 8 * k, bdesc + 2 * k, bx.idesc, 1u);
                if (bx.a_col1 >= 0) {  // bf16x3: B_hi also meets A_lo
                  const uint32_t ta1 = tmem_base + bx.a_col1;
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma_f16_ts(d_tmem, ta1 + 8 * k, bdesc + 2 * k, bx.idesc, 1u);
                }
              }
            }
            umma_commit<1>(b_empty(s));  // frees the stage once these MMAs have read it
          }
          acc = 1;
          __syncwarp();
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit<1>(acc_full(buf));
        __syncwarp();
        if (++buf == NACC) { buf = 0; aph ^= 1; }
      }
      if (elect_one()) umma_commit<1>(a_empty);  // (arrives at once if this issuer had no tile in the unit)
      __syncwarp();
    }
#ifdef RCB_DEBUG
    if (p.prof && lane == 0 && me == 0) {
      p.prof[blockIdx.x * 16 + 3] = clock64() - clk0;
      p.prof[blockIdx.x * 16 + 4] = w0;
      p.prof[blockIdx.x * 16 + 5] = w1;
      p.prof[blockIdx.x * 16 + 6] = w2;
    }
#endif
  } else {
    // =============================== epilogue (+ A loading) ===============================
    const int ew = warp - 2;              // 0..7
    const int band = ew >> 2;             // patch rows 4 * band .. 4 * band + 3
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access; warps (w, w + 4) form a pair
    const int lane_q = quarter * 32;
    // named barriers of the pair (barrier 0 is __syncthreads): "the shared boxes may be overwritten" / "are complete"
    const uint32_t bar_free = 1 + quarter, bar_done = 5 + quarter;
    unsigned char* stg0 = smem + p.stg0_off + ew * STG0_BYTES;
    unsigned char* sb1 = smem + p.stg1_off + quarter * STG1_BYTES;
    unsigned char* xch = smem + p.xch_off + quarter * XCH_BYTES;
    const int swz = lane & 7;             // SWIZZLE_128B: 16-byte chunk c of row `lane` sits at chunk c ^ (lane & 7)
    int buf = 0;
    uint32_t aph = 0;
    auto release_acc = [&]() {  // this warp's TMEM reads of the accumulator are done: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      if (++buf == NACC) { buf = 0; aph ^= 1; }
    };
    uint32_t tile = 0, nunit = 0;
    PH_DECL;
    for (int si = 0; si < nseg; ++si, ++nunit) {
      const Segment sg = segment(si);
      const int q_w = sg.mt * BM + lane_q;  // first query of this warp
      const int q = q_w + lane;
      const bool q_ok = q < p.Q;
      const long long bq = (long long)sg.b * p.Q + q;
      {
        // ---- A operand of this unit: global -> registers -> tensor memory.  Lane i copies the packed row of its
        // query (a_words 32-bit words) into its TMEM lane; the two warps of a pair take alternate 32-word chunks.
        // All loads are in flight before the first store (128 registers), and they are issued before waiting for
        // the previous unit to release A.
        const uint4* row = reinterpret_cast<const uint4*>(p.a_pack + ((long long)sg.b * p.Q + (q_ok ? q : 0)) * p.a_words);
        uint32_t r[4][32];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int ch = 2 * i4 + band;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (q_ok && ch * 32 < p.a_words) v = __ldg(row + ch * 8 + i);
            r[i4][4 * i + 0] = v.x; r[i4][4 * i + 1] = v.y; r[i4][4 * i + 2] = v.z; r[i4][4 * i + 3] = v.w;
          }
        }
        if (nunit > 0) mbar_wait(a_empty, (nunit - 1) & 1);  // previous unit's MMAs are done with A
        tc_fence_after();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int ch = 2 * i4 + band;
          if (ch * 32 < p.a_words) tmem_st32(tmem_base + ((uint32_t)lane_q << 16) + ch * 32, r[i4]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }
      for (int pi = sg.p_begin; pi < sg.p_end; ++pi, ++tile) {
        int py, px;
        sweep_to_patch(pi, p.prows, p.pcols, py, px);
        const int y0 = py * PH, x0 = px * PW;
        PH_START();
        mbar_wait_t(acc_full(buf), aph, pw0);
        tc_fence_after();
        PH_MARK(0);
        const uint32_t taddr = tmem_base + ((uint32_t)lane_q << 16) + p.acc_col0 + buf * BN + band * 64;
        float va[32], vb[32];  // patch rows 4 * band + {0, 1} and + {2, 3}, 16 columns each
        if (RCB_SKIP(p, 64)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) va[i] = vb[i] = 0.f;
        } else {
          tmem_ld32(taddr, va);
          tmem_ld32(taddr + 32, vb);
        }
        PH_MARK(1);
        release_acc();
        PH_MARK(2);
        if (RCB_SKIP(p, 128)) continue;  // debug: no epilogue work at all (pure producer + MMA pipeline)
        // scaled first: level 1 is the mean of the STORED level-0 values, bit for bit
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          va[i] *= p.scale;
          vb[i] *= p.scale;
        }
        // this band's share of the pooled levels: two rows of the patch's 4 x 8 level-1 block, one row of its
        // 2 x 4 level-2 block
        float l1a[8], l1b[8], l2[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          l1a[j] = ((va[2 * j] + va[2 * j + 1]) + (va[16 + 2 * j] + va[16 + 2 * j + 1])) * 0.25f;
          l1b[j] = ((vb[2 * j] + vb[2 * j + 1]) + (vb[16 + 2 * j] + vb[16 + 2 * j + 1])) * 0.25f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) l2[j] = ((l1a[2 * j] + l1a[2 * j + 1]) + (l1b[2 * j] + l1b[2 * j + 1])) * 0.25f;

        const int yy = y0 + 4 * band;
        const bool l0_ok = yy < p.H && q_w < p.Q;  // warp-uniform (x0 < W always holds)
        const int y1 = y0 >> 1, x1 = x0 >> 1;
        const bool l1_ok = p.levels > 1 && y1 < p.Hl[1] && x1 < p.Wl[1] && q_w < p.Q;  // the same in both warps of a pair
        const int y2 = (y0 >> 2) + band, x2 = x0 >> 2;  // this band's level-2 row; x2 is a multiple of 4
        PH_MARK(3);
        // every bulk store this lane 0 issued so far has read its staging buffer (they are one tile old); after the
        // pair's first barrier that also holds for the level-1 box band 0 issued, and band 1 has read the last
        // level-2 row it was handed -- both shared buffers may be overwritten
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        PH_MARK(4);
        named_bar_sync(bar_free, 64);
        if (!p.f16) {
          // ---- fp32 pyramid: the band is four 4x4 tiles = 256 contiguous bytes per query = two boxes
          if (l0_ok && !RCB_SKIP(p, 8)) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              if (x0 + 8 * half >= p.W) break;  // warp-uniform: these tiles do not exist
#pragma unroll
              for (int c = 0; c < 8; ++c) {  // chunk c = tile (c >> 2) of this half, tile row (c & 3)
                const int col = 8 * half + 4 * (c >> 2);
                const int r = c & 3;
                const float* src = (r < 2) ? va : vb;
                const int o = (r & 1) * 16 + col;
                *reinterpret_cast<float4*>(stg0 + half * 4096 + lane * 128 + ((c ^ swz) << 4)) =
                    make_float4(src[o], src[o + 1], src[o + 2], src[o + 3]);
              }
            }
          }
          if (l1_ok) {  // level-1 block 4 x 8 = two tiles = 128 bytes per query; this band owns rows 2 * band + {0, 1}
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              *reinterpret_cast<float4*>(sb1 + lane * 128 + (((t * 4 + 2 * band) ^ swz) << 4)) =
                  make_float4(l1a[4 * t], l1a[4 * t + 1], l1a[4 * t + 2], l1a[4 * t + 3]);
              *reinterpret_cast<float4*>(sb1 + lane * 128 + (((t * 4 + 2 * band + 1) ^ swz) << 4)) =
                  make_float4(l1b[4 * t], l1b[4 * t + 1], l1b[4 * t + 2], l1b[4 * t + 3]);
            }
          }
        } else {
          // ---- fp16 pyramid (tiles of 4 rows x 8 halfs): the band is two tiles = 128 bytes per query = ONE box; the
          // means are formed from the fp32 values and rounded once
          if (l0_ok && !RCB_SKIP(p, 8)) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {  // chunk c = tile (c >> 2), tile row (c & 3): 8 halfs
              const int r = c & 3;
              const float* src = (r < 2) ? va : vb;
              const int o = (r & 1) * 16 + 8 * (c >> 2);
              *reinterpret_cast<uint4*>(stg0 + lane * 128 + ((c ^ swz) << 4)) =
                  pack_half8(src[o], src[o + 1], src[o + 2], src[o + 3], src[o + 4], src[o + 5], src[o + 6], src[o + 7]);
            }
          }
          if (l1_ok) {
            // level-1 block 4 x 8 halfs = ONE tile = 64 bytes per query: SWIZZLE_64B box [32 queries][64 B], 16-byte
            // chunk r of row `lane` sits at chunk r ^ ((lane >> 1) & 3); this band owns rows 2 * band + {0, 1}
            const int s64 = (lane >> 1) & 3;
            *reinterpret_cast<uint4*>(sb1 + lane * 64 + (((2 * band) ^ s64) << 4)) =
                pack_half8(l1a[0], l1a[1], l1a[2], l1a[3], l1a[4], l1a[5], l1a[6], l1a[7]);
            *reinterpret_cast<uint4*>(sb1 + lane * 64 + (((2 * band + 1) ^ s64) << 4)) =
                pack_half8(l1b[0], l1b[1], l1b[2], l1b[3], l1b[4], l1b[5], l1b[6], l1b[7]);
          }
        }
        // band 0 hands its (fp32) level-2 row to its partner, which writes levels 2 and 3 for both
        if (band == 0) *reinterpret_cast<float4*>(xch + lane * 16) = make_float4(l2[0], l2[1], l2[2], l2[3]);
        PH_MARK(5);
        // one proxy fence per tile and warp, then the pair meets: band 0 issues the level-1 box both have written
        fence_proxy_async_smem();
        __syncwarp();
        PH_MARK(6);
        named_bar_sync(bar_done, 64);
        PH_MARK(7);
        if (lane == 0) {
          if (l0_ok && !RCB_SKIP(p, 1)) {
            if (!p.f16) {
              tma_store_4d(&map_l0, smem_u32(stg0), (x0 >> 2) * 16, yy >> 2, q_w, sg.b);
              if (x0 + 8 < p.W) tma_store_4d(&map_l0, smem_u32(stg0 + 4096), ((x0 >> 2) + 2) * 16, yy >> 2, q_w, sg.b);
            } else {
              tma_store_4d(&map_l0, smem_u32(stg0), (x0 >> 3) * 16, yy >> 2, q_w, sg.b);
            }
          }
          if (band == 0 && l1_ok && !RCB_SKIP(p, 2))
            tma_store_4d(&map_l1, smem_u32(sb1), (p.f16 ? (x1 >> 3) : (x1 >> 2)) * 16, y1 >> 2, q_w, sg.b);
          tma_store_commit();
        }
        __syncwarp();
        PH_MARK(8);
        if (band == 1 && p.levels > 2 && q_ok && !RCB_SKIP(p, 4)) {
          // levels 2 and 3 of the patch.  Level 2: rows y2 - 1 (band 0) and y2 of the 2 x 4 block are 32 contiguous
          // bytes of one fp32 tile -- a full sector per query; level 3 = mean of the block, two values.
          const float4 u = *reinterpret_cast<const float4*>(xch + lane * 16);
          if ((y2 - 1) < p.Hl[2] && x2 < p.Wl[2]) {
            if (!p.f16) {
              float* t2 = p.pyr[2] + bq * p.ps[2] + tile_off(y2 - 1, x2, p.tx[2]);
              *reinterpret_cast<float4*>(t2) = u;
              *reinterpret_cast<float4*>(t2 + 4) = make_float4(l2[0], l2[1], l2[2], l2[3]);
            } else {  // tiles of 4 rows x 8 halfs: the two rows are 8 bytes each, 16 bytes apart
              __half* t2 = reinterpret_cast<__half*>(p.pyr[2]) + bq * p.ps[2] + tile_off_h(y2 - 1, x2, p.tx[2]);
              const __half2 a0 = __floats2half2_rn(u.x, u.y), a1 = __floats2half2_rn(u.z, u.w);
              const __half2 b0 = __floats2half2_rn(l2[0], l2[1]), b1 = __floats2half2_rn(l2[2], l2[3]);
              *reinterpret_cast<uint2*>(t2) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1));
              *reinterpret_cast<uint2*>(t2 + 8) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1));
            }
          }
          const int y3 = y0 >> 3, x3 = x0 >> 3;  // x3 is even: both values sit in one tile row
          if (p.levels > 3 && y3 < p.Hl[3] && x3 < p.Wl[3]) {
            const float a = ((u.x + u.y) + (l2[0] + l2[1])) * 0.25f;
            const float c = ((u.z + u.w) + (l2[2] + l2[3])) * 0.25f;
            if (!p.f16)
              *reinterpret_cast<float2*>(p.pyr[3] + bq * p.ps[3] + tile_off(y3, x3, p.tx[3])) = make_float2(a, c);
            else
              *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p.pyr[3]) + bq * p.ps[3] + tile_off_h(y3, x3, p.tx[3])) =
                  __floats2half2_rn(a, c);
          }
        }
        PH_MARK(9);
      }
    }
#ifdef RCB_DEBUG
    if (p.prof && lane == 0 && (ew == 2 || ew == 6)) {  // warps 4 / 8 = TMEM lanes 0-31, band 0 / band 1
      const int o = ew == 2 ? 7 : 11;
      p.prof[blockIdx.x * 16 + o] = clock64() - clk0;
      p.prof[blockIdx.x * 16 + o + 1] = w0;
      p.prof[blockIdx.x * 16 + o + 2] = w1;
      p.prof[blockIdx.x * 16 + o + 3] = tile;
      // phase timers of this warp behind the 16 x 148 role counters
      for (int i = 0; i < 10; ++i) p.prof[16 * 148 + (blockIdx.x * 2 + (ew == 2 ? 0 : 1)) * 10 + i] = ph_t[i];
    }
    (void)w2; (void)w3;
#endif
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace tc

static int padded(int C, int to) { return (C + to - 1) / to * to; }

// Operand workspace: [A image | B tile images], each part 256-byte aligned
struct WsLayout {
  int K16, K8, parts16, a_words, nbox, pcols, prows;
  size_t a_bytes, b_bytes, total;
};
static WsLayout ws_layout(int B, int C, int H, int W, int mode) {
  WsLayout w{};
  w.K16 = padded(C, 64);
  w.K8 = mode == RCB_BUILD_F16F8 ? padded(C, 128) : 0;
  w.parts16 = mode == RCB_BUILD_BF16X3 ? 2 : 1;
  w.a_words = w.parts16 * w.K16 / 2 + w.K8 / 2;
  w.nbox = w.parts16 * w.K16 / 64 + 2 * (w.K8 / 128);
  w.pcols = (W + tc::PW - 1) / tc::PW;
  w.prows = (H + tc::PH - 1) / tc::PH;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  w.a_bytes = up((size_t)B * H * W * w.a_words * 4);
  w.b_bytes = up((size_t)B * w.pcols * w.prows * w.nbox * tc::BOX_BYTES);
  w.total = w.a_bytes + w.b_bytes;
  return w;
}

size_t build_tc_workspace_bytes(int B, int C, int H, int W, int mode) { return ws_layout(B, C, H, W, mode).total; }

// MMA schedule of a tile = the order of the boxes in a tile image; fills p.box / p.nbox and the pack's box tables
static void make_schedule(const WsLayout& wl, int mode, tc::Params& p, tc::PackParams& pk) {
  using namespace tc;
  for (int i = 0; i < 4; ++i) pk.box16_hi[i] = pk.box16_lo[i] = -1;
  for (int i = 0; i < 2; ++i) pk.box8_hi[i] = pk.box8_lo[i] = -1;
  int n = 0;
  const int nk16 = wl.K16 / 64;
  if (mode == RCB_BUILD_F16F8) {
    const int col_hi8 = wl.K16 / 2, col_lo8 = col_hi8 + wl.K8 / 4;
    for (int k8 = 0; k8 < wl.K8 / 128; ++k8) {  // cross terms first, accumulated at a scale of 2^12
      pk.box8_hi[k8] = n;
      p.box[n++] = BoxDesc{col_lo8 + k8 * 32, -1, KIND_F8, make_idesc(0), 0};  // e4m3(B_hi) x A_lo
      pk.box8_lo[k8] = n;
      p.box[n++] = BoxDesc{col_hi8 + k8 * 32, -1, KIND_F8, make_idesc(0), 0};  // B_lo x e4m3(A_hi)
    }
    for (int kb = 0; kb < nk16; ++kb) {
      pk.box16_hi[kb] = n;
      p.box[n++] = BoxDesc{kb * 32, -1, KIND_F16, make_idesc(0), kb == 0 ? 1 : 0};
    }
  } else if (mode == RCB_BUILD_BF16X3) {
    for (int kb = 0; kb < nk16; ++kb) {
      pk.box16_hi[kb] = n;
      p.box[n++] = BoxDesc{kb * 32, wl.K16 / 2 + kb * 32, KIND_F16, make_idesc(1), 0};
      pk.box16_lo[kb] = n;
      p.box[n++] = BoxDesc{kb * 32, -1, KIND_F16, make_idesc(1), 0};
    }
  } else {
    for (int kb = 0; kb < nk16; ++kb) {
      pk.box16_hi[kb] = n;
      p.box[n++] = BoxDesc{kb * 32, -1, KIND_F16, make_idesc(1), 0};
    }
  }
  p.nbox = n;
}

static int check_ws(const WsLayout& wl, const void* ws, size_t ws_bytes) {
  if (wl.a_words > 256 || wl.nbox > tc::MAX_BOX) return RCB_ERR_UNSUPPORTED;  // A stays resident in tensor memory (C <= 256)
  if (!ws || ws_bytes < wl.total || (reinterpret_cast<uintptr_t>(ws) & 127)) return RCB_ERR_WORKSPACE;
  return RCB_OK;
}

// pack: fp32 NCHW -> A image (queries) and B tile images (targets) in `ws`
int launch_pack_tc(const float* f1, const float* f2, int B, int C, int H, int W, int mode, void* ws, size_t ws_bytes,
                   cudaStream_t s) {
  using namespace tc;
  const WsLayout wl = ws_layout(B, C, H, W, mode);
  int st = check_ws(wl, ws, ws_bytes);
  if (st != RCB_OK) return st;
  Params p{};
  PackParams pk{};
  make_schedule(wl, mode, p, pk);
  const bool f8 = mode == RCB_BUILD_F16F8;
  pk.a_img = static_cast<unsigned char*>(ws);
  pk.a_row = (long long)wl.a_words * 4;
  pk.a_lo16 = mode == RCB_BUILD_BF16X3 ? wl.K16 * 2 : -1;
  pk.a_hi8 = wl.K16 * 2;
  pk.a_lo8 = wl.K16 * 2 + wl.K8;
  pk.b_img = pk.a_img + wl.a_bytes;
  pk.nbox = p.nbox;
  pk.pcols = wl.pcols;
  pk.npatch = wl.pcols * wl.prows;
  pk.Hp = wl.prows * PH;
  pk.Wp = wl.pcols * PW;
  pk.K16 = wl.K16;
  pk.K8 = wl.K8;
  dim3 grid((pk.Hp * pk.Wp + 63) / 64, (f8 ? wl.K8 : wl.K16) / 64, 2 * B);
  if (f8) pack_operands_kernel<true><<<grid, 256, 0, s>>>(f1, f2, pk, B, C, H, W);
  else pack_operands_kernel<false><<<grid, 256, 0, s>>>(f1, f2, pk, B, C, H, W);
  return launch_status();
}

// main kernel on operands packed by launch_pack_tc (same B, C, H, W, mode)
int launch_build_tc_packed(const void* ws, size_t ws_bytes, void* const* pyr, const rcb_pyramid_layout& lay, int B,
                           int C, int H, int W, int mode, cudaStream_t s) {
  using namespace tc;
  const bool f16 = lay.dtype == RCB_F16;
  const int esize = f16 ? 2 : 4;
  const WsLayout wl = ws_layout(B, C, H, W, mode);
  int st = check_ws(wl, ws, ws_bytes);
  if (st != RCB_OK) return st;
  if (!encode_fn()) return RCB_ERR_NO_DEVICE;
  const int Q = H * W;
  const unsigned char* a_img = static_cast<const unsigned char*>(ws);
  const unsigned char* b_img = a_img + wl.a_bytes;
  Params p{};
  {
    PackParams pk{};
    make_schedule(wl, mode, p, pk);
  }

  CUtensorMap map_l0, map_l1;
  // stores: a level is viewed as [B][Q][tile rows][tiles_x * 16 words]; one box = 128 B (two fp32 tiles) x 32 queries
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

  // work decomposition
  p.B = B; p.C = C; p.H = H; p.W = W; p.Q = Q;
  p.levels = lay.levels;
  p.mtiles = (Q + BM - 1) / BM;
  p.pcols = (W + PW - 1) / PW;
  p.prows = (H + PH - 1) / PH;
  p.npatch = p.pcols * p.prows;
  p.scale = 1.0f / sqrtf((float)C);
  for (int l = 0; l < RCB_MAX_LEVELS; ++l) {
    p.pyr[l] = l < lay.levels ? static_cast<float*>(pyr[l]) : nullptr;
    p.Hl[l] = lay.H[l]; p.Wl[l] = lay.W[l]; p.tx[l] = lay.tiles_x[l]; p.ps[l] = lay.plane_stride[l];
  }
  p.f16 = f16 ? 1 : 0;
  p.a_words = wl.a_words;
  p.a_pack = reinterpret_cast<const uint32_t*>(a_img);
  p.b_img = b_img;
  p.acc_col0 = (wl.a_words + 127) / 128 * 128;  // A occupies the first a_words TMEM columns
  p.nacc = (512 - p.acc_col0) / BN < MAX_ACC ? (512 - p.acc_col0) / BN : MAX_ACC;
  p.nacc -= p.nacc % NUM_ISSUERS;  // every issuer keeps to its own accumulators (and sees all of their phases)
#ifdef RCB_DEBUG
  p.prof = debug_env_ptr("RCB_TC_PROF_PTR");  // device buffer of 16 x 148 uint64 supplied by tools/time_build.py
  p.debug_skip = debug_env_int("RCB_TC_DEBUG_SKIP", 0);
#endif

  const int stg_total = NUM_EPI_WARPS * STG0_BYTES + 4 * STG1_BYTES + 4 * XCH_BYTES;
  int nstage = (SMEM_BUDGET - BAR_BYTES - stg_total) / STAGE_BYTES;  // 9 boxes of 16 KB
  if (nstage > MAX_STAGE) nstage = MAX_STAGE;
  {
    const int ns = debug_env_int("RCB_TC_NSTAGE", nstage);
    if (ns >= 2 && ns < nstage) nstage = ns;
  }
  if (nstage < 2) return RCB_ERR_UNSUPPORTED;
  p.nstage = nstage;
  p.b_off = 0;
  p.stg0_off = p.b_off + nstage * STAGE_BYTES;
  p.stg1_off = p.stg0_off + NUM_EPI_WARPS * STG0_BYTES;
  p.xch_off = p.stg1_off + 4 * STG1_BYTES;
  p.bar_off = p.xch_off + 4 * XCH_BYTES;
  const int smem_total = p.bar_off + BAR_BYTES;

  // one CTA per SM.  Units (batch, query tile) go round-robin over the CTAs; what is left after the full rounds is
  // cut along the patch sequence so that the last round is short instead of mostly idle.
  const int units = B * p.mtiles;
  int ctas = kNumSMs;
  if (ctas > units * p.npatch) ctas = units * p.npatch;
  p.full_rounds = units / ctas;
  const int left = units % ctas;
  // (pieces are whole patch rows -- row PAIRS in the paired sweep order: the level-3 epilogue pairs horizontally
  // adjacent patches and level-2 lines are completed by 2 x 2 patches)
  const int rows_per = RCB_PAIR_ORDER ? 2 : 1, rgroups = (p.prows + rows_per - 1) / rows_per;
  p.tail_split = left ? ctas / left : 1;
  if (p.tail_split > rgroups) p.tail_split = rgroups;
  p.tail_len = (rgroups + p.tail_split - 1) / p.tail_split * rows_per * p.pcols;
  p.tail_split = (p.npatch + p.tail_len - 1) / p.tail_len;
  p.tail_pieces = left * p.tail_split;

  cudaError_t e = cudaFuncSetAttribute(build_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
  if (e != cudaSuccess) return (int)e;
  build_tc_kernel<<<ctas, THREADS, smem_total, s>>>(map_l0, map_l1, p);
  return launch_status();
}

int launch_build_tc(const float* f1, const float* f2, void* const* pyr, const rcb_pyramid_layout& lay, int B,
                    int C, int H, int W, int mode, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int st = launch_pack_tc(f1, f2, B, C, H, W, mode, ws, ws_bytes, s);
  if (st != RCB_OK) return st;
  return launch_build_tc_packed(ws, ws_bytes, pyr, lay, B, C, H, W, mode, s);
}

}  // namespace rcb
