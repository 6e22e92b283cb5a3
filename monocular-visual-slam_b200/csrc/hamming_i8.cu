// hamming_i8.cu — K2: Hamming kNN-2 + column minimum on the 5th-gen tensor cores.
//
// Same contract as K1 (hamming_popc.cu); replaces the same reference calls
// (/root/reference/feature_pipeline.py.bak:68,82,84; homography.py:12-15,21-23).
//
// Hamming as a dense contraction whose result already IS the sort key.  Query bits map to
// +-64 (bit b -> 64(1-2b)), train bits to -+64, so a 256-byte dot product is
// 8192*ham - 2^20 exactly.  One more K-step multiplies a constant "ones" slice [1, 64, 0..]
// of the A-side tile with an "index" slice [j & 63, j >> 6, 0..] of the B-side tile
// (j = row mod 8192), adding the row index:   acc = 8192*ham + j - 2^20.
// Signed min over accumulators is therefore the lexicographic (distance, index) minimum —
// OpenCV's tie rule — and the epilogue needs no key arithmetic at all, only min/max.
// tcgen05.mma kind::i8 (M=128, N=128, K=32 per instruction, int32 accumulators in TMEM)
// produces a 128x128 tile in 9 instructions.  Indices wrap every 8192 rows, so running
// minima are flushed into the global packed-key domain at 8192-row window boundaries.
//
// Two products per tile pair so that BOTH reductions are per-thread (a TMEM lane is a
// matrix row and each epilogue thread owns one lane):
//     D1 = Qtile . Ttile^T   lanes = queries,  columns = train rows  -> per-row top-2
//     D2 = Ttile . Qtile^T   lanes = train rows, columns = queries   -> per-column minimum
// Both read the same two shared-memory tiles (K-major, canonical no-swizzle core-matrix
// layout), only the A/B descriptor roles swap.
//
// Pipeline (persistent CTAs, one per SM, 320 threads; work item = (pair, 128-query tile)):
//   warp 0   producer : cp.async.bulk (TMA engine) of pre-expanded 40 KB operand tiles into
//                       a 3-stage ring, mbarrier expect_tx / complete_tx
//   warp 1   MMA      : one thread issues 18 tcgen05.mma per train tile, tcgen05.commit
//                       releases the smem stage and publishes the TMEM accumulator stage
//   warps 2-9 epilogue: tcgen05.ld 32x32b.x32 (thread = TMEM lane), 4-way interleaved
//                       top-2 chains at 2.5 min/max per element (pairs + VIMNMX3), 3-input
//                       mins for the column direction, atomicMin for columns
// TMEM: 2 accumulator stages x (D1 128 cols + D2 128 cols) = 512 columns.
#include "common.cuh"

namespace b2s {

constexpr int kI8Tile = 128;                      // rows per operand tile
constexpr int kI8Chunks = 20;                     // 16-byte k-chunks per row: 16 data + ones(2) + index(2)
constexpr int kI8ChunkBytes = kI8Tile * 16;       // 2048: one k-chunk of all 128 rows (= LBO)
constexpr int kI8TileBytes = kI8Chunks * kI8ChunkBytes;  // 40 KB
constexpr int kI8Units = kI8Chunks * kI8Tile;     // 16-byte units per tile
constexpr int kI8Stages = 3;
constexpr int kI8Threads = 320;
constexpr int kWinBits = 13;                      // index window: 8192 rows
constexpr int kWin = 1 << kWinBits;
constexpr int kAccBias = 1 << 20;                 // acc + 2^20 = ham << 13 | (row mod 8192)
constexpr int kAccNone = 0x7FFFFFFF;

// ---- pre-pass: 256 bits -> 256 int8 (+-64) + ones/index slices, UMMA canonical K-major layout
// Tile = 128 rows x 20 k-chunks of 16 bytes.  Unit (row r, k-chunk kc) sits at unit index
// kc*128 + r: 8 rows x 16 B form one 128-byte core matrix, 8-row groups are 128 B apart
// (SBO), k-chunks 2048 B apart (LBO).
__device__ __forceinline__ uint32_t spread4(uint32_t nib, bool train) {
  const uint32_t sp = (nib * 0x00204081u) & 0x01010101u;  // bit i -> byte i (0 or 1)
  // query: +64 / -64 for clear / set;  train: -64 / +64
  return train ? (0xC0C0C0C0u - sp * 0x80u) : (0x40404040u + sp * 0x80u);
}

__global__ void __launch_bounds__(256) expand_pm64_kernel(const uint8_t* __restrict__ desc,
                                                          const int32_t* __restrict__ off,
                                                          const int32_t* __restrict__ src, int tiles_per_pair,
                                                          int train, uint4* __restrict__ out) {
  const int pair = blockIdx.y;
  const int o = off[pair];
  const int n = off[pair + 1] - o;
  const int in0 = src ? src[pair] : o;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;  // 16-byte unit within the pair's tiles
  const int tile = u / kI8Units, w = u - tile * kI8Units;
  if (tile >= tiles_per_pair) return;
  if (tile * kI8Tile >= n) return;  // tile never read
  const int kc = w >> 7, r = w & 127;
  const int row = tile * kI8Tile + r;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (kc < 16) {
    if (row < n) {
      const uint32_t bits =
          *reinterpret_cast<const uint16_t*>(desc + (size_t)(in0 + row) * B2S_DESC_BYTES + 2 * kc);
      v.x = spread4(bits & 15u, train);
      v.y = spread4((bits >> 4) & 15u, train);
      v.z = spread4((bits >> 8) & 15u, train);
      v.w = spread4((bits >> 12) & 15u, train);
    }
  } else if (kc == 16) {
    v.x = 1u | (64u << 8);                               // ones slice: [1, 64, 0, ...]
  } else if (kc == 18) {
    const uint32_t j = (uint32_t)row & (uint32_t)(kWin - 1);
    v.x = (j & 63u) | ((j >> 6) << 8);                   // index slice: [j & 63, j >> 6, 0, ...]
  }
  out[((size_t)pair * tiles_per_pair + tile) * kI8Units + w] = v;
}

// ---- tcgen05 wrappers ------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  // try_wait suspends in hardware for a bounded time; a broken pipeline traps instead of hanging
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, LBO = 2048 B, SBO = 128 B, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)(kI8ChunkBytes >> 4) << 16; // leading (K-direction) byte offset, bits [16,30)
  d |= (uint64_t)(128u >> 4) << 32;          // stride (8-row group) byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  return d;                                  // base offset 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor: S32 accumulate, signed 8-bit A and B, both K-major, M = 128, N = 128
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

#define TMEM_LD_X32(taddr, v)                                                                              \
  asm volatile(                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),    \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
      : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct I8Params {
  const uint8_t* __restrict__ qx;  // expanded query tiles  [pair][q_tiles][40 KB]
  const uint8_t* __restrict__ tx;  // expanded train tiles  [pair][t_tiles][40 KB]
  const int32_t* __restrict__ q_off;
  const int32_t* __restrict__ t_off;
  uint32_t* __restrict__ fwd_best;
  uint32_t* __restrict__ fwd_second;
  uint32_t* __restrict__ bwd_best;
  int q_tiles, t_tiles, n_pairs;
  unsigned long long* dbg;  // optional per-CTA stall counters (b2s_hamming_i8_debug), else nullptr
};
// dbg layout per CTA (8 x u64): [0] MMA thread total, [1] MMA wait tempty, [2] MMA wait full/qfull,
// [3] producer wait empty, [4] epilogue warp 2 total, [5] epilogue wait tfull, [6] tile pairs, [7] -

// 32 accumulator columns folded into 4 interleaved top-2 chains (ILP 4; one chain would
// serialise on the 4-cycle ALU latency).  Two columns per step: the two smallest of
// {best, second, lo, hi} are min(best, lo) and min3(second, max(best, lo), hi) — 5 ops / 2.
template <bool TAIL>
__device__ __forceinline__ void fold_top2(const uint32_t (&v)[32], int c0, int valid, int (&b)[4], int (&s)[4]) {
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    int x0 = (int)v[k], x1 = (int)v[k + 1];
    if (TAIL) {
      x0 = (c0 + k < valid) ? x0 : kAccNone;
      x1 = (c0 + k + 1 < valid) ? x1 : kAccNone;
    }
    const int c = (k >> 1) & 3;
    const int lo = min(x0, x1), hi = max(x0, x1);
    const int mb = max(b[c], lo);
    b[c] = min(b[c], lo);
    s[c] = min(min(s[c], mb), hi);
  }
}
template <bool TAIL>
__device__ __forceinline__ void fold_min(const uint32_t (&v)[32], int c0, int valid, int (&m)[4]) {
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    int x0 = (int)v[k], x1 = (int)v[k + 1];
    if (TAIL) {
      x0 = (c0 + k < valid) ? x0 : kAccNone;
      x1 = (c0 + k + 1 < valid) ? x1 : kAccNone;
    }
    const int c = (k >> 1) & 3;
    m[c] = min(min(m[c], x0), x1);  // VIMNMX3
  }
}
// accumulator domain -> packed key (distance << 22 | pair-local index)
__device__ __forceinline__ uint32_t acc_to_key(int acc, int window_base) {
  const uint32_t u = (uint32_t)(acc + kAccBias);
  return ((u >> kWinBits) << kIdxBits) | ((u & (uint32_t)(kWin - 1)) + (uint32_t)window_base);
}

// Persistent: gridDim.x CTAs (one per SM) walk the (pair, query tile) work items round-robin.
// TMEM, barriers and the train-tile ring are set up once; the ring and the accumulator stages
// keep cycling across work items (global tile counter g), and the query tile is double
// buffered, so the pipeline never drains between items.  (The first version launched one CTA
// per item: ~36k clk per item against 19.6k clk of MMA time — prologue, first-load latency
// and drain were un-overlapped because 200 KB of shared memory allow one CTA per SM.)
__global__ void __launch_bounds__(kI8Threads, 1) hamming_knn2_i8_kernel(const I8Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_q = smem;                                   // 2 x 40 KB (double-buffered query tile)
  uint8_t* s_t = smem + 2 * kI8TileBytes;                // kI8Stages x 40 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + kI8Stages) * kI8TileBytes);
  uint64_t* b_full = bars;                               // [kI8Stages]
  uint64_t* b_empty = bars + kI8Stages;                  // [kI8Stages]
  uint64_t* b_tfull = bars + 2 * kI8Stages;              // [2]
  uint64_t* b_tempty = bars + 2 * kI8Stages + 2;         // [2]
  uint64_t* b_qfull = bars + 2 * kI8Stages + 4;          // [2]
  uint64_t* b_qempty = bars + 2 * kI8Stages + 6;         // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kI8Stages + 8);
  uint2* s_merge = reinterpret_cast<uint2*>(bars + 2 * kI8Stages + 9);  // [128] fwd halves meet here

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_pairs * p.q_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kI8Stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&b_tfull[a], 1);
      mbar_init(&b_tempty[a], 8);  // one elected arrive per epilogue warp
      mbar_init(&b_qfull[a], 1);
      mbar_init(&b_qempty[a], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns (1 CTA per SM by shared-memory footprint)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // every role walks the same item sequence; items past a pair's last query tile are skipped
  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      uint32_t n = 0, g = 0;  // items seen by this CTA, train tiles streamed by this CTA
      long long w_empty = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int pair = w / p.q_tiles, qt = w - pair * p.q_tiles;
        const int nq = p.q_off[pair + 1] - p.q_off[pair];
        if (qt * kI8Tile >= nq) continue;
        const int nt = p.t_off[pair + 1] - p.t_off[pair];
        const int n_tt = (nt + kI8Tile - 1) / kI8Tile;
        const uint32_t qb = n & 1u;
        if (n >= 2) mbar_wait_bounded(&b_qempty[qb], ((n >> 1) - 1u) & 1u);
        mbar_arrive_expect_tx(&b_qfull[qb], kI8TileBytes);
        bulk_g2s(s_q + (size_t)qb * kI8TileBytes, p.qx + ((size_t)pair * p.q_tiles + qt) * kI8TileBytes, kI8TileBytes,
                 &b_qfull[qb]);
        const uint8_t* tsrc = p.tx + (size_t)pair * p.t_tiles * kI8TileBytes;
        for (int t = 0; t < n_tt; ++t, ++g) {
          const uint32_t s = g % kI8Stages;
          const long long c0 = p.dbg ? clock64() : 0;
          if (g >= kI8Stages) mbar_wait_bounded(&b_empty[s], ((g / kI8Stages) - 1u) & 1u);
          if (p.dbg) w_empty += clock64() - c0;
          mbar_arrive_expect_tx(&b_full[s], kI8TileBytes);
          bulk_g2s(s_t + (size_t)s * kI8TileBytes, tsrc + (size_t)t * kI8TileBytes, kI8TileBytes, &b_full[s]);
        }
        ++n;
      }
      if (p.dbg) p.dbg[blockIdx.x * 8 + 3] = (unsigned long long)w_empty;
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t n = 0, g = 0;
      long long w_tempty = 0, w_full = 0;
      const long long t_start = p.dbg ? clock64() : 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int pair = w / p.q_tiles, qt = w - pair * p.q_tiles;
        const int nq = p.q_off[pair + 1] - p.q_off[pair];
        if (qt * kI8Tile >= nq) continue;
        const int nt = p.t_off[pair + 1] - p.t_off[pair];
        const int n_tt = (nt + kI8Tile - 1) / kI8Tile;
        const uint32_t qb = n & 1u;
        long long c0 = p.dbg ? clock64() : 0;
        mbar_wait_bounded(&b_qfull[qb], (n >> 1) & 1u);
        if (p.dbg) w_full += clock64() - c0;
        const uint64_t qdesc = make_smem_desc(smem_u32(s_q + (size_t)qb * kI8TileBytes));
        for (int t = 0; t < n_tt; ++t, ++g) {
          const uint32_t s = g % kI8Stages, a = g & 1u;
          c0 = p.dbg ? clock64() : 0;
          if (g >= 2) mbar_wait_bounded(&b_tempty[a], ((g >> 1) - 1u) & 1u);
          long long c1 = p.dbg ? clock64() : 0;
          mbar_wait_bounded(&b_full[s], (g / kI8Stages) & 1u);
          if (p.dbg) {
            w_tempty += c1 - c0;
            w_full += clock64() - c1;
          }
          tc_fence_after();
          const uint64_t tdesc = make_smem_desc(smem_u32(s_t + (size_t)s * kI8TileBytes));
          const uint32_t d1 = tmem_base + a * 256u, d2 = d1 + 128u;
          // K = 32 bytes per instruction = two k-chunks = 4096 B apart (descriptor units of 16 B: 256).
          // Step 8 pairs the A tile's ones slice (chunk 16) with the B tile's index slice (chunk 18).
          constexpr uint64_t kOnes = 16u * (kI8ChunkBytes >> 4), kIndex = 18u * (kI8ChunkBytes >> 4);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_i8(d1, qdesc + (uint64_t)k * 256u, tdesc + (uint64_t)k * 256u, kIdescI8, k > 0);
          tc_mma_i8(d1, qdesc + kOnes, tdesc + kIndex, kIdescI8, 1);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_i8(d2, tdesc + (uint64_t)k * 256u, qdesc + (uint64_t)k * 256u, kIdescI8, k > 0);
          tc_mma_i8(d2, tdesc + kOnes, qdesc + kIndex, kIdescI8, 1);
          tc_commit(&b_empty[s]);  // smem stage reusable once these MMAs have read it
          tc_commit(&b_tfull[a]);  // accumulators complete
        }
        tc_commit(&b_qempty[qb]);  // query tile buffer reusable once this item's MMAs have read it
        ++n;
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - t_start);
        p.dbg[blockIdx.x * 8 + 1] = (unsigned long long)w_tempty;
        p.dbg[blockIdx.x * 8 + 2] = (unsigned long long)w_full;
        p.dbg[blockIdx.x * 8 + 6] = g;
      }
    }
  } else {
    // ===== epilogue: 8 warps; warp%4 = TMEM lane quarter, (warp-2)/4 = column half =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;  // TMEM lane = tile row
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t g = 0;
    long long w_tfull = 0;
    const long long e_start = p.dbg ? clock64() : 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int pair = w / p.q_tiles, qt = w - pair * p.q_tiles;
      const int qo = p.q_off[pair], nq = p.q_off[pair + 1] - qo;
      const int q0 = qt * kI8Tile;
      if (q0 >= nq) continue;
      const int to = p.t_off[pair], nt = p.t_off[pair + 1] - to;
      const int n_tt = (nt + kI8Tile - 1) / kI8Tile;
      const int nq_valid = min(kI8Tile, nq - q0);
      uint32_t gbest = kNone, gsecond = kNone;
      int b[4] = {kAccNone, kAccNone, kAccNone, kAccNone}, s2[4] = {kAccNone, kAccNone, kAccNone, kAccNone};
      const int q_window = q0 & ~(kWin - 1);
      for (int t = 0; t < n_tt; ++t, ++g) {
        const uint32_t a = g & 1u;
        const int tbase = t * kI8Tile;
        const int nt_valid = min(kI8Tile, nt - tbase);
        const long long c0 = p.dbg ? clock64() : 0;
        mbar_wait_bounded(&b_tfull[a], (g >> 1) & 1u);
        if (p.dbg) w_tfull += clock64() - c0;
        tc_fence_after();
        uint32_t v0[32], v1[32];
        // ---- D1: this thread's query row vs 64 train columns ----
        const uint32_t c1 = lane_addr + a * 256u + (uint32_t)half * 64u;
        TMEM_LD_X32(c1, v0);
        TMEM_LD_X32(c1 + 32u, v1);
        tmem_ld_wait();
        if (nt_valid == kI8Tile) {
          fold_top2<false>(v0, half * 64, kI8Tile, b, s2);
          fold_top2<false>(v1, half * 64 + 32, kI8Tile, b, s2);
        } else {
          fold_top2<true>(v0, half * 64, nt_valid, b, s2);
          fold_top2<true>(v1, half * 64 + 32, nt_valid, b, s2);
        }
        // ---- D2: this thread's train row vs 64 query columns ----
        const uint32_t c2 = c1 + 128u;
        TMEM_LD_X32(c2, v0);
        TMEM_LD_X32(c2 + 32u, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_tempty[a]);  // accumulator stage drained into registers
        int m[4] = {kAccNone, kAccNone, kAccNone, kAccNone};
        if (nq_valid == kI8Tile) {
          fold_min<false>(v0, half * 64, kI8Tile, m);
          fold_min<false>(v1, half * 64 + 32, kI8Tile, m);
        } else {
          fold_min<true>(v0, half * 64, nq_valid, m);
          fold_min<true>(v1, half * 64 + 32, nq_valid, m);
        }
        const int cm = min(min(m[0], m[1]), min(m[2], m[3]));
        if (row < nt_valid && cm != kAccNone) atomicMin(&p.bwd_best[to + tbase + row], acc_to_key(cm, q_window));
        // leave the accumulator domain when the train index window (8192 rows) ends
        if (((t + 1) & (kWin / kI8Tile - 1)) == 0 || t == n_tt - 1) {
          const int window_base = tbase & ~(kWin - 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (b[k] != kAccNone) top2_insert(gbest, gsecond, acc_to_key(b[k], window_base));
            if (s2[k] != kAccNone) top2_insert(gbest, gsecond, acc_to_key(s2[k], window_base));
            b[k] = kAccNone;
            s2[k] = kAccNone;
          }
        }
      }
      // the two column halves of a query row meet in shared memory
      if (half == 1) s_merge[row] = make_uint2(gbest, gsecond);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // epilogue warps only
      if (half == 0 && row < nq_valid) {
        const uint2 o = s_merge[row];
        top2_insert(gbest, gsecond, o.x);
        top2_insert(gbest, gsecond, o.y);
        p.fwd_best[qo + q0 + row] = gbest;
        p.fwd_second[qo + q0 + row] = gsecond;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // s_merge free for the next item
    }
    if (p.dbg && threadIdx.x == 64) {
      p.dbg[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - e_start);
      p.dbg[blockIdx.x * 8 + 5] = (unsigned long long)w_tfull;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- measurement: raw tcgen05.mma kind::i8 issue rate (no epilogue), one CTA per SM ------
// n_dim = 128 or 256; A = 128 rows, B = n_dim rows, both in the canonical no-swizzle layout.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int n_dim) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (3 * kI8TileBytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tm;
  if (warp == 1 && lane == 0) {
    const uint64_t adesc = make_smem_desc(smem_u32(smem));
    const uint64_t bdesc = make_smem_desc(smem_u32(smem + kI8TileBytes));
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)n_dim >> 3) << 17) | ((128u >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tm + (uint32_t)((it & 1) * 256);
#pragma unroll
      for (int k = 0; k < 8; ++k) tc_mma_i8(d, adesc + (uint64_t)k * 256u, bdesc + (uint64_t)k * 256u, idesc, k > 0);
    }
    tc_commit(&bar);
    mbar_wait_bounded(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

// ---- measurement: TMEM -> register read bandwidth (tcgen05.ld 32x32b.x32), `warps` warps per SM
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, uint32_t* sink) {
  __shared__ uint32_t s_tm;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tm;
  const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16);
  const int nw = blockDim.x >> 5;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    // the warps sharing a lane quarter split the 512 columns
    const int per = 512 / (nw / 4);
    const int c0 = (warp >> 2) * per;
    for (int c = c0; c < c0 + per; c += 64) {
      uint32_t v0[32], v1[32];
      TMEM_LD_X32(lane_addr + (uint32_t)c, v0);
      TMEM_LD_X32(lane_addr + (uint32_t)c + 32u, v1);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; k += 8) acc ^= v0[k] ^ v1[k];
    }
  }
  if (acc == 0x12345u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

int tmem_read_launch(int iters, int warps, double* bytes_out, uint32_t* sink, cudaStream_t st) {
  B2S_REQUIRE(warps == 4 || warps == 8 || warps == 16, "warps must be 4, 8 or 16");
  tmem_read_kernel<<<sm_count(), warps * 32, 0, st>>>(iters, sink);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  if (bytes_out) *bytes_out = (double)sm_count() * iters * 128.0 * 512.0 * 4.0;
  return B2S_OK;
}

int mma_rate_launch(int iters, int n_dim, double* macs_out, cudaStream_t st) {
  B2S_REQUIRE(n_dim == 128 || n_dim == 256, "n_dim must be 128 or 256");
  const size_t smem = 3 * (size_t)kI8TileBytes;
  B2S_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate_kernel<<<sm_count(), 128, smem, st>>>(iters, n_dim);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  if (macs_out) *macs_out = (double)sm_count() * iters * 8.0 * 128.0 * n_dim * 32.0;
  return B2S_OK;
}

static unsigned long long* g_i8_dbg = nullptr;  // device buffer, 8 u64 per SM (diagnostics only)
void hamming_i8_set_debug(unsigned long long* dev_buf) { g_i8_dbg = dev_buf; }

constexpr size_t kI8SmemBytes = (size_t)(2 + kI8Stages) * kI8TileBytes + 8 * (2 * kI8Stages + 9) + 128 * sizeof(uint2);

size_t hamming_i8_workspace_bytes(int n_pairs, int max_nq, int max_nt) {
  const size_t qt = (size_t)((max_nq + kI8Tile - 1) / kI8Tile), tt = (size_t)((max_nt + kI8Tile - 1) / kI8Tile);
  return (size_t)n_pairs * (qt + tt) * kI8TileBytes;
}

int hamming_i8_launch(const uint8_t* q, const uint8_t* t, const int32_t* q_off, const int32_t* t_off,
                      const int32_t* q_src, const int32_t* t_src, int n_pairs, int total_nq, int total_nt, int max_nq,
                      int max_nt, uint32_t* fwd_best, uint32_t* fwd_second, uint32_t* bwd_best, void* workspace,
                      size_t workspace_bytes, cudaStream_t st) {
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  if (total_nt > 0) B2S_CUDA(cudaMemsetAsync(bwd_best, 0xFF, sizeof(uint32_t) * (size_t)total_nt, st));
  if (total_nq > 0) {  // rows of pairs without train descriptors keep "none"
    B2S_CUDA(cudaMemsetAsync(fwd_best, 0xFF, sizeof(uint32_t) * (size_t)total_nq, st));
    B2S_CUDA(cudaMemsetAsync(fwd_second, 0xFF, sizeof(uint32_t) * (size_t)total_nq, st));
  }
  if (total_nq == 0 || total_nt == 0 || max_nq == 0 || max_nt == 0) return B2S_OK;
  const int qt = (max_nq + kI8Tile - 1) / kI8Tile, tt = (max_nt + kI8Tile - 1) / kI8Tile;
  const size_t need = hamming_i8_workspace_bytes(n_pairs, max_nq, max_nt);
  B2S_REQUIRE(workspace != nullptr && workspace_bytes >= need,
              "i8 variant needs %zu workspace bytes (b2s_hamming_workspace_bytes_v), got %zu", need, workspace_bytes);
  B2S_REQUIRE(((uintptr_t)workspace & 127u) == 0, "workspace must be 128-byte aligned");
  uint8_t* qx = static_cast<uint8_t*>(workspace);
  uint8_t* tx = qx + (size_t)n_pairs * qt * kI8TileBytes;
  expand_pm64_kernel<<<dim3((qt * kI8Units + 255) / 256, n_pairs), 256, 0, st>>>(q, q_off, q_src, qt, 0,
                                                                                reinterpret_cast<uint4*>(qx));
  B2S_CUDA(cudaGetLastError());
  expand_pm64_kernel<<<dim3((tt * kI8Units + 255) / 256, n_pairs), 256, 0, st>>>(t, t_off, t_src, tt, 1,
                                                                                reinterpret_cast<uint4*>(tx));
  B2S_CUDA(cudaGetLastError());
  note_launch(2);
  static bool attr_set = false;
  if (!attr_set) {
    B2S_CUDA(cudaFuncSetAttribute(hamming_knn2_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kI8SmemBytes));
    attr_set = true;
  }
  I8Params p;
  p.qx = qx;
  p.tx = tx;
  p.q_off = q_off;
  p.t_off = t_off;
  p.fwd_best = fwd_best;
  p.fwd_second = fwd_second;
  p.bwd_best = bwd_best;
  p.q_tiles = qt;
  p.t_tiles = tt;
  p.n_pairs = n_pairs;
  p.dbg = g_i8_dbg;
  const long items = (long)qt * n_pairs;
  const int grid = (int)(items < (long)sm_count() ? items : (long)sm_count());
  hamming_knn2_i8_kernel<<<grid, kI8Threads, kI8SmemBytes, st>>>(p);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

}  // namespace b2s
