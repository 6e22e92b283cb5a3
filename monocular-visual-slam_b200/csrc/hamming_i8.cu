// hamming_i8.cu — K2: Hamming kNN-2 + column minimum on the 5th-gen tensor cores.
//
// Same contract as K1 (hamming_popc.cu); replaces the same reference calls
// (/root/reference/feature_pipeline.py.bak:68,82,84; homography.py:12-15,21-23).
//
// Hamming as a dense contraction whose result already IS a 16-bit sort key.  Query bits map
// to +-8 (bit b -> 8(1-2b)), train bits to -+8, so a 256-byte dot product is 128*ham - 2^14
// exactly.  One more K-step multiplies a constant "ones" block [1, 64 x8, 0..] (A side, lives
// in shared memory for the whole kernel) with the B-side tile's "index" chunk
// [row mod 128, 64 x4, 0..], adding the tile-local column and cancelling the bias:
//     acc = 128*ham + (row mod 128)            0 <= acc <= 32895, i.e. a 16-bit key.
// Rows past the end of a pair carry zero data and the index chunk [127 x9, 0..], which yields
// acc = 65151: larger than every valid key, still 16 bits — no tail masking in the epilogue.
// Unsigned 16-bit min over keys is the lexicographic (distance, index) minimum = OpenCV's tie
// rule, so the epilogue is tcgen05.ld.pack::16b (two columns per register) followed by packed
// VIMNMX.U16x2 / VIMNMX3.U16x2 only: 1.25 instructions per element for the per-row top-2 and
// 0.25 for the column minimum (the 32-bit-key version needed 2.5 + 0.5).  Once per 128-column
// tile the two packed halves are folded and widened into the global packed key.
// tcgen05.mma kind::i8 (M=128, N=128, K=32 per instruction, int32 accumulators in TMEM)
// produces a 128x128 tile in 9 instructions.
//
// Two products per tile pair so that BOTH reductions are per-thread (a TMEM lane is a
// matrix row and each epilogue thread owns one lane):
//     D1 = Qtile . Ttile^T   lanes = queries,  columns = train rows  -> per-row top-2
//     D2 = Ttile . Qtile^T   lanes = train rows, columns = queries   -> per-column minimum
// Both read the same two shared-memory tiles (K-major, canonical no-swizzle core-matrix
// layout), only the A/B descriptor roles swap.
//
// Pipeline (persistent CTAs, one per SM, 320 threads; work item = (pair, 128-query tile)):
//   warp 0   producer : cp.async.bulk (TMA engine) of pre-expanded 34 KB operand tiles into
//                       a 4-stage ring + a double-buffered query tile, mbarrier expect_tx
//   warp 1   MMA      : one thread issues 18 tcgen05.mma per train tile, tcgen05.commit
//                       releases the smem stage and publishes the TMEM accumulator stage
//   warps 2-9 epilogue: two sets of 4 warps; set e owns TMEM stage e (every other tile pair),
//                       thread = TMEM lane, 4 x tcgen05.ld.32x32b.x32.pack::16b per tile pair
//                       software-pipelined against the packed min/max folds
// TMEM: 2 accumulator stages x (D1 128 cols + D2 128 cols) = 512 columns.
#include "common.cuh"

namespace b2s {

constexpr int kI8Tile = 128;                      // rows per operand tile
constexpr int kI8Chunks = 17;                     // 16-byte k-chunks per row: 16 data + 1 index
constexpr int kI8ChunkBytes = kI8Tile * 16;       // 2048: one k-chunk of all 128 rows (= LBO)
constexpr int kI8TileBytes = kI8Chunks * kI8ChunkBytes;  // 34 KB
constexpr int kI8Units = kI8Chunks * kI8Tile;     // 16-byte units per tile
constexpr int kI8Stages = 4;
constexpr int kI8Threads = 320;
constexpr uint32_t kKey16Valid = 32896u;          // 16-bit keys below this are real (ham <= 256, col <= 127)
constexpr uint32_t kKey32Pad = 0x7F000000u;       // widened padding / initial keys land at or above this

// ---- pre-pass: 256 bits -> 256 int8 (+-8) + index chunk, UMMA canonical K-major layout
// Tile = 128 rows x 17 k-chunks of 16 bytes.  Unit (row r, k-chunk kc) sits at unit index
// kc*128 + r: 8 rows x 16 B form one 128-byte core matrix, 8-row groups are 128 B apart
// (SBO), k-chunks 2048 B apart (LBO).
__device__ __forceinline__ uint32_t spread4(uint32_t nib, bool train) {
  const uint32_t sp = (nib * 0x00204081u) & 0x01010101u;  // bit i -> byte i (0 or 1)
  // query: +8 / -8 for clear / set;  train: -8 / +8
  return train ? (0xF8F8F8F8u - sp * 0xF0u) : (0x08080808u + sp * 0xF0u);
}

__global__ void __launch_bounds__(256) expand_pm8_kernel(const uint8_t* __restrict__ desc,
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
  } else if (row < n) {
    v.x = (uint32_t)r | 0x40404000u;                     // index chunk: [r, 64, 64, 64, 64, 0, ...]
    v.y = 0x00000040u;
  } else {
    v.x = 0x7F7F7F7Fu;                                   // padding row: [127 x9, 0, ...] -> acc = 65151
    v.y = 0x7F7F7F7Fu;
    v.z = 0x0000007Fu;
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

// The 18 MMAs of one tile pair as ONE asm block: 8 data K-steps + the index K-step into D1
// (A = query tile, B = train tile), the same into D2 with the roles swapped.  After the 14th
// MMA a non-consumed mbarrier.try_wait probes the barrier of the NEXT tile pair; its predicate
// is only read after the last MMA has been queued, so the probe's latency (and the hardware
// sleep until the phase completes) overlaps the MMA stream instead of idling the tensor pipe
// at the tile boundary (~170 clk per blocking wait, b2s_mma_microbench variants 2 vs 3).
// Returns 1 when the next tile pair's barrier phase was seen complete.
__device__ __forceinline__ uint32_t tc_mma_tile_pair(uint32_t d1, uint32_t d2, uint64_t qdesc, uint64_t tdesc,
                                                     uint64_t odesc, uint32_t idesc, uint64_t* next_bar,
                                                     uint32_t next_parity, uint32_t has_next) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, pnew, pprobe, pdone;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.eq.b32 pacc, 0, 0;\n\t"
      "setp.ne.b32 pnew, 0, 0;\n\t"
      "setp.ne.b32 pprobe, %9, 0;\n\t"
      "setp.ne.b32 pdone, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], %3, %4, %6, pnew;\n\t"
      "add.u64 da, %3, 256;\n\t"
      "add.u64 db, %4, 256;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], da, db, %6, pacc;\n\t"
      "add.u64 da, %3, 512;\n\t"
      "add.u64 db, %4, 512;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], da, db, %6, pacc;\n\t"
      "add.u64 da, %3, 768;\n\t"
      "add.u64 db, %4, 768;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], da, db, %6, pacc;\n\t"
      "add.u64 da, %3, 1024;\n\t"
      "add.u64 db, %4, 1024;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], da, db, %6, pacc;\n\t"
      "add.u64 da, %3, 1280;\n\t"
      "add.u64 db, %4, 1280;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], da, db, %6, pacc;\n\t"
      "add.u64 da, %3, 1536;\n\t"
      "add.u64 db, %4, 1536;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], da, db, %6, pacc;\n\t"
      "add.u64 da, %3, 1792;\n\t"
      "add.u64 db, %4, 1792;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], da, db, %6, pacc;\n\t"
      "add.u64 db, %4, 2048;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], %5, db, %6, pacc;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], %4, %3, %6, pnew;\n\t"
      "add.u64 da, %4, 256;\n\t"
      "add.u64 db, %3, 256;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], da, db, %6, pacc;\n\t"
      "add.u64 da, %4, 512;\n\t"
      "add.u64 db, %3, 512;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], da, db, %6, pacc;\n\t"
      "add.u64 da, %4, 768;\n\t"
      "add.u64 db, %3, 768;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], da, db, %6, pacc;\n\t"
      "add.u64 da, %4, 1024;\n\t"
      "add.u64 db, %3, 1024;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], da, db, %6, pacc;\n\t"
      "@pprobe mbarrier.try_wait.parity.shared::cta.b64 pdone, [%7], %8;\n\t"
      "add.u64 da, %4, 1280;\n\t"
      "add.u64 db, %3, 1280;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], da, db, %6, pacc;\n\t"
      "add.u64 da, %4, 1536;\n\t"
      "add.u64 db, %3, 1536;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], da, db, %6, pacc;\n\t"
      "add.u64 da, %4, 1792;\n\t"
      "add.u64 db, %3, 1792;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], da, db, %6, pacc;\n\t"
      "add.u64 db, %3, 2048;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], %5, db, %6, pacc;\n\t"
      "selp.u32 %0, 1, 0, pdone;\n\t"
      "}"
      : "=r"(ok)
      : "r"(d1), "r"(d2), "l"(qdesc), "l"(tdesc), "l"(odesc), "r"(idesc), "r"(smem_u32(next_bar)), "r"(next_parity),
        "r"(has_next)
      : "memory");
  return ok;
}


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

// 64 accumulator columns -> 32 registers: register j = low16(col 2j+1) << 16 | low16(col 2j)
#define TMEM_LD_X32P(taddr, v)                                                                             \
  asm volatile(                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "                                                  \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),    \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
      : "r"(taddr))

// Empty asm that "rewrites" the 32 registers: ordered after the volatile tcgen05.wait::ld, it
// keeps the compiler from hoisting arithmetic on freshly loaded registers above the wait.
#define TMEM_REGS_READY(v)                                                                                 \
  asm volatile(""                                                                                          \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),       \
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),   \
                 "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]),             \
                 "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]),             \
                 "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]))

struct I8Params {
  const uint8_t* __restrict__ qx;  // expanded query tiles  [pair][q_tiles][34 KB]
  const uint8_t* __restrict__ tx;  // expanded train tiles  [pair][t_tiles][34 KB]
  const int32_t* __restrict__ q_off;
  const int32_t* __restrict__ t_off;
  uint32_t* __restrict__ fwd_best;
  uint32_t* __restrict__ fwd_second;
  uint32_t* __restrict__ bwd_best;
  int q_tiles, t_tiles, n_pairs;
  unsigned long long* dbg;  // optional per-CTA stall counters (b2s_hamming_i8_debug), else nullptr
  int mode;                 // diagnostics only: bit 0 = epilogue does no work, bit 1 = ring is loaded once
};
// dbg layout per CTA (8 x u64): [0] MMA thread total, [1] MMA wait tempty, [2] MMA wait full/qfull,
// [3] producer wait empty, [4] epilogue warp 2 total, [5] epilogue wait tfull, [6] tile pairs, [7] -

// 64 columns (32 packed registers) folded into 2 interleaved packed top-2 chains.  The low and
// the high 16-bit lane of every register are independent streams (even / odd columns).  Two
// registers per step: the two smallest of {best, second, lo, hi} are min(best, lo) and
// min3(second, max(best, lo), hi) — 5 instructions per 4 columns.
__device__ __forceinline__ void fold_top2_p16(const uint32_t (&v)[32], uint32_t (&b)[2], uint32_t (&s)[2]) {
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    const int c = (k >> 1) & 1;
    const uint32_t lo = __vminu2(v[k], v[k + 1]), hi = __vmaxu2(v[k], v[k + 1]);
    const uint32_t mb = __vmaxu2(b[c], lo);
    b[c] = __vminu2(b[c], lo);
    s[c] = __vimin3_u16x2(s[c], mb, hi);
  }
}
__device__ __forceinline__ void fold_min_p16(const uint32_t (&v)[32], uint32_t (&m)[2]) {
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    const int c = (k >> 1) & 1;
    m[c] = __vimin3_u16x2(m[c], v[k], v[k + 1]);
  }
}
__device__ __forceinline__ uint32_t swap16(uint32_t x) { return __byte_perm(x, 0, 0x1032); }
// tile-local 16-bit key (ham << 7 | col) -> packed key (ham << 22 | base + col); padding keys
// (>= kKey16Valid) land at or above kKey32Pad and are turned into "none" by the caller
__device__ __forceinline__ uint32_t key16_to_key32(uint32_t k16, uint32_t base) {
  return ((k16 & 0xFF80u) << (kIdxBits - 7)) | (base + (k16 & 127u));
}

// Persistent: gridDim.x CTAs (one per SM) walk the (pair, query tile) work items round-robin.
// TMEM, barriers and the train-tile ring are set up once; the ring and the accumulator stages
// keep cycling across work items (global tile counter g), and the query tile is double
// buffered, so the pipeline never drains between items.
__global__ void __launch_bounds__(kI8Threads, 1) hamming_knn2_i8_kernel(const I8Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_q = smem;                                   // 2 x 34 KB (double-buffered query tile)
  uint8_t* s_t = smem + 2 * kI8TileBytes;                // kI8Stages x 34 KB
  uint8_t* s_ones = smem + (2 + kI8Stages) * kI8TileBytes;  // 2 k-chunks: [1, 64 x8, 0..] | zeros
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + 2 * kI8ChunkBytes);
  uint64_t* b_full = bars;                               // [kI8Stages]
  uint64_t* b_empty = bars + kI8Stages;                  // [kI8Stages]
  uint64_t* b_tfull = bars + 2 * kI8Stages;              // [2]
  uint64_t* b_qempty = bars + 2 * kI8Stages + 6;         // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kI8Stages + 8);
  uint2* s_merge = reinterpret_cast<uint2*>(bars + 2 * kI8Stages + 9);  // [128] the two epilogue sets meet here

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_pairs * p.q_tiles;

  // constant A-side operand of the index K-step (the second chunk multiplies whatever follows a
  // tile's index chunk in shared memory by zero)
  for (int i = threadIdx.x; i < 2 * kI8Tile; i += kI8Threads)
    reinterpret_cast<uint4*>(s_ones)[i] =
        (i < kI8Tile) ? make_uint4(0x40404001u, 0x40404040u, 0x00000040u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads

  if (threadIdx.x == 0) {
    // full[s] of tile pair g completes when (a) its operand tile(s) have landed (producer
    // expect_tx + TMA complete_tx) AND (b) the four epilogue warps that drained tile pair g-2
    // have released its TMEM stage — ONE wait per tile pair on the MMA thread: every blocking
    // wait there idles the tensor pipe for ~170 clk (b2s_mma_microbench variants 2 vs 3).
    for (int s = 0; s < kI8Stages; ++s) {
      mbar_init(&b_full[s], 5);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&b_tfull[a], 1);
      mbar_init(&b_qempty[a], 1);
      for (int k = 0; k < 4; ++k) mbar_arrive(&b_full[a]);  // tile pairs 0 and 1 find their TMEM stage free
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
        const int nt = p.t_off[pair + 1] - p.t_off[pair];
        if (qt * kI8Tile >= nq || nt == 0) continue;
        const int n_tt = (nt + kI8Tile - 1) / kI8Tile;
        const uint32_t qb = n & 1u;
        const uint8_t* tsrc = p.tx + (size_t)pair * p.t_tiles * kI8TileBytes;
        for (int t = 0; t < n_tt; ++t, ++g) {
          const uint32_t s = g % kI8Stages;
          const long long c0 = p.dbg ? clock64() : 0;
          if (g >= kI8Stages) mbar_wait_bounded(&b_empty[s], ((g / kI8Stages) - 1u) & 1u);
          if (p.dbg) w_empty += clock64() - c0;
          if (t == 0) {  // the item's query tile rides on the barrier of its first train tile
            if (n >= 2) mbar_wait_bounded(&b_qempty[qb], ((n >> 1) - 1u) & 1u);
            mbar_arrive_expect_tx(&b_full[s], 2 * kI8TileBytes);
            bulk_g2s(s_q + (size_t)qb * kI8TileBytes, p.qx + ((size_t)pair * p.q_tiles + qt) * kI8TileBytes,
                     kI8TileBytes, &b_full[s]);
          } else if ((p.mode & 2) && g >= kI8Stages) {
            mbar_arrive(&b_full[s]);
            continue;
          } else {
            mbar_arrive_expect_tx(&b_full[s], kI8TileBytes);
          }
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
      bool probed = false;  // the barrier of the coming tile pair was already seen complete
      long long w_full = 0;
      const long long t_start = p.dbg ? clock64() : 0;
      const uint64_t odesc = make_smem_desc(smem_u32(s_ones));
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int pair = w / p.q_tiles, qt = w - pair * p.q_tiles;
        const int nq = p.q_off[pair + 1] - p.q_off[pair];
        const int nt = p.t_off[pair + 1] - p.t_off[pair];
        if (qt * kI8Tile >= nq || nt == 0) continue;
        const int n_tt = (nt + kI8Tile - 1) / kI8Tile;
        const uint32_t qb = n & 1u;
        const uint64_t qdesc = make_smem_desc(smem_u32(s_q + (size_t)qb * kI8TileBytes));
        for (int t = 0; t < n_tt; ++t, ++g) {
          const uint32_t s = g % kI8Stages, a = g & 1u;
          if (!probed) {  // operands landed AND TMEM stage a drained (first tile pair, item boundaries, late epilogue)
            const long long c0 = p.dbg ? clock64() : 0;
            mbar_wait_bounded(&b_full[s], (g / kI8Stages) & 1u);
            if (p.dbg) w_full += clock64() - c0;
          }
          tc_fence_after();
          const uint64_t tdesc = make_smem_desc(smem_u32(s_t + (size_t)s * kI8TileBytes));
          const uint32_t d1 = tmem_base + a * 256u, d2 = d1 + 128u;
          const uint32_t gn = g + 1u;
          probed = tc_mma_tile_pair(d1, d2, qdesc, tdesc, odesc, kIdescI8, &b_full[gn % kI8Stages],
                                    (gn / kI8Stages) & 1u, (t + 1 < n_tt) ? 1u : 0u) != 0u;
          tc_commit(&b_empty[s]);  // smem stage reusable once these MMAs have read it
          tc_commit(&b_tfull[a]);  // accumulators complete
        }
        tc_commit(&b_qempty[qb]);  // query tile buffer reusable once this item's MMAs have read it
        ++n;
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - t_start);
        p.dbg[blockIdx.x * 8 + 1] = 0ull;
        p.dbg[blockIdx.x * 8 + 2] = (unsigned long long)w_full;
        p.dbg[blockIdx.x * 8 + 6] = g;
      }
    }
  } else {
    // ===== epilogue: 2 sets of 4 warps; warp%4 = TMEM lane quarter, set e handles tile pairs with g%2 == e =====
    const int quarter = warp & 3;
    const uint32_t eset = (uint32_t)(warp - 2) >> 2;
    const int row = quarter * 32 + lane;  // TMEM lane = tile row
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + eset * 256u;
    uint32_t g = 0;
    long long w_tfull = 0;
    const long long e_start = p.dbg ? clock64() : 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int pair = w / p.q_tiles, qt = w - pair * p.q_tiles;
      const int qo = p.q_off[pair], nq = p.q_off[pair + 1] - qo;
      const int q0 = qt * kI8Tile;
      if (q0 >= nq) continue;
      const int to = p.t_off[pair], nt = p.t_off[pair + 1] - to;
      if (nt == 0) continue;
      const int n_tt = (nt + kI8Tile - 1) / kI8Tile;
      const int nq_valid = min(kI8Tile, nq - q0);
      uint32_t gbest = kNone, gsecond = kNone;
      for (int t = 0; t < n_tt; ++t, ++g) {
        if ((g & 1u) != eset) continue;
        const int tbase = t * kI8Tile;
        const int nt_valid = min(kI8Tile, nt - tbase);
        const long long c0 = p.dbg ? clock64() : 0;
        mbar_wait_bounded(&b_tfull[eset], (g >> 1) & 1u);
        if (p.dbg) w_tfull += clock64() - c0;
        tc_fence_after();
        if (p.mode & 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&b_full[(g + 2u) % kI8Stages]);
          continue;
        }
        uint32_t v0[32], v1[32], v2[32], v3[32];
        uint32_t b[2] = {kNone, kNone}, s2[2] = {kNone, kNone}, m[2] = {kNone, kNone};
        // D1 columns 0-63 | 64-127, D2 columns 0-63 | 64-127: 128 packed registers, all four
        // loads in flight at once, and the TMEM stage goes back to the MMA thread as soon as
        // they have landed — before any of the folding (the stage is held for the load latency
        // only, so the MMA thread's probe of the next barrier finds it complete)
        TMEM_LD_X32P(lane_addr, v0);
        TMEM_LD_X32P(lane_addr + 64u, v1);
        TMEM_LD_X32P(lane_addr + 128u, v2);
        TMEM_LD_X32P(lane_addr + 192u, v3);
        tmem_ld_wait();
        TMEM_REGS_READY(v0);
        TMEM_REGS_READY(v1);
        TMEM_REGS_READY(v2);
        TMEM_REGS_READY(v3);
        tc_fence_before();
        __syncwarp();
        // accumulator stage drained into registers: release it to tile pair g+2 (same stage)
        if (lane == 0) mbar_arrive(&b_full[(g + 2u) % kI8Stages]);
        fold_top2_p16(v0, b, s2);
        fold_top2_p16(v1, b, s2);
        fold_min_p16(v2, m);
        fold_min_p16(v3, m);
        // ---- per-row top-2 of this tile: merge the two chains, then the even / odd halves ----
        {
          const uint32_t bb = __vminu2(b[0], b[1]);
          const uint32_t ss = __vimin3_u16x2(__vmaxu2(b[0], b[1]), s2[0], s2[1]);
          const uint32_t bw = swap16(bb);
          const uint32_t best16 = __vminu2(bb, bw);                          // both halves: tile minimum
          const uint32_t sec16 = __vimin3_u16x2(__vmaxu2(bb, bw), ss, swap16(ss));  // both halves: tile second
          top2_insert(gbest, gsecond, key16_to_key32(best16 & 0xFFFFu, (uint32_t)tbase));
          top2_insert(gbest, gsecond, key16_to_key32(sec16 & 0xFFFFu, (uint32_t)tbase));
        }
        // ---- column minimum: this thread's train row against the 128 queries of the item ----
        {
          const uint32_t mm = __vminu2(m[0], m[1]);
          const uint32_t cm = __vminu2(mm, swap16(mm)) & 0xFFFFu;
          if (row < nt_valid && cm < kKey16Valid) atomicMin(&p.bwd_best[to + tbase + row], key16_to_key32(cm, (uint32_t)q0));
        }
      }
      // the two epilogue sets (even / odd tile pairs) meet in shared memory
      if (eset == 1) s_merge[row] = make_uint2(gbest, gsecond);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // epilogue warps only
      if (eset == 0 && row < nq_valid) {
        const uint2 o = s_merge[row];
        top2_insert(gbest, gsecond, o.x);
        top2_insert(gbest, gsecond, o.y);
        p.fwd_best[qo + q0 + row] = gbest >= kKey32Pad ? kNone : gbest;
        p.fwd_second[qo + q0 + row] = gsecond >= kKey32Pad ? kNone : gsecond;
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
// variant 0: 8 accumulating MMAs per iteration, nothing else.
// variant 1: the K2 issue pattern — 9 MMAs into D1 and 9 into D2 per iteration, no barriers.
// variant 2: variant 1 + the two tcgen05.commit of a tile pair (nobody waits on them).
// variant 3: variant 2 + waiting, before iteration i, for the commit of iteration i-2 (the
//            TMEM-stage dependency of the real kernel with an infinitely fast epilogue).
// variant 4: variant 3 with the wait for iteration i+1 taken after the first 4 MMAs of i.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int n_dim, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bars2[4];
  __shared__ uint32_t s_tm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (3 * kI8TileBytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bars2[i], 1);
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
    if (variant == 0) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tm + (uint32_t)((it & 1) * 256);
#pragma unroll
        for (int k = 0; k < 8; ++k) tc_mma_i8(d, adesc + (uint64_t)k * 256u, bdesc + (uint64_t)k * 256u, idesc, k > 0);
      }
    } else {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d1 = tm + (uint32_t)((it & 1) * 256), d2 = d1 + 128u;
        if (variant == 3 && it >= 2) mbar_wait_bounded(&bars2[it & 1], ((it >> 1) - 1) & 1);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          tc_mma_i8(d1, adesc + (uint64_t)(k & 7) * 256u, bdesc + (uint64_t)(k & 7) * 256u, idesc, k > 0);
          if (variant == 4 && k == 3 && it >= 1) mbar_wait_bounded(&bars2[(it + 1) & 1], (((it + 1) >> 1) - 1) & 1);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) tc_mma_i8(d2, bdesc + (uint64_t)(k & 7) * 256u, adesc + (uint64_t)(k & 7) * 256u, idesc, k > 0);
        if (variant >= 2) {
          tc_commit(&bars2[2 + (it & 1)]);
          tc_commit(&bars2[it & 1]);
        }
      }
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
  const int variant = n_dim >> 16;  // diagnostics: upper half selects the issue pattern (see mma_rate_kernel)
  n_dim &= 0xFFFF;
  B2S_REQUIRE(n_dim == 128 || n_dim == 256, "n_dim must be 128 or 256");
  B2S_REQUIRE(variant >= 0 && variant <= 4 && (variant == 0 || n_dim == 128), "bad variant");
  const size_t smem = 3 * (size_t)kI8TileBytes;
  B2S_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate_kernel<<<sm_count(), 128, smem, st>>>(iters, n_dim, variant);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  if (macs_out) *macs_out = (double)sm_count() * iters * (variant ? 18.0 : 8.0) * 128.0 * n_dim * 32.0;
  return B2S_OK;
}

static unsigned long long* g_i8_dbg = nullptr;  // device buffer, 8 u64 per SM (diagnostics only)
static int g_i8_mode = 0;
void hamming_i8_set_debug(unsigned long long* dev_buf, int mode) {
  g_i8_dbg = dev_buf;
  g_i8_mode = mode;
}

constexpr size_t kI8SmemBytes =
    (size_t)(2 + kI8Stages) * kI8TileBytes + 2 * kI8ChunkBytes + 8 * (2 * kI8Stages + 9) + 128 * sizeof(uint2);

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
  expand_pm8_kernel<<<dim3((qt * kI8Units + 255) / 256, n_pairs), 256, 0, st>>>(q, q_off, q_src, qt, 0,
                                                                                reinterpret_cast<uint4*>(qx));
  B2S_CUDA(cudaGetLastError());
  expand_pm8_kernel<<<dim3((tt * kI8Units + 255) / 256, n_pairs), 256, 0, st>>>(t, t_off, t_src, tt, 1,
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
  p.mode = g_i8_mode;
  const long items = (long)qt * n_pairs;
  const int grid = (int)(items < (long)sm_count() ? items : (long)sm_count());
  hamming_knn2_i8_kernel<<<grid, kI8Threads, kI8SmemBytes, st>>>(p);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

}  // namespace b2s
