// ransac_tc.cu — K3t: Sampson scoring of every (hypothesis, correspondence) on the tensor cores.
//
// Same result as K3 / K3h (ransac.cu): per pair and hypothesis the number of correspondences with
// num^2 < th^2 den (homography.py:328-333), decided in float64 arithmetic.  Here the two
// bilinear forms come out of tcgen05.mma kind::tf32:
//     num[h][m] = sum_k n_h[k] phi_m[k]     phi = (ux, uy, u, vx, vy, v, x, y, 1),  n = E row-major
//     den[h][m] = sum_k g_h[k] psi_m[k]     psi = (x^2, xy, y^2, x, y, u^2, uv, v^2, u, v, 1)
//                                            g   = the two 3x3 Gram matrices of E's first two rows /
//                                                  columns (off-diagonals doubled, constants merged)
// Every coefficient is split into a tf32 "hi" and a tf32 "lo" part (22 mantissa bits together)
// and each form is accumulated from three products (lo.hi + hi.lo + hi.hi, K padded to 16), so
// the accumulators carry float32-class accuracy.  The epilogue (thread = TMEM lane = hypothesis)
// evaluates d = num^2 - th^2 den with a bound B on everything the split, the tensor-core
// accumulation and the float32 epilogue can have contributed; |d| > B decides, and only inside
// the band the same thread re-evaluates that one pair with K3's float64 expression.  The counts
// therefore equal K3's (tests/test_gpu_ransac.py), at ~9 float32 instructions per evaluation
// instead of 20 float64 ones.
// Structure = K2 (hamming_i8.cu): persistent CTAs, item = (pair, 128-hypothesis tile), the
// pair's correspondence tiles stream through a bulk-copy ring, 2 TMEM stages x (num 128 + den
// 128 columns), one barrier wait per tile pair on the MMA thread with the probe inside the
// instruction block, two epilogue sets of 4 warps alternating tile pairs.
#include "sampson.cuh"
#include "tcgen05.cuh"

namespace b2s {

constexpr int kTcChunks = 16;                        // per row: n_hi | n_lo | g_hi | g_lo, 16 floats each
constexpr int kTcTileBytes = kTcChunks * kTcChunkBytes;  // 32 KB
constexpr int kTcUnits = kTcChunks * kTcTileRows;
constexpr int kTcStages = 4;
constexpr int kTcThreads = 576;                      // producer warp + MMA warp + 16 epilogue warps
constexpr int kTcWlBlock = 512;                      // work-list entries a warp reserves per atomicAdd
// instruction descriptor: F32 accumulate, TF32 A and B, both K-major, M = 128, N = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
// error of one accumulated form relative to ||coefficients|| * ||monomials||: 3 * 2^-22 from the
// split (dropped lo.lo, lo rounded to tf32) + 6 accumulator updates and 8-term inner sums at
// <= 2^-23 each ~ 2.4e-6 worst case; the measured error stays below 1e-6 (tests/test_gpu_ransac.py
// asserts it is under half of kappa on every case it runs).
constexpr float kTcKappa = 3.8146973e-6f;            // 2^-18

__device__ __forceinline__ void split_tf32(double v, float& hi, float& lo) {
  const float f = (float)v;
  hi = __uint_as_float((__float_as_uint(f) + 0x1000u) & 0xFFFFE000u);   // round to 10 mantissa bits
  const float l = (float)(v - (double)hi);
  lo = __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xFFFFE000u);
}

// ---- pre-pass A: hypotheses -> operand tiles [pair][h_tile][16 chunks][128 rows] + norms ----
__global__ void __launch_bounds__(128) tc_prep_hyp_kernel(const double* __restrict__ E, int H, int h_tiles,
                                                          uint4* __restrict__ out, float2* __restrict__ norms) {
  const int pair = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
  const int h = tile * kTcTileRows + r;
  float hi[32], lo[32];  // [0,16): n, [16,32): g
#pragma unroll
  for (int k = 0; k < 32; ++k) hi[k] = lo[k] = 0.0f;
  float nE = 0.0f, nG = 0.0f;
  if (h < H) {
    double e[9];
    const double* ep = E + ((size_t)pair * H + h) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = ep[k];
    double G[3][3], Gp[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        G[a][b] = fma(e[a], e[b], e[3 + a] * e[3 + b]);              // rows 0, 1 of E:    |(E x1)_{0,1}|^2
        Gp[a][b] = fma(e[3 * a], e[3 * b], e[3 * a + 1] * e[3 * b + 1]);  // columns 0, 1: |(E^T x2)_{0,1}|^2
      }
    const double g[11] = {G[0][0], 2.0 * G[0][1], G[1][1], 2.0 * G[0][2], 2.0 * G[1][2],
                          Gp[0][0], 2.0 * Gp[0][1], Gp[1][1], 2.0 * Gp[0][2], 2.0 * Gp[1][2], G[2][2] + Gp[2][2]};
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      split_tf32(e[k], hi[k], lo[k]);
      s1 = fma(e[k], e[k], s1);
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      split_tf32(g[k], hi[16 + k], lo[16 + k]);
      s2 = fma(g[k], g[k], s2);
    }
    nE = (float)sqrt(s1) * 1.0001f;
    nG = (float)sqrt(s2) * 1.0001f;
    norms[(size_t)pair * H + h] = make_float2(nE, nG);
  }
  uint4* o = out + ((size_t)pair * h_tiles + tile) * kTcUnits + r;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    o[(0 + c) * kTcTileRows] = make_uint4(__float_as_uint(hi[4 * c]), __float_as_uint(hi[4 * c + 1]), __float_as_uint(hi[4 * c + 2]), __float_as_uint(hi[4 * c + 3]));
    o[(4 + c) * kTcTileRows] = make_uint4(__float_as_uint(lo[4 * c]), __float_as_uint(lo[4 * c + 1]), __float_as_uint(lo[4 * c + 2]), __float_as_uint(lo[4 * c + 3]));
    o[(8 + c) * kTcTileRows] = make_uint4(__float_as_uint(hi[16 + 4 * c]), __float_as_uint(hi[16 + 4 * c + 1]), __float_as_uint(hi[16 + 4 * c + 2]), __float_as_uint(hi[16 + 4 * c + 3]));
    o[(12 + c) * kTcTileRows] = make_uint4(__float_as_uint(lo[16 + 4 * c]), __float_as_uint(lo[16 + 4 * c + 1]), __float_as_uint(lo[16 + 4 * c + 2]), __float_as_uint(lo[16 + 4 * c + 3]));
  }
}

// ---- pre-pass B: correspondences -> operand tiles [pair][m_tile][16 chunks][128 rows] + per-pair max norms ----
__global__ void __launch_bounds__(128) tc_prep_corr_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ c_off,
                                                           const int32_t* __restrict__ c_count, int m_tiles,
                                                           uint4* __restrict__ out, int* __restrict__ pairmax) {
  const int pair = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
  const int m = tile * kTcTileRows + r;
  const int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  if (tile * kTcTileRows >= M) return;  // tile never read
  float hi[32], lo[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) hi[k] = lo[k] = 0.0f;
  float p2 = 0.0f, q2 = 0.0f;
  if (m < M) {
    const float4 c = corr[c_off[pair] + m];
    const double x = c.x, y = c.y, u = c.z, v = c.w;
    const double phi[9] = {u * x, u * y, u, v * x, v * y, v, x, y, 1.0};
    const double psi[11] = {x * x, x * y, y * y, x, y, u * u, u * v, v * v, u, v, 1.0};
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      split_tf32(phi[k], hi[k], lo[k]);
      s1 = fma(phi[k], phi[k], s1);
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      split_tf32(psi[k], hi[16 + k], lo[16 + k]);
      s2 = fma(psi[k], psi[k], s2);
    }
    p2 = (float)s1 * 1.0001f;
    q2 = (float)s2 * 1.0001f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    p2 = fmaxf(p2, __shfl_xor_sync(0xFFFFFFFFu, p2, o));
    q2 = fmaxf(q2, __shfl_xor_sync(0xFFFFFFFFu, q2, o));
  }
  if ((r & 31) == 0) {  // non-negative floats order like their bit patterns
    atomicMax(&pairmax[2 * pair], __float_as_int(p2));
    atomicMax(&pairmax[2 * pair + 1], __float_as_int(q2));
  }
  // B side: the hi block meets A's lo block and vice versa, so the row keeps hi | lo in the same slots
  uint4* o = out + ((size_t)pair * m_tiles + tile) * kTcUnits + r;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    o[(0 + c) * kTcTileRows] = make_uint4(__float_as_uint(hi[4 * c]), __float_as_uint(hi[4 * c + 1]), __float_as_uint(hi[4 * c + 2]), __float_as_uint(hi[4 * c + 3]));
    o[(4 + c) * kTcTileRows] = make_uint4(__float_as_uint(lo[4 * c]), __float_as_uint(lo[4 * c + 1]), __float_as_uint(lo[4 * c + 2]), __float_as_uint(lo[4 * c + 3]));
    o[(8 + c) * kTcTileRows] = make_uint4(__float_as_uint(hi[16 + 4 * c]), __float_as_uint(hi[16 + 4 * c + 1]), __float_as_uint(hi[16 + 4 * c + 2]), __float_as_uint(hi[16 + 4 * c + 3]));
    o[(12 + c) * kTcTileRows] = make_uint4(__float_as_uint(lo[16 + 4 * c]), __float_as_uint(lo[16 + 4 * c + 1]), __float_as_uint(lo[16 + 4 * c + 2]), __float_as_uint(lo[16 + 4 * c + 3]));
  }
}

// The 12 MMAs of one (hypothesis tile, correspondence tile) pair as one asm block: per form
// lo.hi, hi.lo, hi.hi (small terms first), two K = 8 steps each; the barrier of the next tile
// pair is probed after the 8th MMA and its predicate read after the last (see hamming_i8.cu).
__device__ __forceinline__ uint32_t tc_mma_score_pair(uint32_t d1, uint32_t d2, uint64_t adesc, uint64_t bdesc,
                                                      uint32_t idesc, uint64_t* next_bar, uint32_t next_parity,
                                                      uint32_t has_next) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, pnew, pprobe, pdone;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.eq.b32 pacc, 0, 0;\n\t"
      "setp.ne.b32 pnew, 0, 0;\n\t"
      "setp.ne.b32 pprobe, %8, 0;\n\t"
      "setp.ne.b32 pdone, 0, 0;\n\t"
      "add.u64 da, %3, 512;\n\t"
      "add.u64 db, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], da, db, %5, pnew;\n\t"
      "add.u64 da, %3, 768;\n\t"
      "add.u64 db, %4, 256;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 0;\n\t"
      "add.u64 db, %4, 512;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 256;\n\t"
      "add.u64 db, %4, 768;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 0;\n\t"
      "add.u64 db, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 256;\n\t"
      "add.u64 db, %4, 256;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 1536;\n\t"
      "add.u64 db, %4, 1024;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%2], da, db, %5, pnew;\n\t"
      "add.u64 da, %3, 1792;\n\t"
      "add.u64 db, %4, 1280;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%2], da, db, %5, pacc;\n\t"
      "@pprobe mbarrier.try_wait.parity.shared::cta.b64 pdone, [%6], %7;\n\t"
      "add.u64 da, %3, 1024;\n\t"
      "add.u64 db, %4, 1536;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%2], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 1280;\n\t"
      "add.u64 db, %4, 1792;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%2], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 1024;\n\t"
      "add.u64 db, %4, 1024;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%2], da, db, %5, pacc;\n\t"
      "add.u64 da, %3, 1280;\n\t"
      "add.u64 db, %4, 1280;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%2], da, db, %5, pacc;\n\t"
      "selp.u32 %0, 1, 0, pdone;\n\t"
      "}"
      : "=r"(ok)
      : "r"(d1), "r"(d2), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(smem_u32(next_bar)), "r"(next_parity), "r"(has_next)
      : "memory");
  return ok;
}

struct TcParams {
  const uint8_t* __restrict__ ax;   // hypothesis tiles      [pair][h_tiles][32 KB]
  const uint8_t* __restrict__ bx;   // correspondence tiles  [pair][m_tiles][32 KB]
  const float2* __restrict__ norms; // [pair][H]: ||n||, ||g||
  const int* __restrict__ pairmax;  // [pair][2]: max ||phi||^2, max ||psi||^2 (float bits)
  const float4* __restrict__ corr;
  const int32_t* __restrict__ c_off;
  const int32_t* __restrict__ c_count;
  const double* __restrict__ E;
  const double* __restrict__ th2_pp;
  double th2_all;
  int32_t* __restrict__ counts;
  float* __restrict__ dbg_num;      // optional [pair][H][ld]: raw accumulators (tests / bring-up)
  float* __restrict__ dbg_den;
  uint2* __restrict__ wl;            // work list of undecidable evaluations: (pair * H + h, m)
  int* __restrict__ wl_count;       // [0] entries appended (may exceed wl_cap: then [1] is set and K3 reruns everything)
  int wl_cap;
  int H, h_tiles, m_tiles, n_pairs, dbg_ld;
};

__global__ void __launch_bounds__(kTcThreads, 1) ransac_score_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_a = smem;                                   // 2 x 32 KB (double-buffered hypothesis tile)
  uint8_t* s_b = smem + 2 * kTcTileBytes;                // kTcStages x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + kTcStages) * kTcTileBytes);
  uint64_t* b_full = bars;                               // [kTcStages]
  uint64_t* b_empty = bars + kTcStages;                  // [kTcStages]
  uint64_t* b_tfull = bars + 2 * kTcStages;              // [2]
  uint64_t* b_aempty = bars + 2 * kTcStages + 2;         // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_pairs * p.h_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&b_full[s], 9);   // producer + the 8 warps that drained tile pair g-2 (see hamming_i8.cu)
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&b_tfull[a], 1);
      mbar_init(&b_aempty[a], 1);
      for (int k = 0; k < 8; ++k) mbar_arrive(&b_full[a]);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // items without correspondences are skipped by every role (their counts stay 0)
  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      uint32_t n = 0, g = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int pair = w / p.h_tiles, ht = w - pair * p.h_tiles;
        const int M = max(p.c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
        if (M <= 0 || ht * kTcTileRows >= p.H) continue;
        const int n_mt = (M + kTcTileRows - 1) / kTcTileRows;
        const uint32_t ab = n & 1u;
        const uint8_t* bsrc = p.bx + (size_t)pair * p.m_tiles * kTcTileBytes;
        for (int t = 0; t < n_mt; ++t, ++g) {
          const uint32_t s = g % kTcStages;
          if (g >= kTcStages) mbar_wait_bounded(&b_empty[s], ((g / kTcStages) - 1u) & 1u);
          if (t == 0) {  // the item's hypothesis tile rides on the barrier of its first correspondence tile
            if (n >= 2) mbar_wait_bounded(&b_aempty[ab], ((n >> 1) - 1u) & 1u);
            mbar_arrive_expect_tx(&b_full[s], 2 * kTcTileBytes);
            bulk_g2s(s_a + (size_t)ab * kTcTileBytes, p.ax + ((size_t)pair * p.h_tiles + ht) * kTcTileBytes, kTcTileBytes,
                     &b_full[s]);
          } else {
            mbar_arrive_expect_tx(&b_full[s], kTcTileBytes);
          }
          bulk_g2s(s_b + (size_t)s * kTcTileBytes, bsrc + (size_t)t * kTcTileBytes, kTcTileBytes, &b_full[s]);
        }
        ++n;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      uint32_t n = 0, g = 0;
      bool probed = false;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int pair = w / p.h_tiles, ht = w - pair * p.h_tiles;
        const int M = max(p.c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
        if (M <= 0 || ht * kTcTileRows >= p.H) continue;
        const int n_mt = (M + kTcTileRows - 1) / kTcTileRows;
        const uint32_t ab = n & 1u;
        const uint64_t adesc = make_smem_desc(smem_u32(s_a + (size_t)ab * kTcTileBytes));
        for (int t = 0; t < n_mt; ++t, ++g) {
          const uint32_t s = g % kTcStages, a = g & 1u;
          if (!probed) mbar_wait_bounded(&b_full[s], (g / kTcStages) & 1u);
          tc_fence_after();
          const uint64_t bdesc = make_smem_desc(smem_u32(s_b + (size_t)s * kTcTileBytes));
          const uint32_t d1 = tmem_base + a * 256u, d2 = d1 + 128u;
          const uint32_t gn = g + 1u;
          probed = tc_mma_score_pair(d1, d2, adesc, bdesc, kIdescTf32, &b_full[gn % kTcStages], (gn / kTcStages) & 1u,
                                     (t + 1 < n_mt) ? 1u : 0u) != 0u;
          tc_commit(&b_empty[s]);
          tc_commit(&b_tfull[a]);
        }
        tc_commit(&b_aempty[ab]);
        ++n;
      }
    }
  } else {
    // ===== epilogue: 4 sets of 4 warps; warp%4 = TMEM lane quarter; set = (tile-pair parity, column half):
    // warps of set (e, half) handle columns [64*half, 64*half + 64) of the tile pairs with g%2 == e =====
    const int quarter = warp & 3;
    const uint32_t set = (uint32_t)(warp - 2) >> 2;
    const uint32_t eset = set & 1u, half = set >> 1;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + eset * 256u + half * 64u;
    uint32_t g = 0;
    int wl_at = 0, wl_end = 0;  // this warp's reserved slice of the work list (uniform across the warp)
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int pair = w / p.h_tiles, ht = w - pair * p.h_tiles;
      const int M = max(p.c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
      if (M <= 0 || ht * kTcTileRows >= p.H) continue;
      const int n_mt = (M + kTcTileRows - 1) / kTcTileRows;
      const int h = ht * kTcTileRows + row;
      const bool live = h < p.H;
      const double th2d = p.th2_pp ? p.th2_pp[pair] : p.th2_all;
      const float th2 = (float)th2d;
      // bound constants of this hypothesis (see the file header): delta on num, eps on den
      float two_delta = 0.0f, c0 = 0.0f;
      if (live) {
        const float2 nn = p.norms[(size_t)pair * p.H + h];
        const float Phi = sqrtf(__int_as_float(p.pairmax[2 * pair])) * 1.0001f;
        const float Psi = sqrtf(__int_as_float(p.pairmax[2 * pair + 1])) * 1.0001f;
        const float delta = kTcKappa * nn.x * Phi, eps = kTcKappa * nn.y * Psi;
        two_delta = 2.0f * delta;
        c0 = fmaf(delta, delta, th2 * eps) * 1.0001f;
      }
      constexpr float kRho = 16.0f * 5.9604645e-8f;  // float32 roundings of the epilogue itself, relative to num^2 + th^2 den
      const uint32_t ph = (uint32_t)pair * (uint32_t)p.H + (uint32_t)h;
      float count_f = 0.0f;  // exact: at most 2^24 increments of 1.0
      for (int t = 0; t < n_mt; ++t, ++g) {
        if ((g & 1u) != eset) continue;
        const int mbase = t * kTcTileRows + 64 * (int)half;
        mbar_wait_bounded(&b_tfull[eset], (g >> 1) & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t xn[32], xd[32];
          TMEM_LD_X32(lane_addr + (uint32_t)(32 * c), xn);
          TMEM_LD_X32(lane_addr + 128u + (uint32_t)(32 * c), xd);
          tmem_ld_wait();
          TMEM_REGS_READY(xn);
          TMEM_REGS_READY(xd);
          if (c == 1) {  // this warp's 64 columns are in registers: its share of the stage goes to tile pair g+2
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[(g + 2u) % kTcStages]);
          }
          const int m0 = mbase + 32 * c;
          const int nv = live ? min(32, M - m0) : 0;
          if (p.dbg_num && nv > 0) {
            float* dn = p.dbg_num + ((size_t)pair * p.H + h) * p.dbg_ld + m0;
            float* dd = p.dbg_den + ((size_t)pair * p.H + h) * p.dbg_ld + m0;
#pragma unroll
            for (int k = 0; k < 32; ++k)  // static register indices (a dynamic one would push the arrays to local memory)
              if (k < nv) {
                dn[k] = __uint_as_float(xn[k]);
                dd[k] = __uint_as_float(xd[k]);
              }
          }
          unsigned amb = 0u;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float num = __uint_as_float(xn[k]), den = __uint_as_float(xd[k]);
            const float T = th2 * den;
            const float d = fmaf(num, num, -T);
            const float sum = fmaf(num, num, fabsf(T));
            const float B = fmaf(two_delta, fabsf(num), fmaf(kRho, sum, c0));
            count_f += (d < -B) ? 1.0f : 0.0f;
            if (!(fabsf(d) > B)) amb |= 1u << k;
          }
          if (nv < 32) {  // tail chunk / dead row: undo what the columns past the end contributed
            const unsigned keep = nv > 0 ? (0xFFFFFFFFu >> (32 - nv)) : 0u;
            unsigned in_mask = 0u;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float num = __uint_as_float(xn[k]), den = __uint_as_float(xd[k]);
              const float T = th2 * den;
              const float d = fmaf(num, num, -T);
              const float sum = fmaf(num, num, fabsf(T));
              const float B = fmaf(two_delta, fabsf(num), fmaf(kRho, sum, c0));
              if (d < -B) in_mask |= 1u << k;
            }
            count_f -= (float)__popc(in_mask & ~keep);
            amb &= keep;
          }
          // undecidable in float32: append to the work list; the float64 pass after this kernel
          // decides those.  Each warp owns a reserved slice and only touches the global counter
          // when the slice runs out, so there is no atomic round trip inside the tile loop.
          const int mine = __popc(amb);
          if (__any_sync(0xFFFFFFFFu, mine != 0)) {
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
              if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (wl_at + total > wl_end) {  // new slice; the unused tail of the old one is marked "no entry"
              if (wl_end <= p.wl_cap)
                for (int i = wl_at + lane; i < wl_end; i += 32) p.wl[i] = make_uint2(0u, 0u);
              int base = 0;
              if (lane == 0) base = atomicAdd(&p.wl_count[0], kTcWlBlock);
              wl_at = __shfl_sync(0xFFFFFFFFu, base, 0);
              wl_end = wl_at + kTcWlBlock;
            }
            if (wl_end > p.wl_cap) {
              if (lane == 0) p.wl_count[1] = 1;           // overflow: everything is redone in float64 afterwards
            } else {
              int at = wl_at + incl - mine;
              while (amb) {
                const int k = __ffs(amb) - 1;
                amb &= amb - 1u;
                p.wl[at++] = make_uint2(ph + 1u, (uint32_t)(m0 + k));   // + 1: zero marks an unused slot
              }
            }
            wl_at += total;
          }
        }
      }
      const int count = (int)count_f;
      if (live && count) atomicAdd(&p.counts[(size_t)pair * p.H + h], count);
    }
    if (wl_end <= p.wl_cap)
      for (int i = wl_at + lane; i < wl_end; i += 32) p.wl[i] = make_uint2(0u, 0u);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- float64 pass over the work list ----
__global__ void __launch_bounds__(256) tc_fixup_kernel(const uint2* __restrict__ wl, const int* __restrict__ wl_count, int wl_cap,
                                                       const float4* __restrict__ corr, const int32_t* __restrict__ c_off,
                                                       const double* __restrict__ E, int H, double th2_all,
                                                       const double* __restrict__ th2_pp, int32_t* __restrict__ counts) {
  if (wl_count[1]) return;  // overflowed: the all-float64 kernel recomputes every count
  const int n = min(wl_count[0], wl_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint2 w = wl[i];
    if (w.x == 0u) continue;  // unused slot of a warp's reserved slice
    w.x -= 1u;
    const int pair = (int)(w.x / (uint32_t)H);
    double e[9];
    const double* ep = E + (size_t)w.x * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = ep[k];
    const float4 c = corr[c_off[pair] + (int)w.y];
    const double th2 = th2_pp ? th2_pp[pair] : th2_all;
    if (sampson_inlier<double>(e, (double)c.x, (double)c.y, (double)c.z, (double)c.w, th2)) atomicAdd(&counts[w.x], 1);
  }
}

// all-float64 K3 (ransac.cu), run only if *run_flag != 0
int ransac_score_fp64_cond_launch(const float4* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs, const double* E,
                                  int H, double th2, const double* th2_pp, int32_t* counts, const int* run_flag, cudaStream_t st);

constexpr size_t kTcSmemBytes = (size_t)(2 + kTcStages) * kTcTileBytes + 8 * (2 * kTcStages + 4) + 16;

static size_t tc_align(size_t x) { return (x + 255) & ~(size_t)255; }
// room for 1 in 16 evaluations to be undecidable in float32 (tracking data: ~1 in 100), plus one
// reserved slice per epilogue warp of every CTA
static size_t tc_worklist_cap(int n_pairs, int H, int max_m) {
  const size_t evals = (size_t)n_pairs * H * max_m;
  const size_t cap = evals / 16 + (size_t)2 * 148 * 16 * kTcWlBlock;
  return cap > 0x7FFFFFFFull ? 0x7FFFFFFFull : cap;
}

}  // namespace b2s

extern "C" {

size_t b2s_ransac_score_tc_workspace_bytes(int n_pairs, int H, int max_m) {
  using namespace b2s;
  const size_t ht = (size_t)((H + kTcTileRows - 1) / kTcTileRows), mt = (size_t)((max_m + kTcTileRows - 1) / kTcTileRows);
  return tc_align((size_t)n_pairs * ht * kTcTileBytes) + tc_align((size_t)n_pairs * mt * kTcTileBytes) +
         tc_align((size_t)n_pairs * H * sizeof(float2)) + tc_align((size_t)n_pairs * 2 * sizeof(int)) + tc_align(256) +
         tc_align(tc_worklist_cap(n_pairs, H, max_m) * sizeof(uint2));
}

int b2s_ransac_score_tc(const float* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs, int max_m,
                        const double* E, int H, double th2, const double* th2_per_pair, int32_t* counts,
                        void* workspace, size_t workspace_bytes, float* dbg_num, float* dbg_den, int dbg_ld,
                        int32_t* dbg_band, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(corr && c_off && c_count && E && counts, "null pointer");
  B2S_REQUIRE(n_pairs >= 0 && H >= 0 && max_m >= 0, "negative size");
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  if (n_pairs == 0 || H == 0) return B2S_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B2S_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_pairs * H, st));
  if (max_m == 0) return B2S_OK;
  const size_t need = b2s_ransac_score_tc_workspace_bytes(n_pairs, H, max_m);
  B2S_REQUIRE(workspace != nullptr && workspace_bytes >= need, "tensor-core scoring needs %zu workspace bytes, got %zu", need,
              workspace_bytes);
  B2S_REQUIRE(((uintptr_t)workspace & 127u) == 0, "workspace must be 128-byte aligned");
  B2S_REQUIRE((dbg_num == nullptr) == (dbg_den == nullptr), "pass both debug buffers or neither");
  const int ht = (H + kTcTileRows - 1) / kTcTileRows, mt = (max_m + kTcTileRows - 1) / kTcTileRows;
  uint8_t* ax = static_cast<uint8_t*>(workspace);
  uint8_t* bx = ax + tc_align((size_t)n_pairs * ht * kTcTileBytes);
  float2* norms = reinterpret_cast<float2*>(bx + tc_align((size_t)n_pairs * mt * kTcTileBytes));
  int* pairmax = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(norms) + tc_align((size_t)n_pairs * H * sizeof(float2)));
  int* wl_count = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(pairmax) + tc_align((size_t)n_pairs * 2 * sizeof(int)));
  uint2* wl = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(wl_count) + tc_align(256));
  const size_t wl_cap = tc_worklist_cap(n_pairs, H, max_m);
  // pairmax and the two work-list counters are adjacent: one memset
  B2S_CUDA(cudaMemsetAsync(pairmax, 0, tc_align((size_t)n_pairs * 2 * sizeof(int)) + 256, st));
  tc_prep_hyp_kernel<<<dim3(ht, n_pairs), 128, 0, st>>>(E, H, ht, reinterpret_cast<uint4*>(ax), norms);
  B2S_CUDA(cudaGetLastError());
  tc_prep_corr_kernel<<<dim3(mt, n_pairs), 128, 0, st>>>(reinterpret_cast<const float4*>(corr), c_off, c_count, mt,
                                                         reinterpret_cast<uint4*>(bx), pairmax);
  B2S_CUDA(cudaGetLastError());
  note_launch(2);
  static bool attr_set[64] = {false};
  if (first_use_on_device(attr_set))
    B2S_CUDA(cudaFuncSetAttribute(ransac_score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  TcParams p;
  p.ax = ax;
  p.bx = bx;
  p.norms = norms;
  p.pairmax = pairmax;
  p.corr = reinterpret_cast<const float4*>(corr);
  p.c_off = c_off;
  p.c_count = c_count;
  p.E = E;
  p.th2_pp = th2_per_pair;
  p.th2_all = th2;
  p.counts = counts;
  p.dbg_num = dbg_num;
  p.dbg_den = dbg_den;
  p.wl = wl;
  p.wl_count = wl_count;
  p.wl_cap = (int)wl_cap;
  p.H = H;
  p.h_tiles = ht;
  p.m_tiles = mt;
  p.n_pairs = n_pairs;
  p.dbg_ld = dbg_ld;
  const long items = (long)ht * n_pairs;
  const int grid = (int)(items < (long)sm_count() ? items : (long)sm_count());
  ransac_score_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(p);
  B2S_CUDA(cudaGetLastError());
  tc_fixup_kernel<<<2 * sm_count(), 256, 0, st>>>(wl, wl_count, (int)wl_cap, p.corr, c_off, E, H, th2, th2_per_pair, counts);
  B2S_CUDA(cudaGetLastError());
  note_launch(2);
  if (dbg_band) B2S_CUDA(cudaMemcpyAsync(dbg_band, wl_count, 2 * sizeof(int), cudaMemcpyDeviceToDevice, st));
  // overflow rescue (exits immediately unless the work list overflowed)
  return ransac_score_fp64_cond_launch(p.corr, c_off, c_count, n_pairs, E, H, th2, th2_per_pair, counts, wl_count + 1, st);
}

}  // extern "C"
