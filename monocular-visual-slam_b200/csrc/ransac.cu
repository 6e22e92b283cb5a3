// ransac.cu — K3/K4: batched RANSAC essential-matrix hypotheses.
//
// Replaces the body of the Python loop of ransac_essential
// (/root/reference/homography.py:324-339): the 8-point minimal solve
// (eight_point_E, homography.py:222-248, including the K^T F K return quirk), the
// Sampson scoring of every correspondence (homography.py:328-333) and the sequential
// best / early-exit bookkeeping (homography.py:335-339), for thousands of hypotheses
// and many frame pairs per launch.
//
// Arithmetic is float64 like the reference (B200 FP64 runs at half the FP32 rate, so
// this costs ~1 us per 2000x500 pair); a float32 scoring variant exists for comparison.
// The inlier test is the division-free form  num^2 < th^2 * den  of the reference's
// num^2 / den < th^2  (den = 0 -> never an inlier in both).
#include <type_traits>

#include "common.cuh"
#include "linalg.cuh"
#include "sampson.cuh"

namespace b2s {

struct Mat3 {
  double m[9];
};

// ---- K3: one thread = one hypothesis, correspondences broadcast from shared memory ----
constexpr int kScoreThreads = 128;
constexpr int kScoreChunk = 1024;

template <typename T>
__global__ void __launch_bounds__(kScoreThreads) ransac_score_kernel(
    const float4* __restrict__ corr, const int32_t* __restrict__ c_off, const int32_t* __restrict__ c_count,
    const double* __restrict__ E, int H, double th2_all, const double* __restrict__ th2_pp,
    int32_t* __restrict__ counts, const int* __restrict__ run_flag) {
  if (run_flag && *run_flag == 0) return;  // conditional rerun after the tensor-core pass (ransac_tc.cu)
  struct alignas(16) P4 { T x, y, u, v; };
  __shared__ P4 s_p[kScoreChunk];
  const int pair = blockIdx.y;
  const int h = blockIdx.x * kScoreThreads + threadIdx.x;
  const int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  const float4* cp = corr + c_off[pair];
  const T th2 = (T)(th2_pp ? th2_pp[pair] : th2_all);
  T e[9];
  const bool live = h < H;
  {
    const double* ep = E + ((size_t)pair * H + (live ? h : 0)) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = (T)ep[k];
  }
  int count = 0;
  for (int base = 0; base < M; base += kScoreChunk) {
    const int n = min(kScoreChunk, M - base);
    __syncthreads();
    for (int m = threadIdx.x; m < n; m += kScoreThreads) {
      const float4 c = __ldg(cp + base + m);
      s_p[m] = P4{(T)c.x, (T)c.y, (T)c.z, (T)c.w};
    }
    __syncthreads();
#pragma unroll 4
    for (int m = 0; m < n; ++m) {
      const P4 c = s_p[m];
      count += sampson_inlier<T>(e, c.x, c.y, c.u, c.v, th2) ? 1 : 0;
    }
  }
  if (live) counts[(size_t)pair * H + h] = count;
}

// ---- K3h: float64-exact decisions at float32 cost -----------------------------------------
// Every (hypothesis, correspondence) is first evaluated in float32 together with a rigorous
// bound B on what float32 rounding (of E and of the 19 operations) can have done to
// d = num^2 - th^2 den.  |d| > B decides; only the band |d| <= B (a few per 10^5 on tracking
// data) is re-evaluated in float64 by the same thread with the very expression of K3, so the
// counts equal the float64 kernel's counts while the FP64 pipe (half the FP32 rate on B200, and
// 2 issue cycles per instruction) is out of the inner loop.
//   u = 2^-24; e = fl32(E); nE = ||e||_F, W1 = max ||(x1, 1)||, W2 = max ||(x2, 1)|| over the pair; the
//   points are float32 already, so only E is rounded on input.  With r_i / c_j the row / column norms of E:
//   a_i = fma(e_i0, x, fma(e_i1, y, e_i2)):  |a_i - fl(a_i)| <= u (2|E_i0 x| + 3|E_i1 y| + 3|E_i2|) <= 3u r_i W1
//   (one input rounding per term + two FMA roundings, Cauchy-Schwarz), likewise |b_j - fl(b_j)| <= 3u c_j W2;
//   num = fma(u, a0, fma(v, a1, a2)):  the a-errors give <= 3u W1 W2 nE, the two FMA roundings <= 2u W2 nE W1:
//       |num - fl(num)| <= 5u nE W1 W2 =: delta;
//   den (four non-negative terms, nested FMAs):  |den - fl(den)| <= 6u nE^2 (W1^2 + W2^2) + 4u den;
//   T = fl(fl32(th^2) den), d = fma(num, num, -T):
//       |d - fl(d)| <= 2 |num| delta + delta^2 + 6u th^2 nE^2 (W1^2 + W2^2) + 7u (num^2 + T)
//   (all to first order in u; the float64 kernel's own rounding is 9 orders of magnitude below).
//   B below takes 7u, 9u and 10u for the 5u, 6u and 7u (1.4x), and every norm is nudged up by 1e-4.
#ifndef B2S_K3H_THREADS
#define B2S_K3H_THREADS 128
#endif
constexpr int kScoreHThreads = B2S_K3H_THREADS;
constexpr int kScoreHChunk = 512;   // 8 KB of shared memory: the usual pair (<= 500 correspondences) is staged once, in the pass that also finds W1 / W2

// B2S_K3H_PACK2: a thread evaluates TWO correspondences per step on packed fma.rn.f32x2 (SASS FFMA2): the same 19
// float32 operations per evaluation in the same order (identical values, identical band), half the issue slots.
// Shared memory then holds correspondences 2k and 2k+1 as (x0, x1, y0, y1) (u0, u1, v0, v1), so that one
// LDS.128 delivers two aligned register pairs and the loop needs no register moves.
#ifndef B2S_K3H_PACK2
#define B2S_K3H_PACK2 1
#endif
// B2S_K3H_GROUPS: blocks of 32 evaluations between two passes over the undecidable ones (1, 2 or 4).
#ifndef B2S_K3H_GROUPS
#define B2S_K3H_GROUPS 2
#endif
constexpr int kK3hGroups = B2S_K3H_GROUPS;
constexpr int kK3hSpan = 32 * kK3hGroups;
__device__ __forceinline__ void k3h_stage(float4* s_p, int m, const float4 c) {
#if B2S_K3H_PACK2
  float* f = reinterpret_cast<float*>(s_p) + (m >> 1) * 8 + (m & 1);
  f[0] = c.x; f[2] = c.y; f[4] = c.z; f[6] = c.w;
#else
  s_p[m] = c;
#endif
}
__device__ __forceinline__ float4 k3h_staged(const float4* s_p, int m) {
#if B2S_K3H_PACK2
  const float* f = reinterpret_cast<const float*>(s_p) + (m >> 1) * 8 + (m & 1);
  return make_float4(f[0], f[2], f[4], f[6]);
#else
  return s_p[m];
#endif
}

// Winner-only scoring (b2s_ransac_winner_batched) evaluates a hypothesis in two pieces: every hypothesis on the
// first screen_len(M) correspondences, and only those that can still win on the rest.
__host__ __device__ inline int screen_len(int M) { return M <= 64 ? M : min(M, ((3 * M) / 8 + 31) & ~31); }

// MODE 0: all correspondences (or the gridDim.z slice), counts written / atomically added.
// MODE 1: the screening prefix [0, screen_len(M)), counts written.
// MODE 2: the rest [screen_len(M), M) for the hypotheses listed in `list` (n_list[pair] of them per pair), added
//         to the count the screening pass left (single writer per hypothesis).
#ifndef B2S_K3H_MIN_CTAS
#define B2S_K3H_MIN_CTAS 7   // 72 registers; left to itself ptxas takes 140-200 for the unrolled spans and runs 10 % slower
#endif
template <int MODE>
__global__ void __launch_bounds__(kScoreHThreads, B2S_K3H_MIN_CTAS) ransac_score_hybrid_kernel(
    const float4* __restrict__ corr, const int32_t* __restrict__ c_off, const int32_t* __restrict__ c_count,
    const double* __restrict__ E, int H, double th2_all, const double* __restrict__ th2_pp,
    int32_t* __restrict__ counts, const int32_t* __restrict__ list, const int32_t* __restrict__ n_list,
    int h_first, int h_end, const int32_t* __restrict__ early) {
  __shared__ float4 s_p[kScoreHChunk];
  __shared__ float s_w[2];
  const int pair = blockIdx.y;
  if (early && early[pair] >= 0) return;   // winner-only: this pair's answer is already known (uniform for the CTA)
  int h = h_first + blockIdx.x * kScoreHThreads + threadIdx.x;   // MODE 0 / 1 score the hypotheses [h_first, h_end)
  int n_h = h_end;
  if (MODE == 2) {
    h -= h_first;
    n_h = n_list[pair];
    if ((int)blockIdx.x * kScoreHThreads >= n_h) return;   // uniform for the CTA
    h = h < n_h ? list[(size_t)pair * H + h] : 0;
  }
  // gridDim.z > 1: the pair's correspondences are cut into gridDim.z slices (multiples of 32) and the counts
  // are summed with atomicAdd into a zeroed array — a lone pair with thousands of correspondences (BASELINE
  // config #4: 4096 hypotheses x ~7000 matches = 32 CTAs otherwise) then fills the machine.  The rounding bound
  // only needs W1 / W2 >= the norms of the correspondences this CTA evaluates, so the slice's own maxima serve.
  int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  const float4* cp = corr + c_off[pair];
  if (MODE == 1) M = screen_len(M);
  if (MODE == 2) {
    const int lo = screen_len(M);
    M -= lo;
    cp += lo;
    if (M <= 0) return;   // uniform for the CTA: the screening pass already saw everything
  }
  if (MODE == 0 && gridDim.z > 1) {
    const int per = (((M + (int)gridDim.z - 1) / (int)gridDim.z) + 31) & ~31;
    const int lo = min(M, (int)blockIdx.z * per);
    M = min(M, lo + per) - lo;
    cp += lo;
    if (M <= 0) return;   // uniform for the CTA
  }
  const double th2d = th2_pp ? th2_pp[pair] : th2_all;
  const float th2 = (float)th2d;
  const bool live = MODE == 2 ? (int)(blockIdx.x * kScoreHThreads + threadIdx.x) < n_h : h < h_end;
  double ed[9];
  float e[9];
  float n2 = 0.0f;
  {
    const double* ep = E + ((size_t)pair * H + (live ? h : h_first)) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      ed[k] = ep[k];
      e[k] = (float)ed[k];
      n2 = fmaf(e[k], e[k], n2);
    }
  }
  // W1^2, W2^2 over the pair's correspondences (block reduction through shared memory)
  if (threadIdx.x < 2) s_w[threadIdx.x] = 0.0f;
  __syncthreads();
  {
    float w1 = 0.0f, w2 = 0.0f;
    for (int m = threadIdx.x; m < M; m += kScoreHThreads) {
      const float4 c = __ldg(cp + m);
      if (m < kScoreHChunk) k3h_stage(s_p, m, c);   // the first chunk is staged by the same pass
      w1 = fmaxf(w1, fmaf(c.x, c.x, fmaf(c.y, c.y, 1.0f)));
      w2 = fmaxf(w2, fmaf(c.z, c.z, fmaf(c.w, c.w, 1.0f)));
    }
    {  // NaN padding of the first chunk's last group (see the chunk loop)
      const int n0 = min(kScoreHChunk, M), n0_32 = (n0 + 31) & ~31;
      const float qn = __int_as_float(0x7FC00000);
      for (int m = n0 + threadIdx.x; m < n0_32; m += kScoreHThreads) k3h_stage(s_p, m, make_float4(qn, qn, qn, qn));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      w1 = fmaxf(w1, __shfl_xor_sync(0xFFFFFFFFu, w1, o));
      w2 = fmaxf(w2, __shfl_xor_sync(0xFFFFFFFFu, w2, o));
    }
    if ((threadIdx.x & 31) == 0) {  // non-negative floats order like their bit patterns
      atomicMax(reinterpret_cast<int*>(&s_w[0]), __float_as_int(w1));
      atomicMax(reinterpret_cast<int*>(&s_w[1]), __float_as_int(w2));
    }
  }
  __syncthreads();
  constexpr float kU = 5.9604645e-8f * 1.0001f;                     // 2^-24, nudged up: the bound itself is rounded
  const float nE = sqrtf(n2) * 1.0001f, W1 = sqrtf(s_w[0]) * 1.0001f, W2 = sqrtf(s_w[1]) * 1.0001f;
#ifndef B2S_K3H_BOUND_SCALE
#define B2S_K3H_BOUND_SCALE 1.0f   // diagnostics only (tools/k3h_check.py): < 1 shrinks the band below what is proven
#endif
  const float delta = B2S_K3H_BOUND_SCALE * 7.0f * kU * nE * W1 * W2;
  const float two_delta_raw = 2.0f * delta;
  const float c0_raw = delta * delta + th2 * (B2S_K3H_BOUND_SCALE * 9.0f * kU * nE * nE * (W1 * W1 + W2 * W2));
  const float rho = B2S_K3H_BOUND_SCALE * 10.0f * kU;                                     // relative part, applied to num^2 + th^2 den
  // The loop does not form q = num^2 on its own: with q + T = d + 2T the relative part is
  // rho (q + T) <= rho |d| + 2 rho T, so |d| > B is implied by |d| (1 - rho) > 2 delta |num| + 2 rho T + c0,
  // i.e. by |d| > (1 + 2 rho)(2 delta |num| + 2 rho T + c0).  The three coefficients below carry that
  // factor (and another 1.001 for the rounding of B itself): two FMA-class instructions fewer per evaluation.
  const float infl = (1.0f + 2.0f * rho) * 1.001f;
  const float rho2 = 2.0f * rho * infl;
  const float two_delta = two_delta_raw * infl, c0 = c0_raw * infl;
  // T = th^2 den never exceeds Tmax = th^2 nE^2 (W1^2 + W2^2) (den = |E x1|^2_xy + |E^T x2|^2_xy): the small
  // relative term 2 rho T is replaced by its bound, which makes B a single FMA and the band ~10 % wider
  // (2 rho Tmax is a tenth of the dominant 2 delta |num| at the threshold).
  const float kappa0 = fmaf(rho2, th2 * nE * nE * (W1 * W1 + W2 * W2) * 1.01f, c0);
  // The inner loop is branch-free: it counts the float32-certain inliers and records the
  // undecidable evaluations of a group of 32 correspondences in a bit mask; the float64
  // re-evaluation runs once per group over the set bits.  (Taking the detour inside the loop made
  // every warp diverge on ~6 % of its steps — 32 lanes x 1.8e-3 on tracking data — and each detour
  // is a ~25-deep dependent DFMA chain: 0.40 ms per bench step, against 0.30 ms this way.)
  const float qnan = __int_as_float(0x7FC00000);
#if B2S_K3H_PACK2
  float2 e2[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) e2[k] = make_float2(e[k], e[k]);
  const float2 nth2_2 = make_float2(-th2, -th2), ntwo_delta2 = make_float2(-two_delta, -two_delta), nkappa2 = make_float2(-kappa0, -kappa0);
#endif
  int count = 0;
  for (int base = 0; base < M; base += kScoreHChunk) {
    const int n = min(kScoreHChunk, M - base);
    const int n32 = (n + 31) & ~31;
    if (base > 0) {   // chunk 0 was staged by the prologue (its barrier is the one after the W1 / W2 reduction)
      __syncthreads();
      // the tail of the last group is padded with NaN: never a certain inlier, and its bits are masked off
      for (int m = threadIdx.x; m < n32; m += kScoreHThreads)
        k3h_stage(s_p, m, m < n ? __ldg(cp + base + m) : make_float4(qnan, qnan, qnan, qnan));
      __syncthreads();
    }
    auto span = [&](auto gc, const int m0) {
      constexpr int G = decltype(gc)::value;
      uint32_t band[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int mg = m0 + 32 * g;
        band[g] = 0u;
#if B2S_K3H_PACK2
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 xy = s_p[mg + 2 * j], uv = s_p[mg + 2 * j + 1];
          const float2 x = make_float2(xy.x, xy.y), y = make_float2(xy.z, xy.w);
          const float2 u = make_float2(uv.x, uv.y), v = make_float2(uv.z, uv.w);
          const float2 a0 = __ffma2_rn(e2[0], x, __ffma2_rn(e2[1], y, e2[2]));
          const float2 a1 = __ffma2_rn(e2[3], x, __ffma2_rn(e2[4], y, e2[5]));
          const float2 a2 = __ffma2_rn(e2[6], x, __ffma2_rn(e2[7], y, e2[8]));
          const float2 b0 = __ffma2_rn(e2[0], u, __ffma2_rn(e2[3], v, e2[6]));
          const float2 b1 = __ffma2_rn(e2[1], u, __ffma2_rn(e2[4], v, e2[7]));
          const float2 num = __ffma2_rn(u, a0, __ffma2_rn(v, a1, a2));
          const float2 den = __ffma2_rn(a0, a0, __ffma2_rn(a1, a1, __ffma2_rn(b0, b0, __fmul2_rn(b1, b1))));
          const float2 nT = __fmul2_rn(nth2_2, den);                          // -(th^2 den): the sign costs no rounding
          const float2 d = __ffma2_rn(num, num, nT);
          const float2 nB = __ffma2_rn(ntwo_delta2, make_float2(fabsf(num.x), fabsf(num.y)), nkappa2);   // -B
          asm("{\n\t.reg .pred p, q;\n\t"
              "setp.lt.f32 p, %2, %3;\n\t"
              "@p add.s32 %0, %0, 1;\n\t"
              "setp.gt.f32 q, %4, %5;\n\t"
              "@!q or.b32 %1, %1, %6;\n\t}"
              : "+r"(count), "+r"(band[g])
              : "f"(d.x), "f"(nB.x), "f"(fabsf(d.x)), "f"(-nB.x), "r"(1u << (2 * j)));
          asm("{\n\t.reg .pred p, q;\n\t"
              "setp.lt.f32 p, %2, %3;\n\t"
              "@p add.s32 %0, %0, 1;\n\t"
              "setp.gt.f32 q, %4, %5;\n\t"
              "@!q or.b32 %1, %1, %6;\n\t}"
              : "+r"(count), "+r"(band[g])
              : "f"(d.y), "f"(nB.y), "f"(fabsf(d.y)), "f"(-nB.y), "r"(2u << (2 * j)));
        }
#else
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 c = s_p[mg + j];
          const float a0 = fmaf(e[0], c.x, fmaf(e[1], c.y, e[2]));
          const float a1 = fmaf(e[3], c.x, fmaf(e[4], c.y, e[5]));
          const float a2 = fmaf(e[6], c.x, fmaf(e[7], c.y, e[8]));
          const float b0 = fmaf(e[0], c.z, fmaf(e[3], c.w, e[6]));
          const float b1 = fmaf(e[1], c.z, fmaf(e[4], c.w, e[7]));
          const float num = fmaf(c.z, a0, fmaf(c.w, a1, a2));
          const float den = fmaf(a0, a0, fmaf(a1, a1, fmaf(b0, b0, b1 * b1)));
          const float T = th2 * den;
          const float d = fmaf(num, num, -T);                              // one rounding fewer than the bound allows for
          const float B = fmaf(two_delta, fabsf(num), kappa0);             // see the notes on q + T and on T <= Tmax above the loop
          // count += d < -B (certain inlier); band |= bit unless |d| > B (undecidable in float32, or not
          // finite).  Two predicated instructions; the compiler's own form took 3.5 per step.
          asm("{\n\t.reg .pred p, q;\n\t"
              "setp.lt.f32 p, %2, %3;\n\t"
              "@p add.s32 %0, %0, 1;\n\t"
              "setp.gt.f32 q, %4, %5;\n\t"
              "@!q or.b32 %1, %1, %6;\n\t}"
              : "+r"(count), "+r"(band[g])
              : "f"(d), "f"(-B), "f"(fabsf(d)), "f"(B), "r"(1u << j));
        }
#endif
        const int left = n - mg;
        if (left < 32) band[g] &= left > 0 ? (1u << left) - 1u : 0u;
      }
      // The float64 expression of K3 decides.  One pass per 32 G evaluations: a warp walks max-over-lanes set bits,
      // and the undecidable evaluations (~2 per 1024) of a longer span mostly sit on different lanes.
      uint64_t lo = band[0], hi = 0;
      if (G > 1) lo |= (uint64_t)band[G > 1 ? 1 : 0] << 32;
      if (G > 2) hi = band[G > 2 ? 2 : 0] | ((uint64_t)band[G > 2 ? 3 : 0] << 32);
      while (lo | hi) {
        int j;
        if (lo) {
          j = __ffsll((long long)lo) - 1;
          lo &= lo - 1;
        } else {
          j = 64 + __ffsll((long long)hi) - 1;
          hi &= hi - 1;
        }
        const float4 c = k3h_staged(s_p, m0 + j);
        count += sampson_inlier<double>(ed, (double)c.x, (double)c.y, (double)c.z, (double)c.w, th2d) ? 1 : 0;
      }
    };
    int m0 = 0;
    for (; m0 + kK3hSpan <= n32; m0 += kK3hSpan) span(std::integral_constant<int, kK3hGroups>{}, m0);
    if (kK3hGroups > 1)
      for (; m0 < n32; m0 += 32) span(std::integral_constant<int, 1>{}, m0);   // what is left of the chunk, one group at a time
  }
  if (live) {
    if (MODE == 2) counts[(size_t)pair * H + h] += count;
    else if (MODE == 0 && gridDim.z > 1) atomicAdd(&counts[(size_t)pair * H + h], count);
    else counts[(size_t)pair * H + h] = count;
  }
}

// ---- winner-only: which hypotheses can still win ---------------------------------------------------------
// Step 1: the first kWinnerChunk0 hypotheses are scored completely; if one of them exceeds 0.8 M the lowest such
// index is the reference's answer and the pair is finished (ransac_early_kernel).  On clean data (most minimal
// samples all-inlier) that is nearly every pair, at 128 / H of the full cost.
// Step 2, for the other pairs, one CTA per pair (ransac_bound_kernel) over the hypotheses [kWinnerChunk0, H):
// c1[h] = inliers among the first m1 = screen_len(M) correspondences, so
// c1[h] <= count[h] <= ub[h] = c1[h] + (M - m1).  The kBoundTop hypotheses with the largest c1 are finished here
// (float64, the rest of the correspondences) and give L = the largest COMPLETE count.  A hypothesis matters to
// the reference's rule (first h above 0.8 M, else the lowest h among the maximum, homography.py:335-339) only if
// it can exceed 0.8 M or reach the maximum; ub[h] < min(L, floor(0.8 M) + 1) rules out both (the maximum is >= L).
// Everything else is listed (ascending h) for the finishing pass.  The abandoned hypotheses keep c1[h] in
// `counts`: a lower bound that is <= 0.8 M and < L, so the ordinary winner kernel run over ALL of `counts`
// returns exactly the winner of the full evaluation.
constexpr int kBoundThreads = 256;
constexpr int kBoundTop = 8;
constexpr int kWinnerChunk0 = 128;   // hypotheses scored to the end before anything else (winner-only mode)

// After the first kWinnerChunk0 hypotheses were scored completely: the first of them above 0.8 M (if any) IS the
// reference's answer — its loop stops there (homography.py:338-339) — and nothing else needs evaluating for this pair.
// early[pair] = that index or -1; chunk_max[pair] = the largest complete count so far (a lower bound of the maximum).
__global__ void __launch_bounds__(128) ransac_early_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ c_count,
                                                           int H, int h_end, int32_t* __restrict__ early, int32_t* __restrict__ chunk_max) {
  __shared__ int s_first, s_max;
  const int pair = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s_first = 0x7FFFFFFF, s_max = 0;
  __syncthreads();
  const double thr = 0.8 * (double)max(c_count[pair], 0);
  int first = 0x7FFFFFFF, mx = 0;
  for (int h = tid; h < h_end; h += blockDim.x) {
    const int c = counts[(size_t)pair * H + h];
    if ((double)c > thr) first = min(first, h);
    mx = max(mx, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, o));
    mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }
  if ((tid & 31) == 0) {
    atomicMin(&s_first, first);
    atomicMax(&s_max, mx);
  }
  __syncthreads();
  if (tid == 0) {
    early[pair] = s_first == 0x7FFFFFFF ? -1 : s_first;
    chunk_max[pair] = s_max;
  }
}

// Hypotheses [h0, H) after the screening pass (see the header of this section); L also takes the complete counts of
// the first chunk.
__global__ void __launch_bounds__(kBoundThreads) ransac_bound_kernel(
    const float4* __restrict__ corr, const int32_t* __restrict__ c_off, const int32_t* __restrict__ c_count,
    const double* __restrict__ E, int H, int h0, double th2_all, const double* __restrict__ th2_pp, int32_t* __restrict__ counts,
    int32_t* __restrict__ list, int32_t* __restrict__ n_list, const int32_t* __restrict__ early,
    const int32_t* __restrict__ chunk_max) {
  __shared__ unsigned long long s_red[kBoundThreads / 32];
  __shared__ unsigned long long s_top[kBoundTop];
  __shared__ int s_cnt[kBoundTop];
  __shared__ int s_scan[kBoundThreads / 32];
  __shared__ int s_need;
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = max(c_count[pair], 0), m1 = screen_len(M);
  int32_t* cnt = counts + (size_t)pair * H;
  const float4* cp = corr + c_off[pair];
  const double th2 = th2_pp ? th2_pp[pair] : th2_all;
  if (M == m1 || early[pair] >= 0 || h0 >= H) {   // nothing left to evaluate, or the answer is already known
    if (tid == 0) n_list[pair] = 0;
    return;
  }
  if (tid < kBoundTop) s_cnt[tid] = 0;
  // ---- the kBoundTop largest c1 (ties: lower h), kBoundTop rounds of a block arg-max below the previous key ----
  unsigned long long last = ~0ull;
  for (int r = 0; r < kBoundTop; ++r) {
    unsigned long long best = 0ull;
    for (int h = h0 + tid; h < H; h += kBoundThreads) {
      const unsigned long long key = ((unsigned long long)(uint32_t)cnt[h] << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)h);
      if (key < last && key > best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, best, o);
      best = v > best ? v : best;
    }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    best = s_red[0];
#pragma unroll
    for (int w = 1; w < kBoundThreads / 32; ++w) best = s_red[w] > best ? s_red[w] : best;
    if (tid == 0) s_top[r] = best;
    last = best;          // 0 when fewer than r + 1 hypotheses exist: later rounds find nothing
    __syncthreads();
  }
  // ---- finish them: the correspondences [m1, M), float64 ----
  for (int r = 0; r < kBoundTop; ++r) {
    const unsigned long long key = s_top[r];
    if (key == 0ull) break;                                   // uniform
    const int h = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    double e[9];
    const double* ep = E + ((size_t)pair * H + h) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = ep[k];
    int mine = 0;
    for (int m = m1 + tid; m < M; m += kBoundThreads) {
      const float4 c = cp[m];
      mine += sampson_inlier<double>(e, (double)c.x, (double)c.y, (double)c.z, (double)c.w, th2) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if (lane == 0 && mine) atomicAdd(&s_cnt[r], mine);
  }
  __syncthreads();
  if (tid == 0) {
    int L = chunk_max[pair];                                   // complete counts of the first chunk
    for (int r = 0; r < kBoundTop; ++r) {
      if (s_top[r] == 0ull) break;
      const int h = (int)(0xFFFFFFFFu - (uint32_t)(s_top[r] & 0xFFFFFFFFull));
      const int full = (int)(s_top[r] >> 32) + s_cnt[r];
      cnt[h] = full;                                           // complete
      L = max(L, full);
    }
    const int early_cnt = (int)floor(0.8 * (double)M) + 1;     // smallest count with count > 0.8 * M (float64, as the reference compares)
    s_need = min(L, early_cnt) - (M - m1);                     // survive iff c1[h] >= need
  }
  __syncthreads();
  // ---- ordered compaction of the survivors (the finished ones excluded) ----
  const int need = s_need;
  const int per = (H - h0 + kBoundThreads - 1) / kBoundThreads;
  const int ha = h0 + tid * per, hb = min(H, ha + per);
  int mine = 0;
  for (int h = ha; h < hb; ++h) {
    bool done = false;
#pragma unroll
    for (int r = 0; r < kBoundTop; ++r) done |= (s_top[r] != 0ull && (int)(0xFFFFFFFFu - (uint32_t)(s_top[r] & 0xFFFFFFFFull)) == h);
    mine += (!done && cnt[h] >= need) ? 1 : 0;
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += s_scan[w];
  int pos = base + incl - mine;
  for (int h = ha; h < hb; ++h) {
    bool done = false;
#pragma unroll
    for (int r = 0; r < kBoundTop; ++r) done |= (s_top[r] != 0ull && (int)(0xFFFFFFFFu - (uint32_t)(s_top[r] & 0xFFFFFFFFull)) == h);
    if (!done && cnt[h] >= need) list[(size_t)pair * H + pos++] = h;
  }
  if (tid == kBoundThreads - 1) n_list[pair] = base + incl;
}

// ---- winner selection + inlier mask ---------------------------------------------------
__global__ void __launch_bounds__(256) ransac_select_kernel(
    const int32_t* __restrict__ counts, const float4* __restrict__ corr, const int32_t* __restrict__ c_off,
    const int32_t* __restrict__ c_count, const double* __restrict__ E, int H, double th2_all,
    const double* __restrict__ th2_pp, int32_t* __restrict__ best_h, int32_t* __restrict__ best_count,
    uint8_t* __restrict__ mask, const int32_t* __restrict__ early_in, int h_chunk, const b2s_record_sink sink, int mask_stride) {
  __shared__ int s_early;
  __shared__ unsigned long long s_best;
  __shared__ int s_cnt;
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int H_all = H;
  if (early_in && early_in[pair] >= 0) H = min(H, h_chunk);   // winner-only: only the first chunk was evaluated for this pair
  const int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  if (tid == 0) {
    s_early = 0x7FFFFFFF;
    s_best = 0ull;
    s_cnt = 0;
  }
  __syncthreads();
  const double early_thr = 0.8 * (double)M;  // "inliers.size > 0.8 * n" in float64
  int early = 0x7FFFFFFF;
  unsigned long long bestk = 0ull;
  for (int h = tid; h < H; h += blockDim.x) {
    const int c = counts[(size_t)pair * H_all + h];
    if ((double)c > early_thr) early = min(early, h);
    // larger count wins; among equal counts the lower h (first strict improvement) wins
    const unsigned long long k = ((unsigned long long)(uint32_t)c << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)h);
    bestk = max(bestk, k);
  }
  if (early != 0x7FFFFFFF) atomicMin(&s_early, early);
  atomicMax(&s_best, bestk);
  __syncthreads();
  int win;
  if (s_early != 0x7FFFFFFF) win = s_early;
  else win = ((s_best >> 32) == 0ull) ? -1 : (int)(0xFFFFFFFFu - (uint32_t)(s_best & 0xFFFFFFFFull));

  const float4* cp = corr + c_off[pair];
  uint8_t* mp = mask + c_off[pair];
  int mine = 0;
  if (win >= 0) {
    double e[9];
    const double* ep = E + ((size_t)pair * H_all + win) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = ep[k];
    const double th2 = th2_pp ? th2_pp[pair] : th2_all;
    for (int m = tid; m < M; m += blockDim.x) {
      const float4 c = cp[m];
      const bool in = sampson_inlier<double>(e, (double)c.x, (double)c.y, (double)c.z, (double)c.w, th2);
      mp[m] = in ? 1 : 0;
      mine += in ? 1 : 0;
    }
  } else {
    for (int m = tid; m < M; m += blockDim.x) mp[m] = 0;
  }
  for (int m = M + tid; m < mask_stride; m += blockDim.x) mp[m] = 0;   // the unused tail of the pair's compact slot
  if (mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (tid == 0) {
    best_h[pair] = win;
    best_count[pair] = s_cnt;
  }
  if (sink.records) {
    // the pair's result record, written by the kernel that already holds everything it needs (csrc/records.cu has
    // the layout and the stand-alone pack kernel used when R | t are appended later): no extra launch at the end of
    // the step, and this IS the send slice of the all-gather / the source of the step's one device->host copy
    uint8_t* rec = sink.records + (size_t)pair * sink.record_bytes;
    const int S = sink.stride;
    const int raw = c_count[pair], n = min(M, S);
    if (tid < 16) {
      int32_t v = 0;
      if (tid == 0) v = raw;
      else if (tid == 1) v = win;
      else if (tid == 2) v = s_cnt;
      else if (tid == 3) v = sink.pair_id0 + pair;
      reinterpret_cast<int32_t*>(rec)[tid] = v;   // R | t stay zero: no pose recovery on this path
    }
    uint16_t* rq = reinterpret_cast<uint16_t*>(rec + 64);
    uint16_t* rt = rq + S;
    uint16_t* rd = rt + S;
    uint8_t* rm = reinterpret_cast<uint8_t*>(rd + S);
    const size_t base = (size_t)pair * S;
    for (int k = tid; k < S; k += blockDim.x) {
      const bool live = k < n;
      rq[k] = live ? (uint16_t)sink.out_q[base + k] : (uint16_t)0;
      rt[k] = live ? (uint16_t)sink.out_t[base + k] : (uint16_t)0;
      rd[k] = live ? (uint16_t)sink.out_d[base + k] : (uint16_t)0;
      rm[k] = live ? mp[k] : (uint8_t)0;          // written above by this CTA (the barrier orders it)
    }
  }
}

// ---- K4: 8-point minimal solver, one thread per hypothesis ---------------------------
constexpr int kEpThreads = 128;

// AFFINE: the last row of Kinv is (0 0 1), so the dehomogenising divisions are by exactly 1.
// KEYE: K = I, so the K^T F K product is the identity map.  Both are decided on the host.
template <bool AFFINE, bool KEYE>
__global__ void __launch_bounds__(kEpThreads, 4) eight_point_kernel(
    const float4* __restrict__ corr, const int32_t* __restrict__ c_off, const int32_t* __restrict__ c_count, int H,
    const int32_t* __restrict__ samples_in, uint64_t seed, int32_t* __restrict__ samples_out, const Mat3 K,
    const Mat3 Kinv, double* __restrict__ E_out, const int32_t* __restrict__ pair_ids, int pair_id0) {
  const int pair = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  const float4* cp = corr + c_off[pair];
  double* eo = E_out + ((size_t)pair * H + h) * 9;
  int idx[8];
  if (samples_in) {
    for (int k = 0; k < 8; ++k) idx[k] = samples_in[((size_t)pair * H + h) * 8 + k];
  } else if (M >= 8) {
    // the stream is keyed by the pair's GLOBAL id, so a pair draws the same samples whichever rank / batch
    // position it is processed at (pair-sharded runs equal the single-GPU run bit for bit)
    draw_distinct<8>(seed, pair_ids ? pair_ids[pair] : pair_id0 + pair, h, M, idx);
  } else {
    for (int k = 0; k < 8; ++k) idx[k] = 0;
  }
  if (samples_out)
    for (int k = 0; k < 8; ++k) samples_out[((size_t)pair * H + h) * 8 + k] = idx[k];
  bool ok = M >= 8;
  for (int k = 0; k < 8; ++k) ok &= (idx[k] >= 0 && idx[k] < M);
  if (!ok) {
    for (int k = 0; k < 9; ++k) eo[k] = 0.0;
    return;
  }

  double A[8][9];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 c = cp[idx[k]];
    const double sx = c.x, sy = c.y, dx = c.z, dy = c.w;
    // x = Kinv [sx sy 1]^T, then dehomogenise (homography.py:228-237)
    const double* ki = Kinv.m;
    double x = fma(ki[0], sx, fma(ki[1], sy, ki[2]));
    double y = fma(ki[3], sx, fma(ki[4], sy, ki[5]));
    double u = fma(ki[0], dx, fma(ki[1], dy, ki[2]));
    double v = fma(ki[3], dx, fma(ki[4], dy, ki[5]));
    if (!AFFINE) {
      const double w1 = fma(ki[6], sx, fma(ki[7], sy, ki[8]));
      const double w2 = fma(ki[6], dx, fma(ki[7], dy, ki[8]));
      x /= w1;
      y /= w1;
      u /= w2;
      v /= w2;
    }
    A[k][0] = u * x; A[k][1] = u * y; A[k][2] = u;
    A[k][3] = v * x; A[k][4] = v * y; A[k][5] = v;
    A[k][6] = x;     A[k][7] = y;     A[k][8] = 1.0;
  }
  double f[9];
  null_vector_8x9(A, f);

  // rank-2 projection: F' = F - (F v3) v3^T, v3 = right-singular vector of the smallest
  // singular value (homography.py:244-246 zeroes S[2] only).
  double v3[3];
  smallest_right_singular3(f, v3);
  double Fp[9];
  for (int r = 0; r < 3; ++r) {
    const double fv = fma(f[3 * r], v3[0], fma(f[3 * r + 1], v3[1], f[3 * r + 2] * v3[2]));
    for (int c = 0; c < 3; ++c) Fp[3 * r + c] = fma(-fv, v3[c], f[3 * r + c]);
  }
  if (KEYE) {
#pragma unroll
    for (int k = 0; k < 9; ++k) eo[k] = Fp[k];
    return;
  }
  // E = K^T F' K  (homography.py:248)
  double T1[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      T1[3 * r + c] = fma(Fp[3 * r], K.m[c], fma(Fp[3 * r + 1], K.m[3 + c], Fp[3 * r + 2] * K.m[6 + c]));
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      eo[3 * r + c] = fma(K.m[r], T1[c], fma(K.m[3 + r], T1[3 + c], K.m[6 + r] * T1[6 + c]));
}

int ransac_score_fp64_cond_launch(const float4* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs, const double* E,
                                  int H, double th2, const double* th2_pp, int32_t* counts, const int* run_flag, cudaStream_t st) {
  dim3 grid((H + kScoreThreads - 1) / kScoreThreads, n_pairs);
  ransac_score_kernel<double><<<grid, kScoreThreads, 0, st>>>(corr, c_off, c_count, E, H, th2, th2_pp, counts, run_flag);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

static int check_sink(const b2s_record_sink* sink, int mask_stride, b2s_record_sink* out) {
  B2S_REQUIRE(mask_stride >= 0, "negative mask_stride");
  if (!sink || !sink->records) return B2S_OK;
  B2S_REQUIRE(sink->out_q && sink->out_t && sink->out_d && sink->stride > 0, "record sink: selection outputs and stride required");
  B2S_REQUIRE(sink->record_bytes >= b2s_record_bytes(sink->stride) && (sink->record_bytes & 15u) == 0 &&
                  ((uintptr_t)sink->records & 15u) == 0, "record sink: record_bytes >= b2s_record_bytes(stride), 16-byte aligned");
  *out = *sink;
  return B2S_OK;
}

}  // namespace b2s

extern "C" {

int b2s_eight_point_batched(const float* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs, int H,
                            const int32_t* samples_in, uint64_t seed, const int32_t* pair_ids, int pair_id0,
                            int32_t* samples_out, const double* K_host, const double* Kinv_host, double* E_out,
                            void* stream) {
  using namespace b2s;
  B2S_REQUIRE(corr && c_off && c_count && E_out, "null pointer");
  B2S_REQUIRE(n_pairs >= 0 && H >= 0, "negative size");
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  B2S_REQUIRE((K_host == nullptr) == (Kinv_host == nullptr), "pass both K and Kinv or neither");
  if (n_pairs == 0 || H == 0) return B2S_OK;
  Mat3 K, Kinv;
  for (int i = 0; i < 9; ++i) {
    K.m[i] = K_host ? K_host[i] : ((i % 4 == 0) ? 1.0 : 0.0);
    Kinv.m[i] = Kinv_host ? Kinv_host[i] : ((i % 4 == 0) ? 1.0 : 0.0);
  }
  dim3 grid((H + kEpThreads - 1) / kEpThreads, n_pairs);
  const bool affine = Kinv.m[6] == 0.0 && Kinv.m[7] == 0.0 && Kinv.m[8] == 1.0;
  bool keye = true;
  for (int i = 0; i < 9; ++i) keye &= (K.m[i] == ((i % 4 == 0) ? 1.0 : 0.0));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float4* c4 = reinterpret_cast<const float4*>(corr);
#define B2S_EP_LAUNCH(A_, K_) \
  eight_point_kernel<A_, K_><<<grid, kEpThreads, 0, st>>>(c4, c_off, c_count, H, samples_in, seed, samples_out, K, Kinv, E_out, pair_ids, pair_id0)
  if (affine && keye) B2S_EP_LAUNCH(true, true);
  else if (affine) B2S_EP_LAUNCH(true, false);
  else B2S_EP_LAUNCH(false, false);
#undef B2S_EP_LAUNCH
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

int b2s_ransac_score_batched(const float* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs,
                             const double* E, int H, double th2, const double* th2_per_pair, int precision,
                             int max_m, int32_t* counts, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(corr && c_off && c_count && E && counts, "null pointer");
  B2S_REQUIRE(precision == 64 || precision == 32 || precision == 6464, "precision must be 64, 32 or 6464");
  B2S_REQUIRE(n_pairs >= 0 && H >= 0, "negative size");
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  if (n_pairs == 0 || H == 0) return B2S_OK;
  dim3 grid((H + kScoreThreads - 1) / kScoreThreads, n_pairs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float4* c4 = reinterpret_cast<const float4*>(corr);
  if (precision == 64) {
    // few pairs with many correspondences each: slice the correspondences too (see the kernel)
    grid.x = (H + kScoreHThreads - 1) / kScoreHThreads;
    const long blocks = (long)grid.x * grid.y, sms = sm_count();
    int zs = 1;
    if (max_m > 2 * kScoreHChunk && blocks < sms) {
      const long want = (2 * sms + blocks - 1) / blocks, most = (max_m + kScoreHChunk - 1) / kScoreHChunk;
      zs = (int)(want < most ? want : most);
    }
    if (zs > 1) {
      B2S_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_pairs * H, st));
      grid.z = zs;
    }
    ransac_score_hybrid_kernel<0><<<grid, kScoreHThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts, nullptr, nullptr, 0, H, nullptr);
  } else if (precision == 6464)
    ransac_score_kernel<double><<<grid, kScoreThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts, nullptr);
  else
    ransac_score_kernel<float><<<grid, kScoreThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts, nullptr);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

size_t b2s_ransac_winner_workspace_bytes(int n_pairs, int H) {
  if (n_pairs <= 0 || H <= 0) return 0;
  // counts | survivor list | survivors per pair | early index per pair | first-chunk maximum per pair
  return sizeof(int32_t) * ((size_t)n_pairs * H * 2 + 3 * (size_t)n_pairs) + 256;
}

int b2s_ransac_winner_batched(const float* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs, const double* E,
                              int H, double th2, const double* th2_per_pair, int32_t* best_h, int32_t* best_count,
                              uint8_t* inlier_mask, void* workspace, size_t workspace_bytes, int32_t* counts_out,
                              int32_t* n_finished_out, const b2s_record_sink* sink, int mask_stride, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(corr && c_off && c_count && E && best_h && best_count && inlier_mask, "null pointer");
  b2s_record_sink sk{};
  if (int rc = check_sink(sink, mask_stride, &sk)) return rc;
  B2S_REQUIRE(n_pairs >= 0 && H >= 0, "negative size");
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  if (n_pairs == 0) return B2S_OK;
  B2S_REQUIRE(workspace && workspace_bytes >= b2s_ransac_winner_workspace_bytes(n_pairs, H) && ((uintptr_t)workspace & 15u) == 0,
              "winner-only scoring needs b2s_ransac_winner_workspace_bytes(n_pairs, H) bytes, 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float4* c4 = reinterpret_cast<const float4*>(corr);
  int32_t* ws = static_cast<int32_t*>(workspace);
  int32_t* counts = counts_out ? counts_out : ws;
  int32_t* list = ws + (size_t)n_pairs * H;
  int32_t* n_list = n_finished_out ? n_finished_out : list + (size_t)n_pairs * H;
  int32_t* early = list + (size_t)n_pairs * H + n_pairs;
  int32_t* chunk_max = early + n_pairs;
  const int h0 = H < kWinnerChunk0 ? H : kWinnerChunk0;
  if (H > 0) {
    if (counts_out) B2S_CUDA(cudaMemsetAsync(counts_out, 0, sizeof(int32_t) * (size_t)n_pairs * H, st));   // unevaluated entries read 0
    // 1. the first chunk to the end, and whether it already holds the answer
    dim3 g0((h0 + kScoreHThreads - 1) / kScoreHThreads, n_pairs);
    ransac_score_hybrid_kernel<0><<<g0, kScoreHThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts, nullptr, nullptr, 0, h0, nullptr);
    B2S_CUDA(cudaGetLastError());
    ransac_early_kernel<<<n_pairs, 128, 0, st>>>(counts, c_count, H, h0, early, chunk_max);
    B2S_CUDA(cudaGetLastError());
    note_launch(2);
    // 2. everything else: screen, bound, finish (pairs whose answer is known skip all three)
    if (H > h0) {
      dim3 grid((H - h0 + kScoreHThreads - 1) / kScoreHThreads, n_pairs);
      ransac_score_hybrid_kernel<1><<<grid, kScoreHThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts, nullptr, nullptr, h0, H, early);
      B2S_CUDA(cudaGetLastError());
      ransac_bound_kernel<<<n_pairs, kBoundThreads, 0, st>>>(c4, c_off, c_count, E, H, h0, th2, th2_per_pair, counts, list, n_list, early, chunk_max);
      B2S_CUDA(cudaGetLastError());
      ransac_score_hybrid_kernel<2><<<grid, kScoreHThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts, list, n_list, 0, H, early);
      B2S_CUDA(cudaGetLastError());
      note_launch(3);
    } else if (n_finished_out) {
      B2S_CUDA(cudaMemsetAsync(n_finished_out, 0, sizeof(int32_t) * (size_t)n_pairs, st));
    }
  }
  ransac_select_kernel<<<n_pairs, 256, 0, st>>>(counts, c4, c_off, c_count, E, H, th2, th2_per_pair, best_h, best_count, inlier_mask,
                                                H > 0 ? early : nullptr, h0, sk, mask_stride);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

int b2s_ransac_select(const int32_t* counts, const float* corr, const int32_t* c_off, const int32_t* c_count,
                      int n_pairs, const double* E, int H, double th2, const double* th2_per_pair, int32_t* best_h,
                      int32_t* best_count, uint8_t* inlier_mask, const b2s_record_sink* sink, int mask_stride, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(counts && corr && c_off && c_count && E && best_h && best_count && inlier_mask, "null pointer");
  b2s_record_sink sk{};
  if (int rc = check_sink(sink, mask_stride, &sk)) return rc;
  B2S_REQUIRE(n_pairs >= 0 && H >= 0, "negative size");
  if (n_pairs == 0) return B2S_OK;
  ransac_select_kernel<<<n_pairs, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      counts, reinterpret_cast<const float4*>(corr), c_off, c_count, E, H, th2, th2_per_pair, best_h, best_count,
      inlier_mask, nullptr, 0, sk, mask_stride);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

}  // extern "C"
