// linalg.cuh — small register-resident linear algebra shared by the minimal solvers
// (8-point essential matrix, 4-point homography) and the counter-based sampler.
#pragma once

#include "common.cuh"

namespace b2s {

__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Null vector of the 8x9 design matrix A: Householder QR of A^T (9x8, one column per sampled
// correspondence); the last column of Q = H0 H1 ... H7 spans the orthogonal complement of the
// rows of A.  Everything is statically indexed (registers only), there is no pivot search and no
// select chain, and no column has to be singled out as the "free" one (for forward motion
// E33 ~ 0, so a fixed free column would make the 8x8 system singular).  The reflector tails
// overwrite the entries they annihilate (LAPACK storage), so the whole solve lives in the 72
// registers of A plus 16 scalars.  The result has unit norm by construction.
// History (us per 592k hypotheses): matrix in local memory 947 -> shared memory + complete
// pivoting 386 -> registers + row-fixed column pivoting (select chains, 7.3k instructions per
// hypothesis) 328 -> this version.
__device__ __forceinline__ void null_vector_8x9(double (&A)[8][9], double (&n)[9]) {
  double v0[8], beta[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // column k of A^T below the diagonal = A[k][k..8]
    double s = 0.0;
#pragma unroll
    for (int i = k; i < 9; ++i) s = fma(A[k][i], A[k][i], s);
    const double nrm = s * rsqrt(s);                 // sqrt(s); NaN for s = 0, masked by beta below
    const double x0 = A[k][k];
    v0[k] = x0 + copysign(nrm, x0);
    beta[k] = (s > 0.0) ? 1.0 / fma(fabs(x0), nrm, s) : 0.0;  // 2 / (v^T v)
    if (!(s > 0.0)) v0[k] = 0.0;
#pragma unroll
    for (int j = k + 1; j < 8; ++j) {
      double d = v0[k] * A[j][k];
#pragma unroll
      for (int i = k + 1; i < 9; ++i) d = fma(A[k][i], A[j][i], d);
      d *= beta[k];
      A[j][k] = fma(-d, v0[k], A[j][k]);
#pragma unroll
      for (int i = k + 1; i < 9; ++i) A[j][i] = fma(-d, A[k][i], A[j][i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) n[i] = (i == 8) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    double d = v0[k] * n[k];
#pragma unroll
    for (int i = k + 1; i < 9; ++i) d = fma(A[k][i], n[i], d);
    d *= beta[k];
    n[k] = fma(-d, v0[k], n[k]);
#pragma unroll
    for (int i = k + 1; i < 9; ++i) n[i] = fma(-d, A[k][i], n[i]);
  }
}


// k distinct indices in [0, M) from a splitmix64 counter stream keyed by (seed, pair, hypothesis)
template <int KSAMP>
__device__ __forceinline__ void draw_distinct(uint64_t seed, int pair, int h, int M, int (&idx)[KSAMP]) {
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(pair + 1)) ^ (0x8CB92BA72F3D8DD7ull * (uint64_t)(h + 1));
  for (int k = 0; k < KSAMP; ++k) {
    int cand;
    bool dup;
    do {
      cand = (int)__umul64hi(splitmix64(s), (uint64_t)M);
      dup = false;
      for (int a = 0; a < k; ++a) dup |= (idx[a] == cand);
    } while (dup);
    idx[k] = cand;
  }
}

}  // namespace b2s
