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


// Right-singular vector of the smallest singular value of the 3x3 matrix f (row-major).
// cof(F) = s2 s3 u1 v1^T + s1 s3 u2 v2^T + s1 s2 u3 v3^T, so v3 is the DOMINANT eigenvector of
// M = cof(F)^T cof(F) with eigenvalue ratio (s3/s2)^2; repeated squaring of the trace-normalised
// M squares that ratio every step.  With tr(M) = 1 the trace of M^2 is 1 - 2*delta (delta = the
// second eigenvalue), so the loop stops after the squaring that saw delta < 1e-8: its result is
// rank one to 1e-16.  3-6 squarings for well-posed samples, 12+ only when s3/s2 > 0.99.
// No division, square root or trigonometry inside the loop (the cyclic Jacobi solver this
// replaces spent ~5k instructions per hypothesis on them).
__device__ __forceinline__ void smallest_right_singular3(const double (&f)[9], double (&v)[3]) {
  double c[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      c[3 * i + j] = fma(f[3 * i1 + j1], f[3 * i2 + j2], -f[3 * i1 + j2] * f[3 * i2 + j1]);
    }
  // symmetric M: m00 m01 m02 m11 m12 m22
  double m00 = fma(c[0], c[0], fma(c[3], c[3], c[6] * c[6]));
  double m01 = fma(c[0], c[1], fma(c[3], c[4], c[6] * c[7]));
  double m02 = fma(c[0], c[2], fma(c[3], c[5], c[6] * c[8]));
  double m11 = fma(c[1], c[1], fma(c[4], c[4], c[7] * c[7]));
  double m12 = fma(c[1], c[2], fma(c[4], c[5], c[7] * c[8]));
  double m22 = fma(c[2], c[2], fma(c[5], c[5], c[8] * c[8]));
  for (int it = 0; it < 40; ++it) {
    const double tr = m00 + m11 + m22;
    if (!(tr > 0.0)) break;                          // cof(F) = 0: F has rank <= 1, nothing to project
    const double inv = 1.0 / tr;
    m00 *= inv; m01 *= inv; m02 *= inv; m11 *= inv; m12 *= inv; m22 *= inv;
    const double n00 = fma(m00, m00, fma(m01, m01, m02 * m02));
    const double n01 = fma(m00, m01, fma(m01, m11, m02 * m12));
    const double n02 = fma(m00, m02, fma(m01, m12, m02 * m22));
    const double n11 = fma(m01, m01, fma(m11, m11, m12 * m12));
    const double n12 = fma(m01, m02, fma(m11, m12, m12 * m22));
    const double n22 = fma(m02, m02, fma(m12, m12, m22 * m22));
    m00 = n00; m01 = n01; m02 = n02; m11 = n11; m12 = n12; m22 = n22;
    if (1.0 - (n00 + n11 + n22) < 2e-8) break;
  }
  // the column with the largest diagonal entry is the best-conditioned copy of v3
  const bool use1 = m11 > m00;
  const double d01 = use1 ? m11 : m00;
  const bool use2 = m22 > d01;
  const double a = use2 ? m02 : (use1 ? m01 : m00);
  const double b = use2 ? m12 : (use1 ? m11 : m01);
  const double g = use2 ? m22 : (use1 ? m12 : m02);
  const double r = rsqrt(fma(a, a, fma(b, b, g * g)));
  v[0] = a * r;
  v[1] = b * r;
  v[2] = g * r;
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
