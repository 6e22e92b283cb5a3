// fivepoint.cu — K8: batched 5-point essential-matrix minimal solver.
//
// The reference's own RANSAC (/root/reference/homography.py:302-345) samples 8 points; its 5-point
// users are the cv2.findEssentialMat call sites (slam_viewer.py:195, web_dashboard_server.py:145,
// visual_slam_offline_entry_point.py:51), i.e. OpenCV's five-point.cpp (Nister 2004 / Stewenius
// 2006).  This kernel is the published algorithm in Nister's form, one thread per 5-sample:
//   1. null space of the 5x9 epipolar system (Householder QR of its transpose; the last four
//      columns of Q): E = x E0 + y E1 + z E2 + E3;
//   2. the ten cubic constraints det E = 0, E E^T E - 1/2 tr(E E^T) E = 0 as a 10x20 matrix over
//      the monomials [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
//      (kept in shared memory, one interleaved copy per thread), Gauss-Jordan with partial pivoting;
//   3. rows (x2z, x2), (y2z, y2), (xyz, xy) give B(z) [x y 1]^T = 0 with a 3x3 matrix of
//      polynomials; det B(z) is the degree-10 polynomial of the problem;
//   4. its real roots by a Sturm chain (exact root counting, bisection to isolation, bisection +
//      Newton to full precision) — no eigen-solver, no complex arithmetic;
//   5. x, y from the null vector of B(z); up to 10 unit-norm E per sample, unused slots zero.
// Points must be calibrated (K^-1-normalised), as cv2 normalises them before its solver.
// oracle/fivepoint_oracle.py restates the same algebra and is pinned against
// cv2.findEssentialMat on exactly five points (tests/golden/fivepoint_golden.npz).
#include "linalg.cuh"

namespace b2s {

constexpr int kFpThreads = 64;
constexpr int kFpMaxSol = 10;

__constant__ int8_t kQuadIdx[4][4] = {{0, 1, 2, 6}, {1, 3, 4, 7}, {2, 4, 5, 8}, {6, 7, 8, 9}};
__constant__ int8_t kCubIdx[10][4] = {{0, 2, 4, 5},     {2, 3, 8, 9},     {4, 8, 10, 11},  {3, 1, 6, 7},    {8, 6, 13, 14},
                                      {10, 13, 16, 17}, {5, 9, 11, 12},   {9, 7, 14, 15},  {11, 14, 17, 18}, {12, 15, 18, 19}};

// polynomials: coefficient arrays, LOWEST power first
__device__ __forceinline__ double horner(const double* p, int deg, double x) {
  double r = p[deg];
  for (int k = deg - 1; k >= 0; --k) r = fma(r, x, p[k]);
  return r;
}

// number of sign changes of the Sturm chain at x
__device__ int sturm_changes(const double (*S)[11], const int* deg, int n_chain, double x) {
  int changes = 0, last = 0;
  for (int k = 0; k < n_chain; ++k) {
    const double v = horner(S[k], deg[k], x);
    const int s = (v > 0.0) - (v < 0.0);
    if (s != 0) {
      if (last != 0 && s != last) ++changes;
      last = s;
    }
  }
  return changes;
}

// real roots of p (degree <= 10, lowest power first) -> roots[], returns their number
__device__ int real_roots_sturm(const double* p_in, int deg_in, double* roots) {
  double S[12][11];
  int deg[12];
  // strip vanishing leading coefficients
  double mx = 0.0;
  for (int k = 0; k <= deg_in; ++k) mx = fmax(mx, fabs(p_in[k]));
  if (!(mx > 0.0) || !isfinite(mx)) return 0;
  int d = deg_in;
  while (d > 0 && fabs(p_in[d]) <= 1e-14 * mx) --d;
  if (d == 0) return 0;
  for (int k = 0; k <= d; ++k) S[0][k] = p_in[k] / mx;
  deg[0] = d;
  for (int k = 1; k <= d; ++k) S[1][k - 1] = (double)k * S[0][k];
  deg[1] = d - 1;
  int n_chain = 2;
  while (deg[n_chain - 1] > 0 && n_chain < 12) {
    // r = -(S[n-2] mod S[n-1])
    const int da = deg[n_chain - 2], db = deg[n_chain - 1];
    double r[11];
    for (int k = 0; k <= da; ++k) r[k] = S[n_chain - 2][k];
    const double lead = S[n_chain - 1][db];
    for (int k = da; k >= db; --k) {
      const double f = r[k] / lead;
      for (int j = 0; j <= db; ++j) r[k - db + j] = fma(-f, S[n_chain - 1][j], r[k - db + j]);
      r[k] = 0.0;
    }
    int dr = db - 1;
    double rm = 0.0;
    for (int k = 0; k <= dr; ++k) rm = fmax(rm, fabs(r[k]));
    if (!(rm > 1e-300)) break;  // exact division: repeated roots; the chain so far still counts distinct roots
    while (dr > 0 && fabs(r[dr]) <= 1e-15 * rm) --dr;
    for (int k = 0; k <= dr; ++k) S[n_chain][k] = -r[k] / rm;   // positive scaling keeps the signs
    deg[n_chain] = dr;
    ++n_chain;
  }
  // Cauchy bound on the root moduli
  double R = 0.0;
  for (int k = 0; k < d; ++k) R = fmax(R, fabs(S[0][k] / S[0][d]));
  R += 1.0;
  // isolate by bisection on the sign-change count
  double lo_s[40], hi_s[40];
  int vlo_s[40], vhi_s[40];
  int sp = 0, n_roots = 0;
  lo_s[0] = -R;
  hi_s[0] = R;
  vlo_s[0] = sturm_changes(S, deg, n_chain, -R);
  vhi_s[0] = sturm_changes(S, deg, n_chain, R);
  sp = 1;
  while (sp > 0 && n_roots < kFpMaxSol) {
    --sp;
    double lo = lo_s[sp], hi = hi_s[sp];
    const int vlo = vlo_s[sp], vhi = vhi_s[sp];
    const int n = vlo - vhi;
    if (n <= 0) continue;
    if (n == 1 || (hi - lo) <= 1e-13 * R) {
      // one root (or an unresolvable cluster, taken as one): bisection on the sign of p, then Newton
      double flo = horner(S[0], d, lo);
      for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        const double fm = horner(S[0], d, mid);
        if ((fm > 0.0) == (flo > 0.0)) {
          lo = mid;
          flo = fm;
        } else {
          hi = mid;
        }
      }
      double x = 0.5 * (lo + hi);
      for (int it = 0; it < 3; ++it) {
        const double f = horner(S[0], d, x), fp = horner(S[1], d - 1, x);
        if (fp != 0.0) {
          const double xn = x - f / fp;
          if (xn >= lo - 1e-9 * R && xn <= hi + 1e-9 * R) x = xn;
        }
      }
      roots[n_roots++] = x;
      continue;
    }
    const double mid = 0.5 * (lo + hi);
    const int vm = sturm_changes(S, deg, n_chain, mid);
    if (sp + 2 > 40) continue;
    lo_s[sp] = mid, hi_s[sp] = hi, vlo_s[sp] = vm, vhi_s[sp] = vhi, ++sp;   // right half first on the stack,
    lo_s[sp] = lo, hi_s[sp] = mid, vlo_s[sp] = vlo, vhi_s[sp] = vm, ++sp;   // left half popped first: ascending roots
  }
  return n_roots;
}

__global__ void __launch_bounds__(kFpThreads) five_point_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ c_off,
                                                                const int32_t* __restrict__ c_count, int S,
                                                                const int32_t* __restrict__ samples_in, uint64_t seed,
                                                                int32_t* __restrict__ samples_out, double* __restrict__ E_out,
                                                                int32_t* __restrict__ n_sol) {
  extern __shared__ double s_A[];  // [10 * 20][kFpThreads]: this thread's constraint matrix, element-major
#define A_(r, c) s_A[((r) * 20 + (c)) * kFpThreads + threadIdx.x]
  const int pair = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  const float4* cp = corr + c_off[pair];
  double* eo = E_out + ((size_t)pair * S + s) * (kFpMaxSol * 9);
  int idx[5];
  if (samples_in) {
    for (int k = 0; k < 5; ++k) idx[k] = samples_in[((size_t)pair * S + s) * 5 + k];
  } else if (M >= 5) {
    draw_distinct<5>(seed, pair, s, M, idx);
  } else {
    for (int k = 0; k < 5; ++k) idx[k] = 0;
  }
  if (samples_out)
    for (int k = 0; k < 5; ++k) samples_out[((size_t)pair * S + s) * 5 + k] = idx[k];
  bool ok = M >= 5;
  for (int k = 0; k < 5; ++k) ok &= (idx[k] >= 0 && idx[k] < M);
  for (int k = 0; k < kFpMaxSol * 9; ++k) eo[k] = 0.0;
  if (n_sol) n_sol[(size_t)pair * S + s] = 0;
  if (!ok) return;

  // ---- 1. null space of the 5x9 system: Householder QR of Q^T, last four columns of the orthogonal factor ----
  double Q[5][9];
  for (int k = 0; k < 5; ++k) {
    const float4 c = cp[idx[k]];
    const double x = c.x, y = c.y, u = c.z, v = c.w;
    Q[k][0] = u * x; Q[k][1] = u * y; Q[k][2] = u;
    Q[k][3] = v * x; Q[k][4] = v * y; Q[k][5] = v;
    Q[k][6] = x;     Q[k][7] = y;     Q[k][8] = 1.0;
  }
  double v0[5], beta[5];
  for (int k = 0; k < 5; ++k) {
    double ss = 0.0;
    for (int i = k; i < 9; ++i) ss = fma(Q[k][i], Q[k][i], ss);
    const double nrm = sqrt(ss), x0 = Q[k][k];
    v0[k] = (ss > 0.0) ? x0 + copysign(nrm, x0) : 0.0;
    beta[k] = (ss > 0.0) ? 1.0 / fma(fabs(x0), nrm, ss) : 0.0;
    for (int j = k + 1; j < 5; ++j) {
      double d = v0[k] * Q[j][k];
      for (int i = k + 1; i < 9; ++i) d = fma(Q[k][i], Q[j][i], d);
      d *= beta[k];
      Q[j][k] = fma(-d, v0[k], Q[j][k]);
      for (int i = k + 1; i < 9; ++i) Q[j][i] = fma(-d, Q[k][i], Q[j][i]);
    }
  }
  double Eb[4][9];  // E = x Eb[0] + y Eb[1] + z Eb[2] + Eb[3]
  for (int b = 0; b < 4; ++b) {
    double n[9];
    for (int i = 0; i < 9; ++i) n[i] = (i == 5 + b) ? 1.0 : 0.0;
    for (int k = 4; k >= 0; --k) {
      double d = v0[k] * n[k];
      for (int i = k + 1; i < 9; ++i) d = fma(Q[k][i], n[i], d);
      d *= beta[k];
      n[k] = fma(-d, v0[k], n[k]);
      for (int i = k + 1; i < 9; ++i) n[i] = fma(-d, Q[k][i], n[i]);
    }
    for (int i = 0; i < 9; ++i) Eb[b][i] = n[i];
  }

  // ---- 2. constraint matrix ----
  // T = E E^T (symmetric, quadratic polynomials), L = T - 1/2 tr(T) I, rows 0..8 = L E, row 9 = det E
  double T[3][3][10];
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      double t[10];
      for (int q = 0; q < 10; ++q) t[q] = 0.0;
      for (int k = 0; k < 3; ++k)
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) t[kQuadIdx[a][b]] = fma(Eb[a][3 * i + k], Eb[b][3 * j + k], t[kQuadIdx[a][b]]);
      for (int q = 0; q < 10; ++q) T[i][j][q] = T[j][i][q] = t[q];
    }
  for (int q = 0; q < 10; ++q) {
    const double h = 0.5 * (T[0][0][q] + T[1][1][q] + T[2][2][q]);
    T[0][0][q] -= h;
    T[1][1][q] -= h;
    T[2][2][q] -= h;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double row[20];
      for (int c = 0; c < 20; ++c) row[c] = 0.0;
      for (int k = 0; k < 3; ++k)
        for (int q = 0; q < 10; ++q)
          for (int l = 0; l < 4; ++l) row[kCubIdx[q][l]] = fma(T[i][k][q], Eb[l][3 * k + j], row[kCubIdx[q][l]]);
      for (int c = 0; c < 20; ++c) A_(3 * i + j, c) = row[c];
    }
  {
    // det E = E00 (E11 E22 - E12 E21) - E01 (E10 E22 - E12 E20) + E02 (E10 E21 - E11 E20)
    double row[20];
    for (int c = 0; c < 20; ++c) row[c] = 0.0;
    const int m1[3][4] = {{4, 8, 5, 7}, {3, 8, 5, 6}, {3, 7, 4, 6}};  // minors of row 0: (a d - b c) with entries a, d, b, c
    for (int col = 0; col < 3; ++col) {
      double mnr[10];
      for (int q = 0; q < 10; ++q) mnr[q] = 0.0;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
          const int q = kQuadIdx[a][b];
          mnr[q] = fma(Eb[a][m1[col][0]], Eb[b][m1[col][1]], mnr[q]);
          mnr[q] = fma(-Eb[a][m1[col][2]], Eb[b][m1[col][3]], mnr[q]);
        }
      const double sgn = (col == 1) ? -1.0 : 1.0;
      for (int q = 0; q < 10; ++q)
        for (int l = 0; l < 4; ++l) row[kCubIdx[q][l]] = fma(sgn * mnr[q], Eb[l][col], row[kCubIdx[q][l]]);
    }
    for (int c = 0; c < 20; ++c) A_(9, c) = row[c];
  }

  // ---- Gauss-Jordan with partial pivoting on the first ten columns ----
  bool singular = false;
  for (int k = 0; k < 10; ++k) {
    int piv = k;
    double big = fabs(A_(k, k));
    for (int r = k + 1; r < 10; ++r) {
      const double a = fabs(A_(r, k));
      if (a > big) big = a, piv = r;
    }
    if (!(big > 0.0)) {
      singular = true;
      break;
    }
    if (piv != k)
      for (int c = k; c < 20; ++c) {
        const double t = A_(k, c);
        A_(k, c) = A_(piv, c);
        A_(piv, c) = t;
      }
    const double inv = 1.0 / A_(k, k);
    for (int c = k; c < 20; ++c) A_(k, c) *= inv;
    for (int r = 0; r < 10; ++r) {
      if (r == k) continue;
      const double f = A_(r, k);
      if (f != 0.0)
        for (int c = k; c < 20; ++c) A_(r, c) = fma(-f, A_(k, c), A_(r, c));
    }
  }
  if (singular) return;

  // ---- 3. B(z): rows <4> - z <5>, <6> - z <7>, <8> - z <9>; polynomials lowest power first ----
  double Bx[3][4], By[3][4], Bc[3][5];
  for (int r = 0; r < 3; ++r) {
    double a[10], b[10];
    for (int c = 0; c < 10; ++c) {
      a[c] = A_(4 + 2 * r, 10 + c);
      b[c] = A_(5 + 2 * r, 10 + c);
    }
    Bx[r][3] = -b[0]; Bx[r][2] = a[0] - b[1]; Bx[r][1] = a[1] - b[2]; Bx[r][0] = a[2];
    By[r][3] = -b[3]; By[r][2] = a[3] - b[4]; By[r][1] = a[4] - b[5]; By[r][0] = a[5];
    Bc[r][4] = -b[6]; Bc[r][3] = a[6] - b[7]; Bc[r][2] = a[7] - b[8]; Bc[r][1] = a[8] - b[9]; Bc[r][0] = a[9];
  }
  // det B = sum_r  Bc[r] * cof_r,  cof_r = (Bx By' - By Bx') of the other two rows (degree 6), cyclic signs
  double p[11];
  for (int k = 0; k <= 10; ++k) p[k] = 0.0;
  for (int r = 0; r < 3; ++r) {
    const int r1 = (r + 1) % 3, r2 = (r + 2) % 3;
    double cof[7];
    for (int k = 0; k < 7; ++k) cof[k] = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) cof[i + j] += Bx[r1][i] * By[r2][j] - By[r1][i] * Bx[r2][j];
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 7; ++j) p[i + j] = fma(Bc[r][i], cof[j], p[i + j]);
  }

  // ---- 4. real roots ----
  double roots[kFpMaxSol];
  const int nr = real_roots_sturm(p, 10, roots);

  // ---- 5. x, y from the null vector of B(z); E = x E0 + y E1 + z E2 + E3 ----
  int n_out = 0;
  for (int k = 0; k < nr; ++k) {
    const double z = roots[k];
    double Bz[3][3];
    for (int r = 0; r < 3; ++r) {
      Bz[r][0] = horner(Bx[r], 3, z);
      Bz[r][1] = horner(By[r], 3, z);
      Bz[r][2] = horner(Bc[r], 4, z);
    }
    double best[3] = {0.0, 0.0, 0.0};
    for (int r = 0; r < 3; ++r) {
      const int r1 = (r + 1) % 3, r2 = (r + 2) % 3;
      const double c0 = Bz[r1][1] * Bz[r2][2] - Bz[r1][2] * Bz[r2][1];
      const double c1 = Bz[r1][2] * Bz[r2][0] - Bz[r1][0] * Bz[r2][2];
      const double c2 = Bz[r1][0] * Bz[r2][1] - Bz[r1][1] * Bz[r2][0];
      if (fabs(c2) > fabs(best[2])) best[0] = c0, best[1] = c1, best[2] = c2;
    }
    if (best[2] == 0.0) continue;
    const double x = best[0] / best[2], y = best[1] / best[2];
    double e[9], n2 = 0.0;
    for (int i = 0; i < 9; ++i) {
      e[i] = fma(x, Eb[0][i], fma(y, Eb[1][i], fma(z, Eb[2][i], Eb[3][i])));
      n2 = fma(e[i], e[i], n2);
    }
    if (!(n2 > 0.0) || !isfinite(n2)) continue;
    const double sc = rsqrt(n2);
    for (int i = 0; i < 9; ++i) eo[n_out * 9 + i] = e[i] * sc;
    ++n_out;
  }
  if (n_sol) n_sol[(size_t)pair * S + s] = n_out;
#undef A_
}

}  // namespace b2s

extern "C" {

int b2s_five_point_batched(const float* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs, int S,
                           const int32_t* samples_in, uint64_t seed, int32_t* samples_out, double* E_out, int32_t* n_sol,
                           void* stream) {
  using namespace b2s;
  B2S_REQUIRE(corr && c_off && c_count && E_out, "null pointer");
  B2S_REQUIRE(n_pairs >= 0 && S >= 0, "negative size");
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  if (n_pairs == 0 || S == 0) return B2S_OK;
  const size_t smem = (size_t)10 * 20 * kFpThreads * sizeof(double);
  static bool attr_set[64] = {false};
  if (first_use_on_device(attr_set))
    B2S_CUDA(cudaFuncSetAttribute(five_point_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((S + kFpThreads - 1) / kFpThreads, n_pairs);
  five_point_kernel<<<grid, kFpThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(corr), c_off, c_count, S, samples_in, seed, samples_out, E_out, n_sol);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

}  // extern "C"
