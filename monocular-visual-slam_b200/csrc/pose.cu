// pose.cu — K7: batched pose recovery from an essential matrix.
//
// Replaces decompose_essential (/root/reference/homography.py:251-299): the SVD of E, the four
// (R, t) candidates and — the expensive part, a Python double loop over 4 x M 4x4 SVDs in the
// reference — the cheirality vote: every inlier is triangulated by DLT against every candidate
// and counted if it lies in front of both cameras.  float64 like the reference.
//   * essential_candidates_kernel: one thread per pair.  Eigen-decomposition of E^T E (cyclic
//     Jacobi, 3x3) -> V, singular values; u_i = E v_i / s_i, u_3 = u_1 x u_2, v_3 = v_1 x v_2
//     (det U = det V = +1 by construction, the reference's sign fix-ups :262-265);
//     R_a = U W V^T, R_b = U W^T V^T, t = +-u_3 in the reference's candidate order (:268-273).
//   * cheirality_vote_kernel: one thread per correspondence, all four candidates.  The DLT
//     system A (4x4, :281-289) is reduced to S = A^T A; the right-singular vector of A's
//     smallest singular value is the dominant eigenvector of adj(S) (= sum_i prod_{j != i}
//     lambda_j v_i v_i^T), found by repeated squaring of the trace-normalised adjugate — no
//     division, square root or iteration-dependent branching per rotation as a Jacobi SVD would
//     need.  Only the signs of the two depths are used (:292-295).
// The candidate order is the reference's up to LAPACK's sign conventions for (u_i, v_i), which
// can only matter when two candidates tie on the vote.
#include "common.cuh"

namespace b2s {

struct PoseMat3 {
  double m[9];
};

// cyclic Jacobi on a symmetric 3x3: S -> diag, V = eigenvectors (columns)
__device__ void jacobi_eig3(double (&S)[3][3], double (&V)[3][3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) V[a][b] = (a == b) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
    if (off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const double apq = S[p][q];
      if (apq == 0.0) continue;
      const double theta = (S[q][q] - S[p][p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
      const double c = rsqrt(fma(t, t, 1.0)), s = t * c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double skp = S[k][p], skq = S[k][q];
        S[k][p] = c * skp - s * skq;
        S[k][q] = s * skp + c * skq;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double spk = S[p][k], sqk = S[q][k];
        S[p][k] = c * spk - s * sqk;
        S[q][k] = s * spk + c * sqk;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq;
        V[k][q] = s * vkp + c * vkq;
      }
    }
  }
}

// out: per pair 4 candidates x (R row-major 9, t 3) = 48 doubles
__global__ void __launch_bounds__(64) essential_candidates_kernel(const double* __restrict__ E, int n_pairs,
                                                                  double* __restrict__ cand) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_pairs) return;
  double e[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) e[k] = E[(size_t)pair * 9 + k];
  double S[3][3], V[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) S[a][b] = fma(e[a], e[b], fma(e[3 + a], e[3 + b], e[6 + a] * e[6 + b]));
  jacobi_eig3(S, V);
  // the two largest eigenvalues -> (v1, v2) in descending order (static selects: no local memory)
  const double l0 = S[0][0], l1 = S[1][1], l2 = S[2][2];
  const int i_min = (l0 <= l1 && l0 <= l2) ? 0 : ((l1 <= l2) ? 1 : 2);
  const int i_a = i_min == 0 ? 1 : 0, i_b = i_min == 2 ? 1 : 2;          // the other two, ascending index
  const double la = i_a == 0 ? l0 : l1, lb = i_b == 1 ? l1 : l2;
  const bool swap = lb > la;
  double v1[3], v2[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double ca = i_a == 0 ? V[k][0] : V[k][1], cb = i_b == 1 ? V[k][1] : V[k][2];
    v1[k] = swap ? cb : ca;
    v2[k] = swap ? ca : cb;
  }
  const double s1 = sqrt(fmax(swap ? lb : la, 0.0)), s2 = sqrt(fmax(swap ? la : lb, 0.0));
  double u1[3], u2[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    u1[r] = fma(e[3 * r], v1[0], fma(e[3 * r + 1], v1[1], e[3 * r + 2] * v1[2])) / s1;
    u2[r] = fma(e[3 * r], v2[0], fma(e[3 * r + 1], v2[1], e[3 * r + 2] * v2[2])) / s2;
  }
  // re-orthonormalise u2 against u1 (s1 ~ s2 for an essential matrix; E v_i are orthogonal up to rounding)
  const double d12 = fma(u1[0], u2[0], fma(u1[1], u2[1], u1[2] * u2[2]));
  double n2 = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    u2[r] = fma(-d12, u1[r], u2[r]);
    n2 = fma(u2[r], u2[r], n2);
  }
  const double in2 = rsqrt(n2);
#pragma unroll
  for (int r = 0; r < 3; ++r) u2[r] *= in2;
  const double u3[3] = {fma(u1[1], u2[2], -u1[2] * u2[1]), fma(u1[2], u2[0], -u1[0] * u2[2]), fma(u1[0], u2[1], -u1[1] * u2[0])};
  const double v3[3] = {fma(v1[1], v2[2], -v1[2] * v2[1]), fma(v1[2], v2[0], -v1[0] * v2[2]), fma(v1[0], v2[1], -v1[1] * v2[0])};
  // U W V^T = u2 v1^T - u1 v2^T + u3 v3^T ;  U W^T V^T = -u2 v1^T + u1 v2^T + u3 v3^T
  double* o = cand + (size_t)pair * 48;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double a = fma(u2[r], v1[c], -u1[r] * v2[c]), b = u3[r] * v3[c];
      o[0 * 12 + 3 * r + c] = o[1 * 12 + 3 * r + c] = a + b;
      o[2 * 12 + 3 * r + c] = o[3 * 12 + 3 * r + c] = b - a;
    }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[0 * 12 + 9 + r] = o[2 * 12 + 9 + r] = u3[r];
    o[1 * 12 + 9 + r] = o[3 * 12 + 9 + r] = -u3[r];
  }
}

// dominant eigenvector of the adjugate of the symmetric 4x4 S = smallest eigenvector of S
__device__ __forceinline__ void smallest_eigvec4(const double (&S)[4][4], double (&x)[4]) {
  // 2x2 minors of rows (0,1) and rows (2,3)
  const double s0 = S[0][0] * S[1][1] - S[1][0] * S[0][1], s1 = S[0][0] * S[1][2] - S[1][0] * S[0][2];
  const double s2 = S[0][0] * S[1][3] - S[1][0] * S[0][3], s3 = S[0][1] * S[1][2] - S[1][1] * S[0][2];
  const double s4 = S[0][1] * S[1][3] - S[1][1] * S[0][3], s5 = S[0][2] * S[1][3] - S[1][2] * S[0][3];
  const double c5 = S[2][2] * S[3][3] - S[3][2] * S[2][3], c4 = S[2][1] * S[3][3] - S[3][1] * S[2][3];
  const double c3 = S[2][1] * S[3][2] - S[3][1] * S[2][2], c2 = S[2][0] * S[3][3] - S[3][0] * S[2][3];
  const double c1 = S[2][0] * S[3][2] - S[3][0] * S[2][2], c0 = S[2][0] * S[3][1] - S[3][0] * S[2][1];
  // adjugate (symmetric because S is): upper triangle
  double m00 = S[1][1] * c5 - S[1][2] * c4 + S[1][3] * c3;
  double m01 = -S[0][1] * c5 + S[0][2] * c4 - S[0][3] * c3;
  double m02 = S[3][1] * s5 - S[3][2] * s4 + S[3][3] * s3;
  double m03 = -S[2][1] * s5 + S[2][2] * s4 - S[2][3] * s3;
  double m11 = S[0][0] * c5 - S[0][2] * c2 + S[0][3] * c1;
  double m12 = -S[3][0] * s5 + S[3][2] * s2 - S[3][3] * s1;
  double m13 = S[2][0] * s5 - S[2][2] * s2 + S[2][3] * s1;
  double m22 = S[3][0] * s4 - S[3][1] * s2 + S[3][3] * s0;
  double m23 = -S[2][0] * s4 + S[2][1] * s2 - S[2][3] * s0;
  double m33 = S[2][0] * s3 - S[2][1] * s1 + S[2][2] * s0;
  for (int it = 0; it < 40; ++it) {
    const double tr = m00 + m11 + m22 + m33;
    if (!(fabs(tr) > 0.0)) break;
    const double inv = 1.0 / tr;
    m00 *= inv; m01 *= inv; m02 *= inv; m03 *= inv; m11 *= inv; m12 *= inv; m13 *= inv; m22 *= inv; m23 *= inv; m33 *= inv;
    const double n00 = fma(m00, m00, fma(m01, m01, fma(m02, m02, m03 * m03)));
    const double n01 = fma(m00, m01, fma(m01, m11, fma(m02, m12, m03 * m13)));
    const double n02 = fma(m00, m02, fma(m01, m12, fma(m02, m22, m03 * m23)));
    const double n03 = fma(m00, m03, fma(m01, m13, fma(m02, m23, m03 * m33)));
    const double n11 = fma(m01, m01, fma(m11, m11, fma(m12, m12, m13 * m13)));
    const double n12 = fma(m01, m02, fma(m11, m12, fma(m12, m22, m13 * m23)));
    const double n13 = fma(m01, m03, fma(m11, m13, fma(m12, m23, m13 * m33)));
    const double n22 = fma(m02, m02, fma(m12, m12, fma(m22, m22, m23 * m23)));
    const double n23 = fma(m02, m03, fma(m12, m13, fma(m22, m23, m23 * m33)));
    const double n33 = fma(m03, m03, fma(m13, m13, fma(m23, m23, m33 * m33)));
    m00 = n00; m01 = n01; m02 = n02; m03 = n03; m11 = n11; m12 = n12; m13 = n13; m22 = n22; m23 = n23; m33 = n33;
    if (1.0 - (n00 + n11 + n22 + n33) < 2e-8) break;
  }
  // the column with the largest diagonal entry
  const double d01 = fmax(m00, m11), d23 = fmax(m22, m33);
  if (d01 >= d23) {
    if (m00 >= m11) { x[0] = m00; x[1] = m01; x[2] = m02; x[3] = m03; }
    else            { x[0] = m01; x[1] = m11; x[2] = m12; x[3] = m13; }
  } else {
    if (m22 >= m33) { x[0] = m02; x[1] = m12; x[2] = m22; x[3] = m23; }
    else            { x[0] = m03; x[1] = m13; x[2] = m23; x[3] = m33; }
  }
}

__global__ void __launch_bounds__(128) cheirality_vote_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ c_off,
                                                              const int32_t* __restrict__ c_count, const uint8_t* __restrict__ mask,
                                                              const double* __restrict__ cand, const PoseMat3 K,
                                                              int32_t* __restrict__ votes) {
  const int pair = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  const bool live = m < M && (mask == nullptr || mask[c_off[pair] + m] != 0);
  int in_front[4] = {0, 0, 0, 0};
  if (live) {
    const float4 c = corr[c_off[pair] + m];
    const double sx = c.x, sy = c.y, dx = c.z, dy = c.w;
    const double* k = K.m;
    // rows of P1 = K [I | 0]:  x P1[2] - P1[0],  y P1[2] - P1[1]
    const double a0[4] = {fma(sx, k[6], -k[0]), fma(sx, k[7], -k[1]), fma(sx, k[8], -k[2]), 0.0};
    const double a1[4] = {fma(sy, k[6], -k[3]), fma(sy, k[7], -k[4]), fma(sy, k[8], -k[5]), 0.0};
#pragma unroll 1
    for (int cd = 0; cd < 4; ++cd) {
      const double* rt = cand + (size_t)pair * 48 + cd * 12;
      double R[9], t[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = rt[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = rt[9 + i];
      // P2 = K [R | t]
      double P2[3][4];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) P2[r][cc] = fma(k[3 * r], R[cc], fma(k[3 * r + 1], R[3 + cc], k[3 * r + 2] * R[6 + cc]));
        P2[r][3] = fma(k[3 * r], t[0], fma(k[3 * r + 1], t[1], k[3 * r + 2] * t[2]));
      }
      double a2[4], a3[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        a2[cc] = fma(dx, P2[2][cc], -P2[0][cc]);
        a3[cc] = fma(dy, P2[2][cc], -P2[1][cc]);
      }
      double S[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[i][j] = fma(a0[i], a0[j], fma(a1[i], a1[j], fma(a2[i], a2[j], a3[i] * a3[j])));
      double X[4];
      smallest_eigvec4(S, X);
      const double iw = 1.0 / X[3];
      const double px = X[0] * iw, py = X[1] * iw, pz = X[2] * iw;
      const double z2 = fma(R[6], px, fma(R[7], py, fma(R[8], pz, t[2])));
      in_front[cd] = (pz > 0.0 && z2 > 0.0) ? 1 : 0;
    }
  }
#pragma unroll
  for (int cd = 0; cd < 4; ++cd) {
    const int n = __popc(__ballot_sync(0xFFFFFFFFu, in_front[cd] != 0));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&votes[pair * 4 + cd], n);
  }
}

// first maximum of the four votes wins (homography.py:296-298)
__global__ void __launch_bounds__(64) pose_pick_kernel(const double* __restrict__ cand, const int32_t* __restrict__ votes,
                                                       int n_pairs, double* __restrict__ R_out, double* __restrict__ t_out) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_pairs) return;
  int win = 0;
  for (int c = 1; c < 4; ++c) win = (votes[pair * 4 + c] > votes[pair * 4 + win]) ? c : win;
  const double* o = cand + (size_t)pair * 48 + win * 12;
  for (int k = 0; k < 9; ++k) R_out[(size_t)pair * 9 + k] = o[k];
  for (int k = 0; k < 3; ++k) t_out[(size_t)pair * 3 + k] = o[9 + k];
}

}  // namespace b2s

extern "C" {

int b2s_pose_pick(const double* candidates, const int32_t* votes, int n_pairs, double* R_out, double* t_out, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(candidates && votes && R_out && t_out, "null pointer");
  if (n_pairs <= 0) return B2S_OK;
  pose_pick_kernel<<<(n_pairs + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(candidates, votes, n_pairs, R_out, t_out);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

int b2s_decompose_essential_batched(const double* E, const float* corr, const int32_t* c_off, const int32_t* c_count,
                                    const uint8_t* inlier_mask, int n_pairs, int max_m, const double* K_host,
                                    double* candidates, int32_t* votes, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(E && corr && c_off && c_count && candidates && votes, "null pointer");
  B2S_REQUIRE(n_pairs >= 0 && max_m >= 0, "negative size");
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  if (n_pairs == 0) return B2S_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PoseMat3 K;
  for (int i = 0; i < 9; ++i) K.m[i] = K_host ? K_host[i] : ((i % 4 == 0) ? 1.0 : 0.0);
  B2S_CUDA(cudaMemsetAsync(votes, 0, sizeof(int32_t) * 4 * (size_t)n_pairs, st));
  essential_candidates_kernel<<<(n_pairs + 63) / 64, 64, 0, st>>>(E, n_pairs, candidates);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  if (max_m > 0) {
    cheirality_vote_kernel<<<dim3((max_m + 127) / 128, n_pairs), 128, 0, st>>>(reinterpret_cast<const float4*>(corr), c_off, c_count,
                                                                              inlier_mask, candidates, K, votes);
    B2S_CUDA(cudaGetLastError());
    note_launch();
  }
  return B2S_OK;
}

}  // extern "C"

// ---- n-point refit of E on the winner's inliers (homography.py:344 -> eight_point_E :222-248) ----
#include "linalg.cuh"

namespace b2s {

constexpr int kRefitThreads = 256;

// One CTA per pair.  The design matrix A (n x 9, one row per inlier) is never formed: its Gram
// matrix A^T A (9x9) is accumulated in a fixed order (per-thread partial sums over a strided
// subset, warp shuffles, then the 8 warps in warp order: deterministic), and the right-singular
// vector of A's smallest singular value is the eigenvector of the smallest eigenvalue of A^T A,
// found by a warp-cooperative cyclic Jacobi in shared memory.  Rank-2 projection and K^T F K as
// in the minimal solver.
__global__ void __launch_bounds__(kRefitThreads) refit_essential_kernel(
    const float4* __restrict__ corr, const int32_t* __restrict__ c_off, const int32_t* __restrict__ c_count,
    const uint8_t* __restrict__ mask, const PoseMat3 K, const PoseMat3 Kinv, double* __restrict__ E_out,
    int32_t* __restrict__ n_used) {
  __shared__ double s_part[8][45];
  __shared__ double s_M[9][9], s_V[9][9];
  __shared__ double s_cs[2];
  __shared__ int s_n[8];
  const int pair = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = max(c_count[pair], 0);   // a negative count is the selection kernel's overflow flag: no model
  const float4* cp = corr + c_off[pair];
  const uint8_t* mp = mask ? mask + c_off[pair] : nullptr;
  double acc[45];
#pragma unroll
  for (int k = 0; k < 45; ++k) acc[k] = 0.0;
  int n = 0;
  const double* ki = Kinv.m;
  for (int m = tid; m < M; m += kRefitThreads) {
    if (mp && mp[m] == 0) continue;
    const float4 c = cp[m];
    const double sx = c.x, sy = c.y, dx = c.z, dy = c.w;
    const double w1 = fma(ki[6], sx, fma(ki[7], sy, ki[8])), w2 = fma(ki[6], dx, fma(ki[7], dy, ki[8]));
    const double x = fma(ki[0], sx, fma(ki[1], sy, ki[2])) / w1, y = fma(ki[3], sx, fma(ki[4], sy, ki[5])) / w1;
    const double u = fma(ki[0], dx, fma(ki[1], dy, ki[2])) / w2, v = fma(ki[3], dx, fma(ki[4], dy, ki[5])) / w2;
    const double a[9] = {u * x, u * y, u, v * x, v * y, v, x, y, 1.0};
    int k = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
      for (int j = i; j < 9; ++j) acc[k] = fma(a[i], a[j], acc[k]), ++k;
    ++n;
  }
#pragma unroll
  for (int k = 0; k < 45; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xFFFFFFFFu, acc[k], o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 45; ++k) s_part[warp][k] = acc[k];
    s_n[warp] = n;
  }
  __syncthreads();
  if (tid < 45) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w][tid];
    // unpack the upper triangle index tid -> (i, j)
    int i = 0, rem = tid;
    while (rem >= 9 - i) rem -= 9 - i, ++i;
    const int j = i + rem;
    s_M[i][j] = t;
    s_M[j][i] = t;
  }
  if (tid < 81) s_V[tid / 9][tid % 9] = (tid / 9 == tid % 9) ? 1.0 : 0.0;
  __syncthreads();
  if (warp != 0) return;
  int total = 0;
  for (int w = 0; w < 8; ++w) total += s_n[w];
  if (lane == 0 && n_used) n_used[pair] = total;
  double* eo = E_out + (size_t)pair * 9;
  if (total < 8) {
    if (lane < 9) eo[lane] = 0.0;
    return;
  }
  // Jacobi on the 9x9 Gram matrix, warp-cooperative, in the round-robin PARALLEL ordering: round r of a sweep rotates
  // the four disjoint index pairs ((r + i) mod 9, (r - i) mod 9), i = 1..4 (nine rounds cover all 36 pairs), so a sweep
  // is 9 dependent steps instead of 36 (the sequential cyclic form took 134 us per 296 pairs, nearly all of it the
  // latency chain angle -> columns -> rows of one rotation at a time).  Lanes 0..3 compute the four angles; the
  // column / row updates are 36 independent (pair, index) tasks.  A rotation is skipped once
  // |a_pq| <= eps sqrt(a_pp a_qq) (the criterion that gives a positive semi-definite matrix its small eigen-pairs to
  // high relative accuracy); a sweep without rotations ends it.
  __shared__ int s_rot;
  __shared__ double s_c4[4], s_s4[4];
  __shared__ int s_p4[4], s_q4[4];
  for (int sweep = 0; sweep < 30; ++sweep) {
    if (lane == 0) s_rot = 0;
    __syncwarp();
    for (int r = 0; r < 9; ++r) {
      if (lane < 4) {
        const int a = (r + lane + 1) % 9, b2 = (r + 9 - (lane + 1)) % 9;
        const int p = min(a, b2), q = max(a, b2);
        const double apq = s_M[p][q], app = s_M[p][p], aqq = s_M[q][q];
        double c = 1.0, sn = 0.0;
        if (fabs(apq) > 1.0e-17 * sqrt(fabs(app * aqq)) && apq != 0.0) {
          const double theta = (aqq - app) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
          c = rsqrt(fma(t, t, 1.0));
          sn = t * c;
          s_rot = 1;
        }
        s_c4[lane] = c, s_s4[lane] = sn, s_p4[lane] = p, s_q4[lane] = q;
      }
      __syncwarp();
      for (int task = lane; task < 36; task += 32) {  // M <- M J (columns p, q of every pair), V <- V J
        const int j = task / 9, k = task - 9 * j, p = s_p4[j], q = s_q4[j];
        const double c = s_c4[j], sn = s_s4[j];
        if (sn != 0.0) {
          const double mkp = s_M[k][p], mkq = s_M[k][q];
          s_M[k][p] = c * mkp - sn * mkq;
          s_M[k][q] = sn * mkp + c * mkq;
          const double vkp = s_V[k][p], vkq = s_V[k][q];
          s_V[k][p] = c * vkp - sn * vkq;
          s_V[k][q] = sn * vkp + c * vkq;
        }
      }
      __syncwarp();
      for (int task = lane; task < 36; task += 32) {  // M <- J^T M (rows p, q of every pair)
        const int j = task / 9, k = task - 9 * j, p = s_p4[j], q = s_q4[j];
        const double c = s_c4[j], sn = s_s4[j];
        if (sn != 0.0) {
          const double mpk = s_M[p][k], mqk = s_M[q][k];
          s_M[p][k] = c * mpk - sn * mqk;
          s_M[q][k] = sn * mpk + c * mqk;
        }
      }
      __syncwarp();
    }
    if (s_rot == 0) break;
  }
  if (lane == 0) {
    int best = 0;
    for (int k = 1; k < 9; ++k) best = (s_M[k][k] < s_M[best][best]) ? k : best;
    double f[9];
    for (int k = 0; k < 9; ++k) f[k] = s_V[k][best];
    double v3[3];
    smallest_right_singular3(f, v3);
    double Fp[9];
    for (int r = 0; r < 3; ++r) {
      const double fv = fma(f[3 * r], v3[0], fma(f[3 * r + 1], v3[1], f[3 * r + 2] * v3[2]));
      for (int c = 0; c < 3; ++c) Fp[3 * r + c] = fma(-fv, v3[c], f[3 * r + c]);
    }
    double T1[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) T1[3 * r + c] = fma(Fp[3 * r], K.m[c], fma(Fp[3 * r + 1], K.m[3 + c], Fp[3 * r + 2] * K.m[6 + c]));
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) eo[3 * r + c] = fma(K.m[r], T1[c], fma(K.m[3 + r], T1[3 + c], K.m[6 + r] * T1[6 + c]));
  }
}

}  // namespace b2s

extern "C" {

int b2s_refit_essential_batched(const float* corr, const int32_t* c_off, const int32_t* c_count, const uint8_t* inlier_mask,
                                int n_pairs, const double* K_host, const double* Kinv_host, double* E_out, int32_t* n_used,
                                void* stream) {
  using namespace b2s;
  B2S_REQUIRE(corr && c_off && c_count && E_out, "null pointer");
  B2S_REQUIRE(n_pairs >= 0, "negative size");
  B2S_REQUIRE((K_host == nullptr) == (Kinv_host == nullptr), "pass both K and Kinv or neither");
  if (n_pairs == 0) return B2S_OK;
  PoseMat3 K, Kinv;
  for (int i = 0; i < 9; ++i) {
    K.m[i] = K_host ? K_host[i] : ((i % 4 == 0) ? 1.0 : 0.0);
    Kinv.m[i] = Kinv_host ? Kinv_host[i] : ((i % 4 == 0) ? 1.0 : 0.0);
  }
  refit_essential_kernel<<<n_pairs, kRefitThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(corr), c_off, c_count, inlier_mask, K, Kinv, E_out, n_used);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

}  // extern "C"
