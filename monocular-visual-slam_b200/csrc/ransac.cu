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
#include "common.cuh"

namespace b2s {

struct Mat3 {
  double m[9];
};

// ---- Sampson test (shared by scoring and the final mask so they agree bit for bit) ----
template <typename T>
__device__ __forceinline__ bool sampson_inlier(const T* e, T x, T y, T u, T v, T th2) {
  const T a0 = fma(e[0], x, fma(e[1], y, e[2]));  // (E x1)_0
  const T a1 = fma(e[3], x, fma(e[4], y, e[5]));  // (E x1)_1
  const T a2 = fma(e[6], x, fma(e[7], y, e[8]));  // (E x1)_2
  const T b0 = fma(e[0], u, fma(e[3], v, e[6]));  // (E^T x2)_0
  const T b1 = fma(e[1], u, fma(e[4], v, e[7]));  // (E^T x2)_1
  const T num = fma(u, a0, fma(v, a1, a2));       // x2^T E x1
  const T den = fma(a0, a0, fma(a1, a1, fma(b0, b0, b1 * b1)));
  return num * num < th2 * den;
}

// ---- K3: one thread = one hypothesis, correspondences broadcast from shared memory ----
constexpr int kScoreThreads = 128;
constexpr int kScoreChunk = 1024;

template <typename T>
__global__ void __launch_bounds__(kScoreThreads) ransac_score_kernel(
    const float4* __restrict__ corr, const int32_t* __restrict__ c_off, const int32_t* __restrict__ c_count,
    const double* __restrict__ E, int H, double th2_all, const double* __restrict__ th2_pp,
    int32_t* __restrict__ counts) {
  struct alignas(16) P4 { T x, y, u, v; };
  __shared__ P4 s_p[kScoreChunk];
  const int pair = blockIdx.y;
  const int h = blockIdx.x * kScoreThreads + threadIdx.x;
  const int M = c_count[pair];
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

// ---- winner selection + inlier mask ---------------------------------------------------
__global__ void __launch_bounds__(256) ransac_select_kernel(
    const int32_t* __restrict__ counts, const float4* __restrict__ corr, const int32_t* __restrict__ c_off,
    const int32_t* __restrict__ c_count, const double* __restrict__ E, int H, double th2_all,
    const double* __restrict__ th2_pp, int32_t* __restrict__ best_h, int32_t* __restrict__ best_count,
    uint8_t* __restrict__ mask) {
  __shared__ int s_early;
  __shared__ unsigned long long s_best;
  __shared__ int s_cnt;
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int M = c_count[pair];
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
    const int c = counts[(size_t)pair * H + h];
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
    const double* ep = E + ((size_t)pair * H + win) * 9;
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
  if (mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (tid == 0) {
    best_h[pair] = win;
    best_count[pair] = s_cnt;
  }
}

// ---- K4: 8-point minimal solver, one thread per hypothesis ---------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Null vector of an 8x9 matrix, entirely in registers with static indexing.
// Gaussian elimination with the pivot ROW fixed (row k at step k) and the pivot COLUMN chosen
// as the largest remaining entry of that row (= partial pivoting on A^T).  Nothing is ever
// swapped: used columns are a bit mask, a dynamically chosen column is read with a 9-way
// select chain, and the free (never chosen) column takes the value 1 in the back
// substitution.  Adaptive columns matter here: for forward motion E33 ~ 0, so fixing the
// last column as the free one would make the 8x8 system singular.
// History: matrix in local memory 947 us -> shared memory + complete pivoting 386 us
// (10.5k instructions per hypothesis, latency bound) -> this version.
constexpr int kEpThreads = 128;

__device__ __forceinline__ double pick9(const double (&a)[9], int c) {
  double v = a[0];
#pragma unroll
  for (int i = 1; i < 9; ++i) v = (c == i) ? a[i] : v;
  return v;
}

__device__ __forceinline__ void null_vector_8x9(double (&A)[8][9], double (&v)[9]) {
  unsigned used = 0u;
  int pcs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int pc = 0;
    double big = -1.0;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      const double a = ((used >> c) & 1u) ? -1.0 : fabs(A[k][c]);
      if (a > big) {
        big = a;
        pc = c;
      }
    }
    used |= 1u << pc;
    pcs[k] = pc;
    const double inv = (big > 0.0) ? 1.0 / pick9(A[k], pc) : 0.0;
#pragma unroll
    for (int r = k + 1; r < 8; ++r) {
      const double f = pick9(A[r], pc) * inv;
#pragma unroll
      for (int c = 0; c < 9; ++c) A[r][c] = fma(-f, A[k][c], A[r][c]);
    }
  }
  int fc = 0;
#pragma unroll
  for (int c = 0; c < 9; ++c) fc = ((used >> c) & 1u) ? fc : c;
  double x[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) x[c] = (c == fc) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    // x of this row's pivot column is still 0 and columns eliminated earlier multiply x = 0
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 9; ++c) acc = fma(A[k][c], x[c], acc);
    const double piv = pick9(A[k], pcs[k]);
    const double val = (piv != 0.0) ? -acc / piv : 0.0;
#pragma unroll
    for (int c = 0; c < 9; ++c) x[c] = (c == pcs[k]) ? val : x[c];
  }
  double n2 = 0.0;
#pragma unroll
  for (int c = 0; c < 9; ++c) n2 = fma(x[c], x[c], n2);
  const double sc = rsqrt(n2);
#pragma unroll
  for (int c = 0; c < 9; ++c) v[c] = x[c] * sc;
}

// Smallest-eigenvalue eigenvector of the symmetric 3x3 S (cyclic Jacobi).
__device__ void smallest_eigvec3(double S[3][3], double* out) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 8; ++sweep) {
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
      for (int k = 0; k < 3; ++k) {  // S <- S J
        const double skp = S[k][p], skq = S[k][q];
        S[k][p] = c * skp - s * skq;
        S[k][q] = s * skp + c * skq;
      }
      for (int k = 0; k < 3; ++k) {  // S <- J^T S
        const double spk = S[p][k], sqk = S[q][k];
        S[p][k] = c * spk - s * sqk;
        S[q][k] = s * spk + c * sqk;
      }
      for (int k = 0; k < 3; ++k) {
        const double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq;
        V[k][q] = s * vkp + c * vkq;
      }
    }
  }
  // static selects (a dynamically indexed V[k][mi] would push S and V to local memory)
  const bool use1 = S[1][1] < S[0][0];
  const double m01 = use1 ? S[1][1] : S[0][0];
  const bool use2 = S[2][2] < m01;
#pragma unroll
  for (int k = 0; k < 3; ++k) out[k] = use2 ? V[k][2] : (use1 ? V[k][1] : V[k][0]);
}

__global__ void __launch_bounds__(kEpThreads) eight_point_kernel(
    const float4* __restrict__ corr, const int32_t* __restrict__ c_off, const int32_t* __restrict__ c_count, int H,
    const int32_t* __restrict__ samples_in, uint64_t seed, int32_t* __restrict__ samples_out, const Mat3 K,
    const Mat3 Kinv, double* __restrict__ E_out) {
  const int pair = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const int M = c_count[pair];
  const float4* cp = corr + c_off[pair];
  double* eo = E_out + ((size_t)pair * H + h) * 9;
  int idx[8];
  if (samples_in) {
    for (int k = 0; k < 8; ++k) idx[k] = samples_in[((size_t)pair * H + h) * 8 + k];
  } else if (M >= 8) {
    uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(pair + 1)) ^ (0x8CB92BA72F3D8DD7ull * (uint64_t)(h + 1));
    for (int k = 0; k < 8; ++k) {
      int cand;
      bool dup;
      do {
        cand = (int)__umul64hi(splitmix64(s), (uint64_t)M);
        dup = false;
        for (int a = 0; a < k; ++a) dup |= (idx[a] == cand);
      } while (dup);
      idx[k] = cand;
    }
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
    const double w1 = fma(ki[6], sx, fma(ki[7], sy, ki[8]));
    const double x = fma(ki[0], sx, fma(ki[1], sy, ki[2])) / w1;
    const double y = fma(ki[3], sx, fma(ki[4], sy, ki[5])) / w1;
    const double w2 = fma(ki[6], dx, fma(ki[7], dy, ki[8]));
    const double u = fma(ki[0], dx, fma(ki[1], dy, ki[2])) / w2;
    const double v = fma(ki[3], dx, fma(ki[4], dy, ki[5])) / w2;
    A[k][0] = u * x; A[k][1] = u * y; A[k][2] = u;
    A[k][3] = v * x; A[k][4] = v * y; A[k][5] = v;
    A[k][6] = x;     A[k][7] = y;     A[k][8] = 1.0;
  }
  double f[9];
  null_vector_8x9(A, f);

  // rank-2 projection: F' = F - (F v3) v3^T, v3 = right-singular vector of the smallest
  // singular value (homography.py:244-246 zeroes S[2] only).
  double S[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) S[a][b] = fma(f[a], f[b], fma(f[3 + a], f[3 + b], f[6 + a] * f[6 + b]));
  double v3[3];
  smallest_eigvec3(S, v3);
  double Fp[9];
  for (int r = 0; r < 3; ++r) {
    const double fv = fma(f[3 * r], v3[0], fma(f[3 * r + 1], v3[1], f[3 * r + 2] * v3[2]));
    for (int c = 0; c < 3; ++c) Fp[3 * r + c] = fma(-fv, v3[c], f[3 * r + c]);
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

}  // namespace b2s

extern "C" {

int b2s_eight_point_batched(const float* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs, int H,
                            const int32_t* samples_in, uint64_t seed, int32_t* samples_out, const double* K_host,
                            const double* Kinv_host, double* E_out, void* stream) {
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
  eight_point_kernel<<<grid, kEpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(corr), c_off, c_count, H, samples_in, seed, samples_out, K, Kinv, E_out);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

int b2s_ransac_score_batched(const float* corr, const int32_t* c_off, const int32_t* c_count, int n_pairs,
                             const double* E, int H, double th2, const double* th2_per_pair, int precision,
                             int32_t* counts, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(corr && c_off && c_count && E && counts, "null pointer");
  B2S_REQUIRE(precision == 64 || precision == 32, "precision must be 64 or 32");
  B2S_REQUIRE(n_pairs >= 0 && H >= 0, "negative size");
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);
  if (n_pairs == 0 || H == 0) return B2S_OK;
  dim3 grid((H + kScoreThreads - 1) / kScoreThreads, n_pairs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float4* c4 = reinterpret_cast<const float4*>(corr);
  if (precision == 64)
    ransac_score_kernel<double><<<grid, kScoreThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts);
  else
    ransac_score_kernel<float><<<grid, kScoreThreads, 0, st>>>(c4, c_off, c_count, E, H, th2, th2_per_pair, counts);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

int b2s_ransac_select(const int32_t* counts, const float* corr, const int32_t* c_off, const int32_t* c_count,
                      int n_pairs, const double* E, int H, double th2, const double* th2_per_pair, int32_t* best_h,
                      int32_t* best_count, uint8_t* inlier_mask, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(counts && corr && c_off && c_count && E && best_h && best_count && inlier_mask, "null pointer");
  B2S_REQUIRE(n_pairs >= 0 && H >= 0, "negative size");
  if (n_pairs == 0) return B2S_OK;
  ransac_select_kernel<<<n_pairs, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      counts, reinterpret_cast<const float4*>(corr), c_off, c_count, E, H, th2, th2_per_pair, best_h, best_count,
      inlier_mask);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

}  // extern "C"
