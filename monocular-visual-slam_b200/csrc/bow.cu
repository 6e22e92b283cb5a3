// bow.cu — K9: bag-of-words candidate ranking.
//
// The step in front of the relocalizer's matching: compute_bow_histogram
// (/root/reference/persistent_map.py:82-96; the same arithmetic as BoWDatabase._compute_hist,
// /root/reference/loop_closure.py:36-48) assigns every ORB descriptor, read as 32 float values
// 0..255, to its nearest vocabulary centroid (sklearn pairwise_distances_argmin_min: squared
// Euclidean distance through ||x||^2 - 2 x.y + ||y||^2 in float64, first index on ties),
// counts the words and divides by the number of descriptors in float32; the ranking is
// sklearn's cosine_similarity of the query histogram against every map histogram
// (persistent_map.py:235, loop_closure.py:63).
//
// Both are small dense float problems (2000 x k x 32 per frame, k = 64..500), far from any
// roofline; what the device version buys is that the query descriptors are already resident
// for the Hamming kernels and that a whole map (4541 keyframes, BASELINE config #5) is
// histogrammed in one launch.  Arithmetic is float64 like sklearn's: with integer-valued x and
// float32 centroids every product is exact and the 32-term sums differ from sklearn's GEMM order
// by rounding at the 1e-13 level, so the argmin differs only on exact ties, which both resolve
// to the lowest index.
#include "common.cuh"

namespace b2s {

constexpr int kBowThreads = 128;
constexpr int kBowChunk = 64;   // centroids staged per pass: 64 x 32 doubles = 16 KB
constexpr int kBowDim = 32;

__global__ void __launch_bounds__(kBowThreads) bow_words_kernel(
    const uint8_t* __restrict__ desc, const int32_t* __restrict__ f_off, const float* __restrict__ vocab, int k,
    int32_t* __restrict__ words, int32_t* __restrict__ counts) {
  __shared__ double s_v[kBowChunk][kBowDim];
  __shared__ double s_n[kBowChunk];
  const int frame = blockIdx.y;
  const int o = f_off[frame], n = f_off[frame + 1] - o;
  if ((int)blockIdx.x * kBowThreads >= n) return;  // block-uniform
  const int i = blockIdx.x * kBowThreads + threadIdx.x;
  const bool live = i < n;
  double m2x[kBowDim];  // -2 x_j (exact)
  {
    const uint4* p = reinterpret_cast<const uint4*>(desc + (size_t)(o + (live ? i : 0)) * kBowDim);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < kBowDim; ++j) m2x[j] = -2.0 * (double)((w[j >> 2] >> (8 * (j & 3))) & 0xFFu);
  }
  double best = 1.0e300;
  int bi = 0;
  for (int c0 = 0; c0 < k; c0 += kBowChunk) {
    const int nc = min(kBowChunk, k - c0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nc * kBowDim; idx += kBowThreads)
      s_v[idx / kBowDim][idx % kBowDim] = (double)__ldg(vocab + (size_t)c0 * kBowDim + idx);
    __syncthreads();
    if (threadIdx.x < nc) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < kBowDim; ++j) s = fma(s_v[threadIdx.x][j], s_v[threadIdx.x][j], s);
      s_n[threadIdx.x] = s;
    }
    __syncthreads();
    for (int c = 0; c < nc; ++c) {
      double acc = s_n[c];  // ||y||^2 - 2 x.y  (+ ||x||^2, the same for every centroid)
#pragma unroll
      for (int j = 0; j < kBowDim; ++j) acc = fma(m2x[j], s_v[c][j], acc);
      if (acc < best) {  // strict: the lowest index wins a tie
        best = acc;
        bi = c0 + c;
      }
    }
  }
  if (live) {
    if (words) words[o + i] = bi;
    atomicAdd(&counts[(size_t)frame * k + bi], 1);
  }
}

// hist = counts.astype(float32) / float32(n)   (np.bincount(...).astype(np.float32); hist /= hist.sum():
// the float32 sum of integer counts < 2^24 is exact)
__global__ void __launch_bounds__(256) bow_normalise_kernel(const int32_t* __restrict__ counts,
                                                            const int32_t* __restrict__ f_off, int n_frames, int k,
                                                            float* __restrict__ hist) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (size_t)n_frames * k) return;
  const int frame = (int)(idx / k);
  const int n = f_off[frame + 1] - f_off[frame];
  hist[idx] = n > 0 ? __fdiv_rn((float)counts[idx], (float)n) : 0.0f;
}

// scores[r] = <q, h_r> / (||q|| ||h_r||), zero rows score 0 (sklearn normalize leaves them zero);
// one warp per map histogram, float64 accumulation, rounded once to float32
__global__ void __launch_bounds__(256) bow_cosine_kernel(const float* __restrict__ q, const float* __restrict__ hists,
                                                         int n, int k, float* __restrict__ scores) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  const float* h = hists + (size_t)r * k;
  double dot = 0.0, hh = 0.0, qq = 0.0;
  for (int j = lane; j < k; j += 32) {
    const double a = (double)__ldg(q + j), b = (double)__ldg(h + j);
    dot = fma(a, b, dot);
    hh = fma(b, b, hh);
    qq = fma(a, a, qq);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    dot += __shfl_xor_sync(0xFFFFFFFFu, dot, s);
    hh += __shfl_xor_sync(0xFFFFFFFFu, hh, s);
    qq += __shfl_xor_sync(0xFFFFFFFFu, qq, s);
  }
  if (lane == 0) scores[r] = (hh > 0.0 && qq > 0.0) ? (float)(dot / (sqrt(qq) * sqrt(hh))) : 0.0f;
}

}  // namespace b2s

extern "C" {

int b2s_bow_histogram_batched(const uint8_t* desc, const int32_t* f_off, int n_frames, int max_n, const float* vocab,
                              int k, int32_t* words, int32_t* counts, float* hist, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(n_frames >= 0 && max_n >= 0 && k > 0, "bad size");
  B2S_REQUIRE(n_frames <= 65535, "n_frames %d exceeds grid.y limit 65535; split the batch", n_frames);
  if (n_frames == 0) return B2S_OK;
  B2S_REQUIRE(f_off && vocab && counts && hist, "null pointer");
  B2S_REQUIRE(max_n == 0 || desc, "null descriptor pointer");
  B2S_REQUIRE(((uintptr_t)desc & 15u) == 0, "descriptor buffer must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B2S_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_frames * k * sizeof(int32_t), st));
  if (max_n > 0) {
    bow_words_kernel<<<dim3((max_n + kBowThreads - 1) / kBowThreads, n_frames), kBowThreads, 0, st>>>(desc, f_off, vocab, k,
                                                                                                      words, counts);
    B2S_CUDA(cudaGetLastError());
    note_launch();
  }
  const size_t total = (size_t)n_frames * k;
  bow_normalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(counts, f_off, n_frames, k, hist);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

int b2s_bow_cosine(const float* hist_q, const float* hists, int n, int k, float* scores, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(n >= 0 && k > 0, "bad size");
  if (n == 0) return B2S_OK;
  B2S_REQUIRE(hist_q && hists && scores, "null pointer");
  bow_cosine_kernel<<<(n + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(hist_q, hists, n, k, scores);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

}  // extern "C"
