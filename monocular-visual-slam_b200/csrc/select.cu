// select.cu — match selection on packed keys: ratio LUT, cross-check, stable sort, truncate.
//
// Replaces the Python post-processing of ORBFeaturePipeline.match
// (/root/reference/feature_pipeline.py.bak:85-94: ratio test in float64, list.sort by
// distance — stable, so ties keep ascending queryIdx — and [:max_matches]), the emit rule
// of cv::BFMatcher(crossCheck=True) (keep i iff argmin_i' D[i', f(i)] == i, ascending i),
// the ratio + symmetry tests of match_orb_descriptors (/root/reference/homography.py:16-25)
// and the keypoint gather of matches_to_points (feature_pipeline.py.bak:104-111).
//
// One CTA per pair; the stable order of the reference's sort is reproduced by a counting sort
// (see the kernel).
#include "common.cuh"

namespace b2s {

struct RatioLut {
  uint16_t v[257];  // keep iff d1 < v[d2]
};

struct SelectParams {
  const uint32_t* __restrict__ fwd_best;
  const uint32_t* __restrict__ fwd_second;
  const uint32_t* __restrict__ bwd_best;
  const int32_t* __restrict__ q_off;
  const int32_t* __restrict__ t_off;
  const float2* __restrict__ kp_q;
  const float2* __restrict__ kp_t;
  const int32_t* __restrict__ q_src;  // optional row of each pair's first keypoint (shared frames)
  const int32_t* __restrict__ t_src;
  int out_stride;                     // > 0: pair p's outputs start at p*out_stride, else at q_off[p]
  int32_t* __restrict__ out_q;
  int32_t* __restrict__ out_t;
  int32_t* __restrict__ out_d;
  float4* __restrict__ out_corr;
  int32_t* __restrict__ out_count;
  int32_t* __restrict__ out_total;    // optional: survivors before truncation
  int use_ratio, use_cross, sort_by_distance, max_matches;
  int n_sort;  // power of two >= max_nq
};

// Distances are integers 0..256, so the reference's stable sort by distance is a COUNTING sort.  Every warp of the
// CTA owns a contiguous segment of the query rows: (1) it histograms the distances of its segment's survivors into
// its own row of a [warps][bins] table, (2) the table is turned into output cursors — bins ascending, within a bin
// the warps (= row segments) ascending — and (3) every warp places its segment 32 rows at a time: __match_any_sync
// groups the lanes of equal distance, the lowest lane of a group advances that distance's cursor, the others take
// the slots behind it in lane (= query index) order.  Ties therefore keep ascending queryIdx exactly like the
// reference's stable list.sort.  (Round 1 placed with ONE warp: 63 dependent iterations for 2000 rows, 313 for the
// 10 000 rows of BASELINE config #4 — 0.10 ms, half of a lone 10k x 10k pair.)  sort_by_distance = 0 is the same
// with a single bin (ascending query index).
constexpr int kSelBins = 257;
constexpr int kSelBinsPad = 264;   // row stride of the per-warp table

__global__ void __launch_bounds__(1024) select_matches_kernel(const SelectParams p, const RatioLut lut) {
  extern __shared__ uint16_t s_sel[];       // [n_sort] distance (0xFFFF = dropped) | [n_sort] query index by output position | int [warps][kSelBinsPad]
  __shared__ int s_tot[kSelBinsPad + 32];   // survivors per bin, then the bin's first output position
  __shared__ int s_count;
  uint16_t* s_d = s_sel;
  uint16_t* s_sorted = s_sel + p.n_sort;
  int* s_hist = reinterpret_cast<int*>(s_sel + 2 * p.n_sort);
  const int pair = blockIdx.x;
  const int qo = p.q_off[pair];
  const int nq = p.q_off[pair + 1] - qo;
  const int to = p.t_off[pair];
  const int ob = p.out_stride > 0 ? pair * p.out_stride : qo;  // output base
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  if (nq > p.n_sort) {  // caller's max_nq was too small for this pair: flag it, never overrun smem
    if (tid == 0) {
      p.out_count[pair] = -1;
      if (p.out_total) p.out_total[pair] = -1;
    }
    return;
  }
  for (int b = tid; b < nwarps * kSelBinsPad; b += nthr) s_hist[b] = 0;
  for (int b = tid; b < kSelBinsPad + 32; b += nthr) s_tot[b] = 0;
  __syncthreads();

  // ---- (1) survivors and the histogram of their distances, per warp segment ----
  const int seg = (((nq + nwarps - 1) / nwarps) + 31) & ~31;      // rows per warp, a multiple of 32
  const int r0 = warp * seg, r1 = min(nq, r0 + seg);
  int* my_hist = s_hist + warp * kSelBinsPad;
  for (int i = r0 + lane; i < r1; i += 32) {
    const uint32_t b = p.fwd_best[qo + i];
    bool keep = b < kInvalidRow;
    const uint32_t d1 = b >> kIdxBits, j = b & kIdxMask;
    if (keep && p.use_ratio) {
      const uint32_t s2 = p.fwd_second[qo + i];
      keep = (s2 < kInvalidRow) && (d1 < (uint32_t)lut.v[min(s2 >> kIdxBits, 256u)]);
    }
    if (keep && p.use_cross) {
      const uint32_t bw = p.bwd_best[to + j];
      keep = (bw & kIdxMask) == (uint32_t)i && bw < kInvalidRow;
    }
    const uint32_t bin = p.sort_by_distance ? min(d1, 256u) : 0u;
    s_d[i] = keep ? (uint16_t)bin : (uint16_t)0xFFFFu;
    if (keep) atomicAdd(&my_hist[bin], 1);
  }
  __syncthreads();

  // ---- (2) cursors: within a bin the warps in order, then the bins in order ----
  for (int b = tid; b < kSelBins; b += nthr) {
    int run = 0;
    for (int w = 0; w < nwarps; ++w) {
      const int c = s_hist[w * kSelBinsPad + b];
      s_hist[w * kSelBinsPad + b] = run;     // exclusive over the warps
      run += c;
    }
    s_tot[b] = run;
  }
  __syncthreads();
  if (tid < 32) {   // exclusive prefix sum over the 257 bins (9 per lane)
    int local[9], sum = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      local[k] = s_tot[lane * 9 + k];   // bins >= 257 are zero padding
      sum += local[k];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - sum;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      s_tot[lane * 9 + k] = run;        // now: first output position of this distance
      run += local[k];
    }
    if (lane == 31) s_count = incl;
  }
  __syncthreads();

  // ---- (3) stable placement, every warp its own segment and its own cursors ----
  for (int base = r0; base < r1; base += 32) {
    const int i = base + lane;
    const uint32_t dd = i < r1 ? (uint32_t)s_d[i] : 0xFFFFu;
    const bool keep = dd != 0xFFFFu;
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, keep);
    if (keep) {
      const uint32_t grp = __match_any_sync(active, dd);
      const int leader = __ffs((int)grp) - 1;
      int pos = 0;
      if (lane == leader) {
        pos = my_hist[dd];
        my_hist[dd] = pos + __popc(grp);
      }
      pos = __shfl_sync(grp, pos, leader);
      s_sorted[s_tot[dd] + pos + __popc(grp & ((1u << lane) - 1u))] = (uint16_t)i;
    }
    __syncwarp();
  }
  __syncthreads();

  int count = s_count;
  if (tid == 0 && p.out_total) p.out_total[pair] = count;
  if (p.max_matches > 0 && count > p.max_matches) count = p.max_matches;
  if (p.out_stride > 0 && count > p.out_stride) {  // caller's stride cannot hold this pair: flag, never overrun
    if (tid == 0) p.out_count[pair] = -1;
    return;
  }
  if (tid == 0) p.out_count[pair] = count;
  const int kq0 = p.q_src ? p.q_src[pair] : qo, kt0 = p.t_src ? p.t_src[pair] : to;
  for (int k = tid; k < count; k += nthr) {
    const uint32_t i = s_sorted[k];
    const uint32_t b = p.fwd_best[qo + i];
    const uint32_t j = b & kIdxMask;
    p.out_q[ob + k] = (int32_t)i;
    p.out_t[ob + k] = (int32_t)j;
    p.out_d[ob + k] = (int32_t)(b >> kIdxBits);
    if (p.out_corr) {
      const float2 a = p.kp_q[kq0 + i], c = p.kp_t[kt0 + j];
      p.out_corr[ob + k] = make_float4(a.x, a.y, c.x, c.y);
    }
  }
}

}  // namespace b2s

extern "C" int b2s_select_matches(const uint32_t* fwd_best, const uint32_t* fwd_second, const uint32_t* bwd_best,
                                  const int32_t* q_off, const int32_t* t_off, int n_pairs, int max_nq,
                                  int use_ratio, int use_cross, const int32_t* ratio_lut_host,
                                  int sort_by_distance, int max_matches, const float* kp_q, const float* kp_t,
                                  const int32_t* kp_q_src_row, const int32_t* kp_t_src_row, int out_stride,
                                  int32_t* out_q, int32_t* out_t, int32_t* out_d, float* out_corr,
                                  int32_t* out_count, int32_t* out_total, void* stream) {
  using namespace b2s;
  B2S_REQUIRE(n_pairs >= 0 && max_nq >= 0 && out_stride >= 0, "negative size");
  B2S_REQUIRE(max_nq <= B2S_SELECT_MAX_QUERIES, "select: %d queries per pair exceeds %d", max_nq,
              B2S_SELECT_MAX_QUERIES);
  B2S_REQUIRE(!use_ratio || ratio_lut_host != nullptr, "use_ratio needs ratio_lut_host");
  B2S_REQUIRE((out_corr == nullptr) || (kp_q && kp_t), "out_corr needs kp_q and kp_t");
  B2S_REQUIRE(fwd_best && q_off && t_off && out_q && out_t && out_d && out_count, "null pointer");
  B2S_REQUIRE(!use_ratio || fwd_second, "use_ratio needs fwd_second");
  B2S_REQUIRE(!use_cross || bwd_best, "use_cross needs bwd_best");
  if (n_pairs == 0) return B2S_OK;
  RatioLut lut;
  for (int d = 0; d <= 256; ++d) {
    int v = use_ratio ? ratio_lut_host[d] : 0;
    lut.v[d] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
  }
  SelectParams p;
  p.fwd_best = fwd_best;
  p.fwd_second = fwd_second;
  p.bwd_best = bwd_best;
  p.q_off = q_off;
  p.t_off = t_off;
  p.kp_q = reinterpret_cast<const float2*>(kp_q);
  p.kp_t = reinterpret_cast<const float2*>(kp_t);
  p.q_src = kp_q_src_row;
  p.t_src = kp_t_src_row;
  p.out_stride = out_stride;
  p.out_q = out_q;
  p.out_t = out_t;
  p.out_d = out_d;
  p.out_corr = reinterpret_cast<float4*>(out_corr);
  p.out_count = out_count;
  p.out_total = out_total;
  p.use_ratio = use_ratio;
  p.use_cross = use_cross;
  p.sort_by_distance = sort_by_distance;
  p.max_matches = max_matches;
  int n = 32;
  while (n < max_nq) n <<= 1;
  p.n_sort = n;   // capacity of the shared-memory arrays (rows per pair)
  const int threads = n >= 1024 ? 1024 : (n < 64 ? 64 : n);
  const size_t smem = 2 * sizeof(uint16_t) * (size_t)n + sizeof(int) * (size_t)(threads / 32) * kSelBinsPad;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024) {
    B2S_CUDA(cudaFuncSetAttribute(select_matches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  select_matches_kernel<<<n_pairs, threads, smem, st>>>(p, lut);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}
