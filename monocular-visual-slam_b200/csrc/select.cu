// select.cu — match selection on packed keys: ratio LUT, cross-check, stable sort, truncate.
//
// Replaces the Python post-processing of ORBFeaturePipeline.match
// (/root/reference/feature_pipeline.py.bak:85-94: ratio test in float64, list.sort by
// distance — stable, so ties keep ascending queryIdx — and [:max_matches]), the emit rule
// of cv::BFMatcher(crossCheck=True) (keep i iff argmin_i' D[i', f(i)] == i, ascending i),
// the ratio + symmetry tests of match_orb_descriptors (/root/reference/homography.py:16-25)
// and the keypoint gather of matches_to_points (feature_pipeline.py.bak:104-111).
//
// One CTA per pair.  Survivors become sort keys (distance<<22 | queryIdx) — unique, so a
// plain bitonic sort in shared memory reproduces the reference's stable order exactly.
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
  int use_ratio, use_cross, sort_by_distance, max_matches;
  int n_sort;  // power of two >= max_nq
};

__global__ void __launch_bounds__(1024) select_matches_kernel(const SelectParams p, const RatioLut lut) {
  extern __shared__ uint32_t s_key[];
  __shared__ int s_count;
  const int pair = blockIdx.x;
  const int qo = p.q_off[pair];
  const int nq = p.q_off[pair + 1] - qo;
  const int to = p.t_off[pair];
  const int ob = p.out_stride > 0 ? pair * p.out_stride : qo;  // output base
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (tid == 0) s_count = 0;
  __syncthreads();

  // smallest power of two covering this pair (CTA-uniform)
  int n = 32;
  while (n < nq) n <<= 1;
  if (n > p.n_sort) {  // caller's max_nq was too small for this pair: flag it, never overrun smem
    if (tid == 0) p.out_count[pair] = -1;
    return;
  }

  int mine = 0;
  for (int i = tid; i < n; i += nthr) {
    uint32_t key = kNone;
    if (i < nq) {
      const uint32_t b = p.fwd_best[qo + i];
      bool keep = b < kInvalidRow;
      const uint32_t d1 = b >> kIdxBits, j = b & kIdxMask;
      if (keep && p.use_ratio) {
        const uint32_t s2 = p.fwd_second[qo + i];
        keep = (s2 < kInvalidRow) && (d1 < (uint32_t)lut.v[min(s2 >> kIdxBits, 256u)]);
      }
      if (keep && p.use_cross) keep = (p.bwd_best[to + j] & kIdxMask) == (uint32_t)i && p.bwd_best[to + j] < kInvalidRow;
      if (keep) {
        key = p.sort_by_distance ? ((d1 << kIdxBits) | (uint32_t)i) : (uint32_t)i;
        ++mine;
      }
    }
    s_key[i] = key;
  }
  if (mine) atomicAdd(&s_count, mine);
  __syncthreads();

  // bitonic sort, ascending; "none" keys sink to the end
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n; i += nthr) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = s_key[i], b = s_key[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            s_key[i] = b;
            s_key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }

  int count = s_count;
  if (p.max_matches > 0 && count > p.max_matches) count = p.max_matches;
  if (p.out_stride > 0 && count > p.out_stride) {  // caller's stride cannot hold this pair: flag, never overrun
    if (tid == 0) p.out_count[pair] = -1;
    return;
  }
  if (tid == 0) p.out_count[pair] = count;
  const int kq0 = p.q_src ? p.q_src[pair] : qo, kt0 = p.t_src ? p.t_src[pair] : to;
  for (int k = tid; k < count; k += nthr) {
    const uint32_t i = s_key[k] & kIdxMask;
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
                                  int32_t* out_count, void* stream) {
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
  p.use_ratio = use_ratio;
  p.use_cross = use_cross;
  p.sort_by_distance = sort_by_distance;
  p.max_matches = max_matches;
  int n = 32;
  while (n < max_nq) n <<= 1;
  p.n_sort = n;
  const size_t smem = sizeof(uint32_t) * (size_t)n;
  const int threads = n >= 2048 ? 1024 : (n / 2 < 32 ? 32 : n / 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024) {
    B2S_CUDA(cudaFuncSetAttribute(select_matches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  select_matches_kernel<<<n_pairs, threads, smem, st>>>(p, lut);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}
