// hamming_popc.cu — K1: brute-force Hamming kNN-2 + column minimum on the integer pipe.
//
// Replaces cv::BFMatcher(NORM_HAMMING) knnMatch(k=2) / crossCheck match
// (/root/reference/feature_pipeline.py.bak:68,82,84) and the NumPy distance loops of
// match_orb_descriptors (/root/reference/homography.py:12-15,21-23).
//
// Mapping.  One CTA = one (pair, query tile, train split).  A thread owns R query
// descriptors in registers (8 x u32 each) and walks the pair's train descriptors, which
// the TMA engine stages into shared memory in 4 KB chunks (128 descriptors) through a
// 4-deep mbarrier ring of cp.async.bulk copies.  A train descriptor is read by a whole
// warp with two broadcast LDS.128.  Per descriptor pair: 8 XOR, a carry-save tree that
// trades POPCs for LOP3s (CSA level 0..3 -> 8/6/5/4 POPC), one IMAD to form each packed
// key, a 3-instruction top-2 update per row, and a running minimum per column that is
// finished with one REDUX.MIN per warp and column.  Column minima of the warps meet in
// shared memory and leave the CTA as one atomicMin per column and chunk.
//
// Every reduction is a min over packed (distance<<22 | index) keys, so the result does
// not depend on arrival order: bit-exact against OpenCV's lowest-index tie rule.
#include "common.cuh"

namespace b2s {

struct HammingParams {
  const uint4* __restrict__ q;  // 2 x uint4 per descriptor
  const uint4* __restrict__ t;
  const int32_t* __restrict__ q_off;
  const int32_t* __restrict__ t_off;
  const int32_t* __restrict__ q_src;  // optional input row of each pair's first descriptor
  const int32_t* __restrict__ t_src;
  uint32_t* __restrict__ fwd_best;
  uint32_t* __restrict__ fwd_second;
  uint32_t* __restrict__ bwd_best;
  uint2* __restrict__ partial;  // [t_split][total_nq] when t_split > 1
  int total_nq;
  int t_split;
};

constexpr int kTC = 128;    // train descriptors per chunk (4 KB)
constexpr int kStages = 4;  // bulk-copy ring depth

// Explicit LOP3s: NVVM otherwise pushes the XORs through the carry-save tree and emits
// ~20 LOP3 per descriptor pair instead of 12-20 (measured: the 64-lane ALU pipe, not the
// 16-lane POPC pipe, became the binder).  asm keeps the tree as written.
__device__ __forceinline__ uint32_t xor2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, 0, 0x3c;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t and2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, 0, 0xc0;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void csa(uint32_t& s, uint32_t& c, uint32_t a, uint32_t b, uint32_t d) {
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(a), "r"(b), "r"(d));  // a ^ b ^ d
  asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(c) : "r"(a), "r"(b), "r"(d));  // majority
}

// 256-bit Hamming distance.  LEVEL selects how many POPCs are traded for LOP3s.
template <int LEVEL>
__device__ __forceinline__ uint32_t ham256(const uint4& qa, const uint4& qb, const uint4& ta, const uint4& tb) {
  const uint32_t x0 = xor2(qa.x, ta.x), x1 = xor2(qa.y, ta.y), x2 = xor2(qa.z, ta.z), x3 = xor2(qa.w, ta.w);
  const uint32_t x4 = xor2(qb.x, tb.x), x5 = xor2(qb.y, tb.y), x6 = xor2(qb.z, tb.z), x7 = xor2(qb.w, tb.w);
  if constexpr (LEVEL == 0) {
    return (__popc(x0) + __popc(x1) + __popc(x2)) + (__popc(x3) + __popc(x4) + __popc(x5)) +
           (__popc(x6) + __popc(x7));
  } else {
    uint32_t s1, c1, s2, c2;
    csa(s1, c1, x0, x1, x2);
    csa(s2, c2, x3, x4, x5);
    if constexpr (LEVEL == 1) {  // 6 POPC, 12 LOP3
      return (__popc(s1) + __popc(s2) + __popc(x6)) + __popc(x7) + 2u * (__popc(c1) + __popc(c2));
    } else {
      uint32_t s3, c3;
      csa(s3, c3, s1, s2, x6);
      if constexpr (LEVEL == 2) {  // 5 POPC, 14 LOP3
        return (__popc(s3) + __popc(x7)) + 2u * (__popc(c1) + __popc(c2) + __popc(c3));
      } else {  // 4 POPC, 18 LOP3
        const uint32_t s4 = xor2(s3, x7), c4 = and2(s3, x7);
        uint32_t s5, c5;
        csa(s5, c5, c1, c2, c3);
        return __popc(s4) + 2u * (__popc(s5) + __popc(c4)) + 4u * __popc(c5);
      }
    }
  }
}

template <int R, int W, int CSA>
__global__ void __launch_bounds__(W * 32) hamming_knn2_popc_kernel(const HammingParams p) {
  constexpr int NT = W * 32;
  constexpr int QT = NT * R;
  __shared__ __align__(128) uint4 s_t[kStages][kTC * 2];
  __shared__ __align__(8) uint64_t s_full[kStages];
  __shared__ uint32_t s_col[2][W][kTC];

  const int pair = blockIdx.y;
  const int qout = p.q_off[pair];
  const int nq = p.q_off[pair + 1] - qout;
  const int q0 = blockIdx.x * QT;
  if (q0 >= nq) return;  // CTA-uniform, before any barrier
  const int tout = p.t_off[pair];
  const int nt = p.t_off[pair + 1] - tout;
  const int qin = p.q_src ? p.q_src[pair] : qout;
  const int tin = p.t_src ? p.t_src[pair] : tout;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int z = blockIdx.z;
  const int nchunks = (nt + kTC - 1) / kTC;
  const int per = (nchunks + p.t_split - 1) / p.t_split;
  const int c_begin = min(z * per, nchunks);
  const int c_end = min(c_begin + per, nchunks);

  // this thread's query rows
  uint4 qa[R], qb[R];
  uint32_t iq[R], best[R], second[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = q0 + r * NT + tid;
    const bool valid = i < nq;
    const uint4* src = p.q + 2 * (size_t)(qin + (valid ? i : 0));
    qa[r] = __ldg(src);
    qb[r] = __ldg(src + 1);
    iq[r] = valid ? (uint32_t)i : kInvalidRow;
    best[r] = kNone;
    second[r] = kNone;
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint4* tsrc = p.t + 2 * (size_t)tin;
  auto issue = [&](int c) {  // thread 0 only
    const int s = (c - c_begin) % kStages;
    const int rows = min(kTC, nt - c * kTC);
    const uint32_t bytes = (uint32_t)rows * B2S_DESC_BYTES;
    mbar_arrive_expect_tx(&s_full[s], bytes);
    bulk_g2s(&s_t[s][0], tsrc + 2 * (size_t)c * kTC, bytes, &s_full[s]);
  };
  if (tid == 0) {
    for (int c = c_begin; c < min(c_begin + kStages, c_end); ++c) issue(c);
  }

  for (int c = c_begin; c < c_end; ++c) {
    const int it = c - c_begin;
    const int s = it % kStages;
    const int buf = it & 1;
    const int rows = min(kTC, nt - c * kTC);
    mbar_wait(&s_full[s], (uint32_t)(it / kStages) & 1u);

    const uint4* st = &s_t[s][0];
    const uint32_t jbase = (uint32_t)c * kTC;
#pragma unroll 2
    for (int j = 0; j < rows; ++j) {
      const uint4 ta = st[2 * j], tb = st[2 * j + 1];  // warp-broadcast LDS.128 x2
      const uint32_t jg = jbase + (uint32_t)j;
      uint32_t cmin = kNone;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t d = ham256<CSA>(qa[r], qb[r], ta, tb);
        top2_insert(best[r], second[r], d * (1u << kIdxBits) + jg);
        cmin = min(cmin, d * (1u << kIdxBits) + iq[r]);
      }
      cmin = __reduce_min_sync(0xffffffffu, cmin);
      if (lane == 0) s_col[buf][warp][j] = cmin;
    }
    __syncthreads();  // chunk consumed by every warp; column minima visible
    if (tid == 0 && c + kStages < c_end) issue(c + kStages);
    for (int j = tid; j < rows; j += NT) {
      uint32_t v = s_col[buf][0][j];
#pragma unroll
      for (int w = 1; w < W; ++w) v = min(v, s_col[buf][w][j]);
      atomicMin(&p.bwd_best[tout + (int)jbase + j], v);
    }
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (iq[r] != kInvalidRow) {
      const size_t o = (size_t)qout + iq[r];
      if (p.t_split == 1) {
        p.fwd_best[o] = best[r];
        p.fwd_second[o] = second[r];
      } else {
        p.partial[(size_t)z * p.total_nq + o] = make_uint2(best[r], second[r]);
      }
    }
  }
}

__global__ void hamming_merge_kernel(const uint2* __restrict__ partial, int total_nq, int t_split,
                                     uint32_t* __restrict__ fwd_best, uint32_t* __restrict__ fwd_second) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_nq) return;
  uint32_t best = kNone, second = kNone;
  for (int z = 0; z < t_split; ++z) {
    const uint2 v = partial[(size_t)z * total_nq + i];
    top2_insert(best, second, v.x);
    top2_insert(best, second, v.y);
  }
  fwd_best[i] = best;
  fwd_second[i] = second;
}

int hamming_merge_launch(const uint2* partial, int total_nq, int t_split, uint32_t* fwd_best, uint32_t* fwd_second,
                         cudaStream_t st) {   // also folds the train-axis split of the tensor-core kernel (hamming_i8.cu)
  if (total_nq <= 0) return B2S_OK;
  hamming_merge_kernel<<<(total_nq + 255) / 256, 256, 0, st>>>(partial, total_nq, t_split, fwd_best, fwd_second);
  B2S_CUDA(cudaGetLastError());
  note_launch();
  return B2S_OK;
}

// ---- host side -----------------------------------------------------------------
static int g_csa = 2, g_rows = 2, g_warps = 4;

template <int R, int W>
static cudaError_t launch_rw(int csa, dim3 grid, cudaStream_t st, const HammingParams& p) {
  switch (csa) {
    case 0: hamming_knn2_popc_kernel<R, W, 0><<<grid, W * 32, 0, st>>>(p); break;
    case 1: hamming_knn2_popc_kernel<R, W, 1><<<grid, W * 32, 0, st>>>(p); break;
    case 2: hamming_knn2_popc_kernel<R, W, 2><<<grid, W * 32, 0, st>>>(p); break;
    default: hamming_knn2_popc_kernel<R, W, 3><<<grid, W * 32, 0, st>>>(p); break;
  }
  return cudaGetLastError();
}

int hamming_popc_launch(const uint8_t* q, const uint8_t* t, const int32_t* q_off, const int32_t* t_off,
                        const int32_t* q_src, const int32_t* t_src, int n_pairs, int total_nq, int total_nt,
                        int max_nq, int max_nt, uint32_t* fwd_best, uint32_t* fwd_second, uint32_t* bwd_best,
                        int t_split, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int R = g_rows, W = g_warps, QT = R * W * 32;
  const int qtiles = (max_nq + QT - 1) / QT;
  const int nchunks = (max_nt + kTC - 1) / kTC;
  if (t_split <= 0) {  // fill the machine ~4 CTAs deep when the batch alone cannot
    const long ctas = (long)qtiles * n_pairs;
    const long target = 4L * sm_count();
    t_split = (int)((target + ctas - 1) / (ctas > 0 ? ctas : 1));
    if (workspace == nullptr) t_split = 1;
  }
  if (t_split > nchunks) t_split = nchunks;
  if (t_split < 1) t_split = 1;
  if (t_split > 65535) t_split = 65535;
  if (t_split > 1) {
    // shrink to what the caller's workspace can hold rather than fail
    while (t_split > 1 && b2s_hamming_workspace_bytes(total_nq, t_split) > workspace_bytes) --t_split;
  }
  B2S_REQUIRE(n_pairs <= 65535, "n_pairs %d exceeds grid.y limit 65535; split the batch", n_pairs);

  HammingParams p;
  p.q = reinterpret_cast<const uint4*>(q);
  p.t = reinterpret_cast<const uint4*>(t);
  p.q_off = q_off;
  p.t_off = t_off;
  p.q_src = q_src;
  p.t_src = t_src;
  p.fwd_best = fwd_best;
  p.fwd_second = fwd_second;
  p.bwd_best = bwd_best;
  p.partial = reinterpret_cast<uint2*>(workspace);
  p.total_nq = total_nq;
  p.t_split = t_split;

  if (total_nt > 0) B2S_CUDA(cudaMemsetAsync(bwd_best, 0xFF, sizeof(uint32_t) * (size_t)total_nt, st));
  if (total_nq == 0 || n_pairs == 0) return B2S_OK;
  if (max_nt == 0 || nchunks == 0) {  // no train rows anywhere: every neighbour is "none"
    B2S_CUDA(cudaMemsetAsync(fwd_best, 0xFF, sizeof(uint32_t) * (size_t)total_nq, st));
    B2S_CUDA(cudaMemsetAsync(fwd_second, 0xFF, sizeof(uint32_t) * (size_t)total_nq, st));
    return B2S_OK;
  }
  if (t_split > 1) {  // splits that own no chunk of a short pair still report "none"
    B2S_CUDA(cudaMemsetAsync(workspace, 0xFF, sizeof(uint2) * (size_t)total_nq * t_split, st));
  }
  dim3 grid(qtiles, n_pairs, t_split);
  cudaError_t e;
  if (R == 2 && W == 4) e = launch_rw<2, 4>(g_csa, grid, st, p);
  else if (R == 2 && W == 8) e = launch_rw<2, 8>(g_csa, grid, st, p);
  else if (R == 4 && W == 4) e = launch_rw<4, 4>(g_csa, grid, st, p);
  else e = launch_rw<4, 8>(g_csa, grid, st, p);
  B2S_CUDA(e);
  note_launch();
  if (t_split > 1) {
    hamming_merge_kernel<<<(total_nq + 255) / 256, 256, 0, st>>>(p.partial, total_nq, t_split, fwd_best,
                                                                fwd_second);
    B2S_CUDA(cudaGetLastError());
    note_launch();
  }
  return B2S_OK;
}

}  // namespace b2s

extern "C" {

size_t b2s_hamming_workspace_bytes(int total_nq, int t_split) {
  if (t_split <= 1 || total_nq <= 0) return 0;
  return sizeof(uint2) * (size_t)total_nq * (size_t)t_split;
}

int b2s_hamming_set_config(int csa_level, int rows_per_thread, int warps) {
  if (csa_level >= 0) {
    B2S_REQUIRE(csa_level <= 3, "csa_level must be 0..3");
    b2s::g_csa = csa_level;
  }
  if (rows_per_thread >= 0) {
    B2S_REQUIRE(rows_per_thread == 2 || rows_per_thread == 4, "rows_per_thread must be 2 or 4");
    b2s::g_rows = rows_per_thread;
  }
  if (warps >= 0) {
    B2S_REQUIRE(warps == 4 || warps == 8, "warps must be 4 or 8");
    b2s::g_warps = warps;
  }
  return B2S_OK;
}

int b2s_hamming_get_config(int* csa_level, int* rows_per_thread, int* warps) {
  if (csa_level) *csa_level = b2s::g_csa;
  if (rows_per_thread) *rows_per_thread = b2s::g_rows;
  if (warps) *warps = b2s::g_warps;
  return B2S_OK;
}

}  // extern "C"
