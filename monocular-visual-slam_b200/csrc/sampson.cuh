// sampson.cuh — the Sampson inlier test of the reference (homography.py:328-333), shared by
// every scoring kernel and by the final inlier mask so that they agree bit for bit.
#pragma once

#include "common.cuh"

namespace b2s {

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

}  // namespace b2s
