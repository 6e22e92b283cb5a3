// Probe: what does tcgen05.ld.32x32b.xN.pack::16b return?  (build: nvcc -arch=sm_100a)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(uint32_t* out) {
  __shared__ uint32_t s_tm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tm;
  const uint32_t la = tm + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c = 0; c < 64; c += 4) {
    uint32_t v[4];
    for (int j = 0; j < 4; ++j) v[j] = ((0xA000u + (c + j)) << 16) | (uint32_t)(row << 8 | (c + j));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(la + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(la + 8));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 8; ++j) out[row * 8 + j] = r[j];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 8 * 4);
  k<<<1, 128>>>(d);
  uint32_t h[128 * 8];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("err=%d\n", (int)e);
  for (int row : {0, 1, 33, 127}) { printf("row %3d:", row); for (int j = 0; j < 8; ++j) printf(" %08x", h[row * 8 + j]); printf("\n"); }
}
