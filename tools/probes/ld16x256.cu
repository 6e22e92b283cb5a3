// Probe: register <- (TMEM lane, column) map of tcgen05.ld.16x256b.xN[.pack::16b]  (build: nvcc -arch=sm_100a)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(uint32_t* out, uint32_t* out2) {
  __shared__ uint32_t s_tm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tm;
  const uint32_t la = tm + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c = 0; c < 128; c += 4) {
    uint32_t v[4];
    for (int j = 0; j < 4; ++j) v[j] = ((0xA000u + (c + j)) << 16) | (uint32_t)(row << 8 | (c + j));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(la + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // packed: x8 -> 32 registers
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.pack::16b.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                 "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                 "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(la));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = r[j];
  // second half of the warp's lanes (TMEM lanes +16)
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.pack::16b.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                 "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                 "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(la + (16u << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out2[(warp * 32 + lane) * 32 + j] = r[j];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tm) : "memory");
}
int main() {
  uint32_t *d, *d2; cudaMalloc(&d, 128 * 32 * 4); cudaMalloc(&d2, 128 * 32 * 4);
  k<<<1, 128>>>(d, d2);
  static uint32_t h[128 * 32], h2[128 * 32];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost);
  printf("err=%d (%s)\n", (int)e, cudaGetErrorString(e));
  // decode: low half = (row<<8|col) of one cell, high half = low 16 bits of the next cell?
  for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 37, 127}) {
    printf("thr %3d A:", t);
    for (int j = 0; j < 32; ++j) { uint32_t v = h[t * 32 + j]; printf(" [%d,%d|%d,%d]", (v & 0xFFFF) >> 8, v & 0xFF, (v >> 24) & 0xFF, (v >> 16) & 0xFF); }
    printf("\nthr %3d B:", t);
    for (int j = 0; j < 8; ++j) { uint32_t v = h2[t * 32 + j]; printf(" [%d,%d|%d,%d]", (v & 0xFFFF) >> 8, v & 0xFF, (v >> 24) & 0xFF, (v >> 16) & 0xFF); }
    printf("\n");
  }
}
