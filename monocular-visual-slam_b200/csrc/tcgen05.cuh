// tcgen05.cuh — thin PTX wrappers shared by the tensor-core kernels of libb2s (sm_100a):
// bounded mbarrier waits, tcgen05 fences / commit / election, the no-swizzle K-major shared
// memory matrix descriptor, and TMEM loads.
#pragma once

#include "common.cuh"

namespace b2s {

constexpr int kTcTileRows = 128;                     // rows per operand tile (M = N = 128)
constexpr int kTcChunkBytes = kTcTileRows * 16;      // one 16-byte k-chunk of all 128 rows (= LBO)

// ---- tcgen05 wrappers ------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  // try_wait suspends in hardware for a bounded time; a broken pipeline traps instead of hanging
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// one elected lane of a converged warp (ptxas then knows the tcgen05 instructions below it are
// issued by a single thread and drops the per-instruction election loop it emits otherwise)
__device__ __forceinline__ bool elect_one() {
  uint32_t is_leader;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
  return is_leader != 0u;
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, LBO = 2048 B, SBO = 128 B, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)(kTcChunkBytes >> 4) << 16; // leading (K-direction) byte offset, bits [16,30)
  d |= (uint64_t)(128u >> 4) << 32;          // stride (8-row group) byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  return d;                                  // base offset 0, layout type 0 = SWIZZLE_NONE
}

#define TMEM_LD_X32(taddr, v)                                                                              \
  asm volatile(                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),    \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
      : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// Empty asm that "rewrites" the 32 registers: ordered after the volatile tcgen05.wait::ld, it
// keeps the compiler from hoisting arithmetic on freshly loaded registers above the wait.
#define TMEM_REGS_READY(v)                                                                                 \
  asm volatile(""                                                                                          \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),       \
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),   \
                 "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]),             \
                 "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]),             \
                 "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]))


}  // namespace b2s
