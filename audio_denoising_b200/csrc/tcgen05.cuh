// tcgen05.cuh -- PTX wrappers for the 5th-generation tensor cores (sm_100a): TMEM allocation, tcgen05.mma (kind::tf32) with
// shared-memory operand descriptors in the canonical K-major no-swizzle core-matrix layout, tcgen05.commit -> mbarrier,
// tcgen05.ld, and the fences the proxies need.  Shared by the fused U-Net kernels (unet_tc.cu).
#pragma once

#include <stdint.h>
#include <string.h>

#include "tma.cuh"

namespace b2d {
namespace tc5 {

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp; ncols a power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_addr(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, TF32 operands (K = 8 per instruction), fp32 accumulate; one thread issues
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with the accumulate flag as an immediate predicate (no setp per instruction in the single issuing thread)
template <bool ACC>
__device__ __forceinline__ void mma_tf32_imm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  if (ACC) {
    asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc),
                 "l"(b_desc), "r"(idesc)
                 : "memory");
  } else {
    asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc),
                 "l"(b_desc), "r"(idesc)
                 : "memory");
  }
}
// the mbarrier gets one arrival when every tcgen05.mma this thread issued so far has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core / TMA (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32 (w % 4) ...): issue only; wait with ld_wait()
__device__ __forceinline__ void ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the loaded registers may be used after this (one wait covers every tcgen05.ld issued before it)
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// canonical K-major no-swizzle layout (TF32: 4 elements per 16-byte core-matrix row):
// element (row, k) of a [rows, KP] operand sits at ((row / 8) * (KP / 4) + k / 4) * 32 + (row % 8) * 4 + k % 4   [floats]
__host__ __device__ inline int canon_off(int row, int k, int KP) { return (((row >> 3) * (KP >> 2) + (k >> 2)) << 5) + ((row & 7) << 2) + (k & 3); }
// descriptor of an operand starting at shared address saddr: LBO = 128 B (next k-chunk), SBO = (KP / 4) * 128 B (next 8 rows)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int KP) {
  const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((KP >> 2) * 128) >> 4;
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46);  // version 1, SWIZZLE_NONE
}
// instruction descriptor: D fp32, A / B TF32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
// TF32 split of an fp32 value: big = the 19 leading bits, small = the (exactly representable) rest, itself cut to TF32
__host__ __device__ inline float tf32_big(float v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
#else
  uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r;
#endif
}

}  // namespace tc5
}  // namespace b2d
