// tcgen05 / mbarrier / TF32-split helpers shared by the tensor-core kernels (ddk_conv_tc.cu, ddk_conv_tcr.cu).
#pragma once

#include <cstring>

#include "ddk_conv.cuh"

namespace ddk {

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// TF32 split a = hi + lo in three instructions: hi = a rounded to 10 mantissa bits (add half an ulp of the TF32 grid, clear the
// 13 low bits), lo = a - hi (exact in fp32).  lo is passed as it is: the tensor core reads the upper 19 bits of a TF32 operand,
// i.e. truncates lo to 10 mantissa bits -- an error of 2^-21 |a|, the size of the lo*lo term the 3-pass product drops anyway.
// (cvt.rna.tf32.f32 is emulated with ~6 instructions on sm_100a; the split was the bottleneck of the row warps.)
__host__ __device__ __forceinline__ void tc_split(float a, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  hi = (__float_as_uint(a) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
#else
  uint32_t b; memcpy(&b, &a, 4);
  hi = (b + 0x1000u) & 0xffffe000u;
  float h, l; memcpy(&h, &hi, 4);
  l = a - h; memcpy(&lo, &l, 4);
#endif
}
// the same split with lo ROUNDED to the TF32 grid (round half away, like hi) instead of leaving the truncation to the tensor core:
// the representation error of an operand drops from 2^-21 |a| (biased towards zero) to 2^-22 |a| (unbiased) for two more integer
// instructions per value.  hi + lo == a no longer holds exactly; |a - hi - lo| <= 2^-22 |a|.
__host__ __device__ __forceinline__ void tc_split_rn(float a, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  hi = (__float_as_uint(a) + 0x1000u) & 0xffffe000u;
  lo = (__float_as_uint(a - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
#else
  uint32_t b; memcpy(&b, &a, 4);
  hi = (b + 0x1000u) & 0xffffe000u;
  float h, l; memcpy(&h, &hi, 4);
  l = a - h; memcpy(&lo, &l, 4);
  lo = (lo + 0x1000u) & 0xffffe000u;
#endif
}
__device__ __forceinline__ void tc_mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(unsigned long long* b) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(tc_smem(b)) : "memory");
}
// Waiting warps back off between probes: in the first k_conv_tcr capture 42 % of all issued instructions were try_wait / branch
// pairs of warps spinning on these barriers, taken from the issue slots of the warps they were waiting for.
__device__ __forceinline__ void tc_mbar_wait(unsigned long long* b, uint32_t parity) {
  uint32_t ok = 0;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(tc_smem(b)), "r"(parity) : "memory");
  while (!ok) {
    __nanosleep(40);
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tc_smem(b)), "r"(parity) : "memory");
  }
}
// K-major, no-swizzle shared-memory descriptor: start address, K-direction (leading) and 8-row-group (stride) byte offsets in
// 16-byte units, descriptor version 1 (cute/arch/mma_sm100_desc.hpp; validated in tools/microbench/umma_tf32x3.cu)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(b)) : "memory");
}


__device__ __forceinline__ void tc_mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(tc_smem(b)), "r"(bytes) : "memory");
}
// bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP), completes on an mbarrier; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc_smem(dst)), "l"(src), "r"(bytes), "r"(tc_smem(bar)) : "memory");
}
// waits that back off between probes (warps that wait for a whole segment must not take issue slots from the producers)
__device__ __forceinline__ void tc_mbar_wait_sleep(unsigned long long* b, uint32_t parity) {
  uint32_t ok = 0;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(tc_smem(b)), "r"(parity) : "memory");
  while (!ok) {
    __nanosleep(200);
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tc_smem(b)), "r"(parity) : "memory");
  }
}

}  // namespace ddk
