// Inline PTX helpers (mbarrier, 1-D TMA bulk copies, L2 policies) and the fused SpMV epilogue.
#pragma once
#include "kb_common.cuh"

// ---------------------------------------------------------------- epilogue --
__device__ __forceinline__ double kb_spmv_epilogue(double t, int mode, const double* z,
                                                   double coef, size_t idx) {
  if (mode == 1) return kb_mul_sub(coef, z[idx], t);  // t - coef*z
  if (mode == 2) return __dsub_rn(z[idx], t);         // z - t
  return t;
}

// ------------------------------------------------------------- PTX helpers --
__device__ __forceinline__ uint32_t kb_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void kb_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void kb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kb_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void kb_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void kb_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "KB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra KB_DONE_%=;\n"
      "bra KB_WAIT_%=;\n"
      "KB_DONE_%=:\n"
      "}\n" ::"r"(kb_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void kb_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(kb_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(kb_smem_u32(bar))
      : "memory");
}

// same, with an L2 eviction-priority hint (createpolicy): the matrix stream is read
// once (evict_first) while x windows are re-read by later tiles (evict_last)
__device__ __forceinline__ void kb_bulk_g2s_hint(void* dst, const void* src, uint32_t bytes,
                                                 uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(kb_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(kb_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t kb_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t kb_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

