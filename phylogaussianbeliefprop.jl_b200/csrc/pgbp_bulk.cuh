// Bulk asynchronous copies (TMA engine, 1-D) and the mbarrier that counts their bytes: shared by the shared-memory
// message kernels (pgbp_coop.cuh) and the element pass of the shared-precision path (pgbp_shared.cu).
#pragma once
#ifndef PGBP_HOST_EMUL
namespace pgbp {
// ---- bulk asynchronous copies (TMA engine, 1-D): one instruction moves a whole slot row of a tile ------------------
// With the batch-innermost layout the TILE elements of a tile are contiguous in every slot row (8 TILE bytes, 16-byte
// aligned), so one cp.async.bulk per row, issued by ONE thread, replaces TILE per-lane 8-byte cp.async (address
// arithmetic, predicate and LDGSTS per lane).  Completion is counted in bytes on an mbarrier.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");  // visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
}  // namespace pgbp
#endif
