// tma_stage.cuh -- 1-D bulk async copies (TMA, cp.async.bulk) global -> shared with mbarrier
// completion, as hand-written PTX for sm_100a.  Used by the wavefront kernels to stage tiles of
// the subject sequence through shared memory.  SASS: UBLKCP / SYNCS.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsq {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders earlier generic-proxy accesses of shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Warp-uniform bounded wait: every lane of a converged warp polls the same barrier and the loop
// condition is a warp vote, so control flow stays uniform.  try_wait suspends in hardware between
// probes.  A copy that has not landed after 2^29 clocks (~0.3 s; a tile takes microseconds) is a
// protocol failure: the warp raises the context's device fault word and returns false (warp-uniform),
// whereupon the caller abandons its task; other warps stop fetching tasks, and the host turns the word
// into TSQ_ERR_CUDA and discards the results (tsq_api.cpp: check_device_fault) -- never a silently
// wrong score.
__device__ __forceinline__ bool mbar_wait_warp(uint64_t* bar, uint32_t parity, int* fault) {
  if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return true;
  const long long t0 = clock64();
  for (;;) {
    if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return true;
    const bool give_up = (clock64() - t0 > (1ll << 29)) || *reinterpret_cast<volatile int*>(fault) != 0;
    if (__any_sync(0xffffffffu, give_up)) break;
  }
  if ((threadIdx.x & 31) == 0) atomicOr(fault, 1);
  return false;
}

}  // namespace tsq
