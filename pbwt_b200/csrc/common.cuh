// common.cuh -- shared device helpers (sm_100a): mbarrier + 1-D TMA bulk copies, warp utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pbwtgpu.h"

namespace pbwtgpu {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
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
// order earlier generic-proxy accesses to shared memory before a later async-proxy (TMA) write
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }

// pack3 byte count for one run of length n (pbwtCore.c:240-252)
__host__ __device__ __forceinline__ int p3_run_bytes(int n) {
  int nb = n / 63488;
  n -= nb * 63488;
  nb += (n >= 2048);
  n &= 2047;
  nb += (n >= 64);
  n &= 63;
  nb += (n != 0);
  return nb;
}
// emit the bytes of one run; returns count (pbwtCore.c:240-252)
__device__ __forceinline__ int p3_run_emit(uint8_t *out, int v, int n) {
  const uint8_t top = (uint8_t)(v << 7);
  int nb = 0;
  while (n >= 63488) { out[nb++] = top | 0x7f; n -= 63488; }
  if (n >= 2048) { out[nb++] = top | 0x60 | (uint8_t)(n >> 11); n &= 2047; }
  if (n >= 64) { out[nb++] = top | 0x40 | (uint8_t)(n >> 6); n &= 63; }
  if (n) out[nb++] = top | (uint8_t)n;
  return nb;
}
// run length coded by one pack3 byte (pbwtCore.c:232-238)
__host__ __device__ __forceinline__ int p3_byte_len(uint8_t b) {
  const int code = b & 0x7f;
  return code < 64 ? code : code < 96 ? (code - 64) << 6 : (code - 96) << 11;
}

}  // namespace pbwtgpu
