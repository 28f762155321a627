// chunk_kernel_helpers.cuh -- what is left of the round-1 trie-walk kernels' shared code: shared-
// memory layout constants, the TMA table staging, 4-byte lower-casing and the capcode character
// classes.  Used only by the chunk kernel for long documents (tokenize_thread_kernel.cuh,
// CHUNK = true); everything else runs the position-table pipeline (pipeline_kernels.cuh).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_tables.h"
#include "kernels.h"

namespace tmgpu {

constexpr int kSmemTableBytes = 2304;  // root_child[256] u32 + begin_byte[256] + sp_child[256] u32
constexpr int kSmemRingBase = 2560;    // tables + mbarrier + class table, 512-aligned ring area


// ---- small PTX helpers -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void stage_tables_tma(uint8_t* smem, const uint8_t* image) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSmemTableBytes);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(kSmemTableBytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem)),
        "l"(image), "r"(kSmemTableBytes), "r"(smem_u32(bar))
        : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar))
        : "memory");
  }
}

__device__ __forceinline__ uint32_t lower4(uint32_t w) {  // ASCII A-Z -> a-z in 4 packed bytes
  uint32_t hi = w & 0x80808080u;
  uint32_t x = w & 0x7F7F7F7Fu;
  uint32_t ge_a = x + 0x3F3F3F3Fu;   // bit7 set if byte >= 'A'
  uint32_t gt_z = x + 0x25252525u;   // bit7 set if byte >  'Z'
  uint32_t is_up = (ge_a & ~gt_z & ~hi) & 0x80808080u;
  return w | (is_up >> 2);
}

// character classes of the capcode encoder (javascript/tokenmonster.js:880-898, ASCII)
enum : uint32_t { CL_O = 0, CL_U = 1, CL_L = 2, CL_N = 3, CL_A = 4 };
__device__ __forceinline__ uint32_t char_class(uint32_t c) {
  if (c - 'a' < 26u) return CL_L;
  if (c - 'A' < 26u) return CL_U;
  if (c - '0' < 10u) return CL_N;
  if (c == '\'') return CL_A;
  return CL_O;
}

}  // namespace tmgpu
