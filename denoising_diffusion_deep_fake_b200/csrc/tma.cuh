// tma.cuh — host-side TMA tensor-map construction (driver entry point, cached) and device-side
// cp.async.bulk.tensor wrappers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace d3fk {

// Cached cuTensorMapEncodeTiled for bf16 tensors, SWIZZLE_128B (or 64B/32B), zero OOB fill.
// dims/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
int get_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

#ifdef __CUDACC__
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int x0, int x1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(x0), "r"(x1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int x0, int x1, int x2, int x3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(bar)
      : "memory");
}
// L2 prefetch of a 2-D box (no shared-memory destination, no barrier): used before griddepcontrol.wait for operands the
// previous kernel does not produce (packed weights), so the first real loads of the main loop hit L2
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int x0, int x1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(x0), "r"(x1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
#endif

}  // namespace d3fk
