// Host-side TMA tensor-map encoding (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link) and
// the device-side tensor copy instruction.
#pragma once
#include <cuda.h>

#include "async.cuh"
#include "common.cuh"

namespace forge {

// Encode a tiled tensor map over `base`: dims[rank] elements (innermost first), strides_bytes[rank - 1] (of dims 1..),
// box[rank] elements.  Returns 0 or a fail() code.
int encode_tensor_map(const char* fn, CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                      const unsigned long long* dims, const unsigned long long* strides_bytes, const unsigned* box,
                      CUtensorMapSwizzle swizzle);

namespace async_ {

// 4-D tiled TMA load global -> shared (UTMALDG); completes the box's bytes on the mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(map)) : "memory");
}

}  // namespace async_
}  // namespace forge
