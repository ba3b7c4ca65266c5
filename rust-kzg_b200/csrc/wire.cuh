// wire.cuh -- big-endian wire formats of field elements (Fr::from_bytes / to_bytes, blst/src/types/fr.rs:64-136).
#pragma once
#include "mont.cuh"

namespace b200 {

// 32 big-endian bytes -> 8 little-endian words (p must be 4-byte aligned)
__device__ __forceinline__ void load_be32(const uint8_t* p, uint32_t w[8]) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);  // blobs are 32-byte aligned inside a 128 KiB array
#pragma unroll
    for (int i = 0; i < 8; i++) w[7 - i] = __byte_perm(q[i], 0, 0x0123);
}
__device__ __forceinline__ bool lt_r(const uint32_t w[8]) {  // w < r ?
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        uint32_t m = FrParams::mod(i);
        if (w[i] < m) return true;
        if (w[i] > m) return false;
    }
    return false;
}
// Montgomery Fr -> 32 canonical big-endian bytes (p 4-byte aligned)
__device__ __forceinline__ void store_fr_be32(uint8_t* p, const fr_t& mont) {
    fr_t v = mont.from_mont();
    uint32_t* o = reinterpret_cast<uint32_t*>(p);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = __byte_perm(v.v[7 - k], 0, 0x0123);
}

}  // namespace b200
