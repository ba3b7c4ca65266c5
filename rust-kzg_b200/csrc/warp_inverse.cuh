// warp_inverse.cuh -- one field inversion shared by the 32 lanes of a warp.
#pragma once
#include "mont.cuh"

namespace b200 {

// Field inversion shared by a warp.  The binary-Euclid inverse is a data-dependent loop: 32 lanes inverting 32
// different values run the union of their branch sequences (2.3x the time of one inversion, measured).  Instead the
// warp forms prefix and suffix products of its inputs with two shuffle scans, every lane inverts the SAME total (no
// divergence), and 1/z_i = (1/total) * prefix_{i-1} * suffix_{i+1}.
template <class F>
__device__ __forceinline__ F shfl_field(const F& a, int src) {
    F r;
#pragma unroll
    for (int i = 0; i < F::N; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
    return r;
}
// 1/z on every lane (z != 0); must be called by all 32 lanes of the warp
template <class F>
__device__ __forceinline__ F warp_inverse(const F& z) {
    const int lane = threadIdx.x & 31;
    F pre = z, suf = z;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        F t = shfl_field(pre, lane >= d ? lane - d : lane);
        F u = shfl_field(suf, lane + d < 32 ? lane + d : lane);
        if (lane >= d) pre = pre * t;
        if (lane + d < 32) suf = suf * u;
    }
    F inv = shfl_field(pre, 31).inverse();                 // identical on all lanes
    F pl = shfl_field(pre, lane ? lane - 1 : 0), sr = shfl_field(suf, lane < 31 ? lane + 1 : 31);
    if (lane) inv = inv * pl;
    if (lane < 31) inv = inv * sr;
    return inv;
}


}  // namespace b200
