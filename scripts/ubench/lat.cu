// single-warp latency of dependent group / field operations (warm instruction cache: one code copy in a loop),
// and the same operations "cold" (a long straight-line sequence of distinct copies), to separate pipe latency
// from instruction-fetch cost in the latency-bound reduce tails.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../rust-kzg_b200/csrc/g1.cuh"
using namespace b200;

__device__ __forceinline__ fp_t shfl_fp(const fp_t& v, int d) {
    fp_t o;
#pragma unroll
    for (int k = 0; k < 12; k++) o.v[k] = __shfl_down_sync(0xffffffffu, v.v[k], d);
    return o;
}
__device__ __forceinline__ xyzz_t shfl_down_xyzz(const xyzz_t& v, int d) {
    xyzz_t o;
    o.x = shfl_fp(v.x, d); o.y = shfl_fp(v.y, d); o.zzz = shfl_fp(v.zzz, d); o.zz = shfl_fp(v.zz, d);
    return o;
}
__device__ __forceinline__ affine_t gen_affine(int t) {
    affine_t p;
    p.x = fp_t::one(); p.y = fp_t::rr();
    p.x.v[0] ^= t + 1; p.y.v[1] ^= 3 * t + 7;
    return p;
}
template <int OP>
__global__ void __launch_bounds__(128) k_lat(uint8_t* sink, int iters) {
    affine_t p = gen_affine(threadIdx.x);
    xyzz_t acc = affine_to_xyzz(p), b = affine_to_xyzz(gen_affine(threadIdx.x + 77));
    fp_t f = p.x, g = p.y;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (OP == 0) xyzz_add(acc, b);
        if (OP == 1) xyzz_add_affine(acc, p);
        if (OP == 2) xyzz_dbl(acc);
        if (OP == 3) { f = f * g; }
        if (OP == 4) { f = f * g; g = g * f; f = f * g; g = g * f; f = f * g; g = g * f; f = f * g; g = g * f; }
        if (OP == 6) { f = f + g; g = g - f; f = f + g; g = g - f; f = f + g; g = g - f; f = f + g; g = g - f; }
        if (OP == 7) { f = f.sqr(); }
        if (OP == 5) { xyzz_t o = shfl_down_xyzz(acc, 1); if ((threadIdx.x & 31) == 31) o = b; xyzz_add(acc, o); }
    }
    if (OP == 3 || OP == 4 || OP == 6 || OP == 7) { acc.x = f + g; acc.y = f - g; }
    if (acc.x.v[0] == 0x12345 && acc.y.v[3] == 7) store_xyzz(sink, acc);
}
// straight-line: 8 distinct inlined copies per iteration
template <int OP>
__global__ void __launch_bounds__(128) k_lat8(uint8_t* sink, int iters) {
    affine_t p = gen_affine(threadIdx.x);
    xyzz_t acc = affine_to_xyzz(p), b = affine_to_xyzz(gen_affine(threadIdx.x + 77));
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) xyzz_add(acc, b);
            if (OP == 2) xyzz_dbl(acc);
        }
    }
    if (acc.x.v[0] == 0x12345 && acc.y.v[3] == 7) store_xyzz(sink, acc);
}
template <class K> float timeit(K launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    uint8_t* sink; cudaMalloc(&sink, 4096);
    const char* names[] = {"xyzz_add", "xyzz_add_affine", "xyzz_dbl", "fp_mul", "fp_mul x8 (unrolled)", "shfl+xyzz_add", "fp add/sub x8", "fp_sqr"};
    for (int threads : {32, 128}) {
        int iters = 200;
#define RUN(OP, PER) { float ms = timeit([&] { k_lat<OP><<<1, threads>>>(sink, iters); }); \
        printf("threads=%3d %-24s warm loop : %8.3f us/op\n", threads, names[OP], ms * 1e3 / (iters * PER)); }
        RUN(0, 1) RUN(1, 1) RUN(2, 1) RUN(3, 1) RUN(4, 8) RUN(5, 1) RUN(6, 8) RUN(7, 1)
        { float ms = timeit([&] { k_lat8<0><<<1, threads>>>(sink, 1); });
          printf("threads=%3d xyzz_add x8 straight-line, 1 pass (cold): %8.3f us/op\n", threads, ms * 1e3 / 8); }
        { float ms = timeit([&] { k_lat8<0><<<1, threads>>>(sink, 25); });
          printf("threads=%3d xyzz_add x8 straight-line, 25 passes    : %8.3f us/op\n", threads, ms * 1e3 / 200); }
        { float ms = timeit([&] { k_lat8<2><<<1, threads>>>(sink, 1); });
          printf("threads=%3d xyzz_dbl x8 straight-line, 1 pass (cold): %8.3f us/op\n", threads, ms * 1e3 / 8); }
    }
    // empty-launch floor
    { float ms = timeit([&] { k_lat<3><<<1, 32>>>(sink, 0); }); printf("launch floor: %.3f us\n", ms * 1e3); }
    return 0;
}
