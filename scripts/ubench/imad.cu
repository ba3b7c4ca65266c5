// Microbenchmarks of the sm_100a integer pipe forms used by the Montgomery multiplier (not part of the product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rust-kzg_b200/csrc/mont.cuh"
using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// (i) plain 32x32+64 multiply-add, 8 independent accumulators
__global__ void __launch_bounds__(256) k_wide(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[8];
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x), "r"(y));
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567) sink[0] = s;
}
// (ii) carry chains: 4 independent chains of 6 pairs (like one Montgomery row x4), 48 wide-mads per iteration
__global__ void __launch_bounds__(256) k_chain(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t A[4][12], v[12];
    for (int c = 0; c < 4; c++) for (int k = 0; k < 12; k++) A[c][k] = threadIdx.x + k + c;
    for (int k = 0; k < 12; k++) v[k] = a + k * 77 + threadIdx.x;
    uint32_t y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int c = 0; c < 4; c++) Chain<6, false>::mad(A[c], v, y + c);
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 4; c++) for (int k = 0; k < 12; k++) s ^= A[c][k];
    if (s == 0x1234567) sink[0] = s;
}
// (iii) lo and hi halves as separate multiplies (IMAD + IMAD.HI), no carries: 2 multiplier ops per product
__global__ void __launch_bounds__(256) k_lohi(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t lo[8], hi[8];
    for (int k = 0; k < 8; k++) { lo[k] = threadIdx.x + k; hi[k] = k; }
    uint32_t x = a + threadIdx.x, y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(x), "r"(y));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[k]) : "r"(x), "r"(y));
            }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= lo[k] ^ hi[k];
    if (s == 0x1234567) sink[0] = s;
}

// (vi) carry-OUT only wide multiply-adds + counter capture on the ALU pipe: 8 independent 64-bit columns
__global__ void __launch_bounds__(256) k_cout(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t lo[8], hi[8], cnt[8];
    for (int k = 0; k < 8; k++) { lo[k] = threadIdx.x + k; hi[k] = k; cnt[k] = 0; }
    uint32_t x = a + threadIdx.x, y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                             : "+r"(lo[k]), "+r"(hi[k]), "+r"(cnt[k]) : "r"(x), "r"(y));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= lo[k] ^ hi[k] ^ cnt[k];
    if (s == 0x1234567) sink[0] = s;
}
// (vii) IMAD.HI alone and IMAD lo alone
__global__ void __launch_bounds__(256) k_hi(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t hi[8];
    for (int k = 0; k < 8; k++) hi[k] = k + threadIdx.x;
    uint32_t y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(hi[k]) : "r"(y));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= hi[k];
    if (s == 0x1234567) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_lo(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t lo[8];
    for (int k = 0; k < 8; k++) lo[k] = k + threadIdx.x;
    uint32_t y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(lo[k]) : "r"(y));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= lo[k];
    if (s == 0x1234567) sink[0] = s;
}
// (viii) DFMA issue rate: 8 independent chains
__global__ void __launch_bounds__(256) k_dfma(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    double acc[8];
    for (int k = 0; k < 8; k++) acc[k] = 1.0 + threadIdx.x + k;
    double x = 1.0000001 + a * 1e-9, y = 1e-7 * (b + blockIdx.x);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = __fma_rz(acc[k], x, y);
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 0.12345) sink[0] = 1;
}
// (iv) full Fp / Fr multiplications, 2 independent streams per thread
template <class F>
__global__ void __launch_bounds__(128) k_mul(uint8_t* sink, int iters) {
    F x = F::one(), y = F::rr(), z = F::one(), w = F::rr();
    x.v[0] ^= threadIdx.x; y.v[1] ^= blockIdx.x; z.v[2] ^= threadIdx.x; w.v[3] ^= blockIdx.x + 1;
    for (int it = 0; it < iters; it++) {
        x = x * y; z = z * w; y = y * x; w = w * z;
    }
    if (x.v[0] == 0x12345 && y.v[3] == 7 && z.v[1] == 1 && w.v[2] == 3) store_field(sink, x);
}
// (v) IADD3 carry chains only (12-limb add), to see the ALU-pipe cost
__global__ void __launch_bounds__(128) k_add(uint8_t* sink, int iters) {
    fp_t x = fp_t::one(), y = fp_t::rr();
    x.v[0] ^= threadIdx.x; y.v[1] ^= blockIdx.x;
    for (int it = 0; it < iters; it++) { x = x + y; y = y - x; }
    if (x.v[0] == 0x12345 && y.v[3] == 7) store_field(sink, x);
}

template <class L>
static float timeit(L launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    uint64_t* sink; CK(cudaMalloc(&sink, 4096));
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
    for (int occ : {2, 4, 8}) {
        int blocks = sms * occ, th = 256, iters = 2000;
        float ms = timeit([&] { k_wide<<<blocks, th>>>(sink, 3, 5, iters); });
        printf("wide  occ=%d: %.2f T mad/s\n", occ, (double)blocks * th * iters * 64 / ms / 1e9);
        ms = timeit([&] { k_chain<<<blocks, th>>>(sink, 3, 5, iters); });
        printf("chain occ=%d: %.2f T mad/s\n", occ, (double)blocks * th * iters * 48 / ms / 1e9);
        ms = timeit([&] { k_lohi<<<blocks, th>>>(sink, 3, 5, iters); });
        printf("lohi  occ=%d: %.2f T mul-ops/s\n", occ, (double)blocks * th * iters * 64 / ms / 1e9);
        ms = timeit([&] { k_cout<<<blocks, th>>>(sink, 3, 5, iters); });
        printf("cout  occ=%d: %.2f T mad/s\n", occ, (double)blocks * th * iters * 64 / ms / 1e9);
        ms = timeit([&] { k_dfma<<<blocks, th>>>(sink, 3, 5, iters); });
        printf("dfma  occ=%d: %.2f T dfma/s\n", occ, (double)blocks * th * iters * 64 / ms / 1e9);
        ms = timeit([&] { k_hi<<<blocks, th>>>(sink, 3, 5, iters); });
        printf("hi    occ=%d: %.2f T mad/s\n", occ, (double)blocks * th * iters * 64 / ms / 1e9);
        ms = timeit([&] { k_lo<<<blocks, th>>>(sink, 3, 5, iters); });
        printf("lo    occ=%d: %.2f T mad/s\n", occ, (double)blocks * th * iters * 64 / ms / 1e9);
    }
    for (int occ : {2, 4, 8, 12}) {
        int blocks = sms * occ, th = 128, iters = 100;
        float ms = timeit([&] { k_mul<fp_t><<<blocks, th>>>((uint8_t*)sink, iters); });
        printf("fpmul occ=%d (warps/SM=%d): %.2f G mul/s\n", occ, occ * 4, (double)blocks * th * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_mul<fpc_t><<<blocks, th>>>((uint8_t*)sink, iters); });
        printf("fpmul-compact occ=%d: %.2f G mul/s\n", occ, (double)blocks * th * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_mul<fpd_t><<<blocks, th>>>((uint8_t*)sink, iters); });
        printf("fpmul-dfma occ=%d: %.2f G mul/s\n", occ, (double)blocks * th * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_mul<frd_t><<<blocks, th>>>((uint8_t*)sink, iters); });
        printf("frmul-dfma occ=%d: %.2f G mul/s\n", occ, (double)blocks * th * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_mul<fp28_t><<<blocks, th>>>((uint8_t*)sink, iters); });
        printf("fpmul-r28 occ=%d: %.2f G mul/s\n", occ, (double)blocks * th * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_mul<fr28_t><<<blocks, th>>>((uint8_t*)sink, iters); });
        printf("frmul-r28 occ=%d: %.2f G mul/s\n", occ, (double)blocks * th * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_mul<fr_t><<<blocks, th>>>((uint8_t*)sink, iters); });
        printf("frmul occ=%d: %.2f G mul/s\n", occ, (double)blocks * th * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_add<<<blocks, th>>>((uint8_t*)sink, iters * 10); });
        printf("fpadd occ=%d: %.2f G add/s\n", occ, (double)blocks * th * iters * 10 * 2 / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
