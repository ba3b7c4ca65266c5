// two-kernel target for ncu: the carry-chain and the radix-2^28 Fp multipliers in an identical loop
#include <cstdio>
#include <cuda_runtime.h>
#include "../../rust-kzg_b200/csrc/mont.cuh"
using namespace b200;
template <class F>
__global__ void __launch_bounds__(128) k_mulloop(uint8_t* sink, int iters) {
    F x = F::one(), y = F::rr(), z = F::one(), w = F::rr();
    x.v[0] ^= threadIdx.x; y.v[1] ^= blockIdx.x; z.v[2] ^= threadIdx.x; w.v[3] ^= blockIdx.x + 1;
    for (int it = 0; it < iters; it++) { x = x * y; z = z * w; y = y * x; w = w * z; }
    if (x.v[0] == 0x12345 && y.v[3] == 7 && z.v[1] == 1 && w.v[2] == 3) store_field(sink, x);
}
int main() {
    uint8_t* sink; cudaMalloc(&sink, 4096);
    int sms = 148;
    for (int r = 0; r < 2; r++) {
        k_mulloop<fpu_t><<<sms * 8, 128>>>(sink, 100);
        k_mulloop<fp_t><<<sms * 8, 128>>>(sink, 100);
    }
    cudaDeviceSynchronize();
    printf("done\n");
    return 0;
}
