// selftest.cu -- device self-test hooks (elementwise field / point kernels) and the integer-pipe microbenchmark.
// These exist so the parity tests can pin the device arithmetic against the CPU oracle one operation at a time
//  and so bench.py can report a MEASURED integer-multiply peak next to the HBM peak.
#include "../../include/b200_kzg.h"
#include "capi_common.cuh"
#include "g1.cuh"
#include "util.cuh"

using namespace b200;

// default: unrolled carry-chain multiplier; op + 16 selects the compact one, op + 32 the radix-2^28 one
template <class F>
__global__ void k_field_op(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int B = F::N * 4;
    F x = load_field<F>(a + i * B), y = b ? load_field<F>(b + i * B) : F::zero(), r;
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.neg(); break;
        case 4: r = x.inverse(); break;
        case 7: r = x.inverse_fermat(); break;
        case 9: r = x.inverse_euclid(); break;
        case 8: r = x.sqr(); break;
        case 5: r = x.to_mont(); break;
        default: r = x.from_mont(); break;
    }
    store_field(out + i * B, r);
}

__global__ void k_p1_add(uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, int mixed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    jac_t pa = load_jac(a + i * 144), pb = load_jac(b + i * 144);
    xyzz_t acc = jac_to_xyzz(pa);
    if (mixed) {
        affine_t q = jac_to_affine(pb);
        xyzz_add_affine(acc, q);
    } else {
        xyzz_t q = jac_to_xyzz(pb);
        xyzz_add(acc, q);
    }
    store_jac(out + i * 144, xyzz_to_jac(acc));
}

__global__ void k_p1_compress(uint8_t* out, const uint8_t* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_compress(out + i * 48, jac_to_affine(load_jac(p + i * 144)));
}

// IMAD.WIDE issue rate of the FMA-heavy pipe: four independent carry chains of six 32x32+64 multiply-adds per thread
// (exactly the instruction the Montgomery rows are made of).  The operands are loop-carried, so ptxas cannot hoist
// the products (a loop-invariant product turns the loop into 64-bit additions and reports a bogus 2x rate).
__global__ void __launch_bounds__(256) k_imad_bench(uint64_t* sink, uint32_t a, uint32_t b, int iters) {
    uint32_t A[4][12], v[12];
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 12; k++) A[c][k] = threadIdx.x + k + c;
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
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 12; k++) s ^= A[c][k];
    if (s == 0x1234567) sink[0] = s;
}
template <class F>
__global__ void __launch_bounds__(128) k_mul_bench(uint8_t* sink, int iters) {
    F x = F::one(), y = F::rr();
    x.v[0] ^= threadIdx.x;
    y.v[1] ^= blockIdx.x;
    for (int it = 0; it < iters; it++) {
        x = x * y;
        y = y * x;
    }
    if (x.v[0] == 0x12345 && y.v[3] == 7) store_field(sink, x);
}
__global__ void __launch_bounds__(128) k_fpmul_bench(uint8_t* sink, int iters) {
    fp_t x = fp_t::one(), y = fp_t::rr();
    x.v[0] ^= threadIdx.x;
    y.v[1] ^= blockIdx.x;
    for (int it = 0; it < iters; it++) {
        x = x * y;
        y = y * x;
    }
    if (x.v[0] == 0x12345 && y.v[3] == 7) store_field(sink, x);
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    explicit DevBuf(size_t bytes) { p = reinterpret_cast<T*>(dev_alloc<uint8_t>(bytes)); }
    ~DevBuf() { cudaFree(p); }
};

template <class F>
static RustError field_op(int op, void* out, const void* a, const void* b, size_t n) {
    return guarded([&] {
        require_device();
        if (n == 0) return;
        size_t bytes = n * F::N * 4;
        DevBuf<uint8_t> da(bytes), db(bytes), dout(bytes);
        B200_CUDA_CHECK(cudaMemcpy(da.p, a, bytes, cudaMemcpyHostToDevice));
        if (b) B200_CUDA_CHECK(cudaMemcpy(db.p, b, bytes, cudaMemcpyHostToDevice));
        if (op & 128)
            k_field_op<Mont<typename F::params_t, MONT_KARA>><<<div_up(n, 128), 128>>>(op & 15, dout.p, da.p, b ? db.p : nullptr, n);
        else if (op & 64)
            k_field_op<Mont<typename F::params_t, MONT_DFMA>><<<div_up(n, 128), 128>>>(op & 15, dout.p, da.p, b ? db.p : nullptr, n);
        else if (op & 32)
            k_field_op<Mont<typename F::params_t, MONT_R28>><<<div_up(n, 128), 128>>>(op & 15, dout.p, da.p, b ? db.p : nullptr, n);
        else if (op & 16)
            k_field_op<Mont<typename F::params_t, MONT_COMPACT>><<<div_up(n, 128), 128>>>(op & 15, dout.p, da.p, b ? db.p : nullptr, n);
        else
            k_field_op<F><<<div_up(n, 128), 128>>>(op, dout.p, da.p, b ? db.p : nullptr, n);
        B200_LAUNCH_CHECK();
        B200_CUDA_CHECK(cudaMemcpy(out, dout.p, bytes, cudaMemcpyDeviceToHost));
    });
}

extern "C" {

RustError b200_selftest_fp(int op, blst_fp* out, const blst_fp* a, const blst_fp* b, size_t n) {
    return field_op<fp_t>(op, out, a, b, n);
}
RustError b200_selftest_fr(int op, blst_fr* out, const blst_fr* a, const blst_fr* b, size_t n) {
    return field_op<fr_t>(op, out, a, b, n);
}
RustError b200_selftest_p1_add(blst_p1* out, const blst_p1* a, const blst_p1* b, size_t n, int mixed) {
    return guarded([&] {
        require_device();
        if (n == 0) return;
        DevBuf<uint8_t> da(n * 144), db(n * 144), dout(n * 144);
        B200_CUDA_CHECK(cudaMemcpy(da.p, a, n * 144, cudaMemcpyHostToDevice));
        B200_CUDA_CHECK(cudaMemcpy(db.p, b, n * 144, cudaMemcpyHostToDevice));
        k_p1_add<<<div_up(n, 64), 64>>>(dout.p, da.p, db.p, n, mixed);
        B200_LAUNCH_CHECK();
        B200_CUDA_CHECK(cudaMemcpy(out, dout.p, n * 144, cudaMemcpyDeviceToHost));
    });
}
RustError b200_selftest_p1_compress(uint8_t* out48, const blst_p1* p, size_t n) {
    return guarded([&] {
        require_device();
        if (n == 0) return;
        DevBuf<uint8_t> dp(n * 144), dout(n * 48);
        B200_CUDA_CHECK(cudaMemcpy(dp.p, p, n * 144, cudaMemcpyHostToDevice));
        k_p1_compress<<<div_up(n, 64), 64>>>(dout.p, dp.p, n);
        B200_LAUNCH_CHECK();
        B200_CUDA_CHECK(cudaMemcpy(out48, dout.p, n * 48, cudaMemcpyDeviceToHost));
    });
}

RustError b200_microbench_int(double* imad_per_s, double* fpmul_per_s) {
    return guarded([&] {
        require_device();
        cudaDeviceProp prop;
        B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, 0));
        int sms = prop.multiProcessorCount;
        DevBuf<uint64_t> sink(4096);
        cudaEvent_t e0, e1;
        B200_CUDA_CHECK(cudaEventCreate(&e0));
        B200_CUDA_CHECK(cudaEventCreate(&e1));
        float ms = 0;
        {
            int iters = 2000, blocks = sms * 8, threads = 256;
            k_imad_bench<<<blocks, threads>>>(sink.p, 3, 5, 50);  // warm-up
            B200_CUDA_CHECK(cudaEventRecord(e0));
            k_imad_bench<<<blocks, threads>>>(sink.p, 3, 5, iters);
            B200_CUDA_CHECK(cudaEventRecord(e1));
            B200_CUDA_CHECK(cudaEventSynchronize(e1));
            B200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (imad_per_s) *imad_per_s = (double)blocks * threads * iters * 48.0 / (ms * 1e-3);
        }
        {
            int iters = 200, blocks = sms * 16, threads = 128;
            k_fpmul_bench<<<blocks, threads>>>((uint8_t*)sink.p, 5);
            B200_CUDA_CHECK(cudaEventRecord(e0));
            k_fpmul_bench<<<blocks, threads>>>((uint8_t*)sink.p, iters);
            B200_CUDA_CHECK(cudaEventRecord(e1));
            B200_CUDA_CHECK(cudaEventSynchronize(e1));
            B200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (fpmul_per_s) *fpmul_per_s = (double)blocks * threads * iters * 2.0 / (ms * 1e-3);
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

/* multiplications per second of a multiplier variant in a dependent-chain loop on all SMs.
 * field: 0 Fp, 1 Fr; mode: 0 unrolled CIOS (default), 5 Karatsuba a*b + row-wise reduction */
RustError b200_microbench_mul(int field, int mode, double* mul_per_s) {
    return guarded([&] {
        require_device();
        int dev = 0, sms = 0;
        B200_CUDA_CHECK(cudaGetDevice(&dev));
        B200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        DevBuf<uint8_t> sink(64);
        cudaEvent_t e0, e1;
        B200_CUDA_CHECK(cudaEventCreate(&e0));
        B200_CUDA_CHECK(cudaEventCreate(&e1));
        const int iters = 200, blocks = sms * 16, threads = 128;
        auto launch = [&](int it) {
            if (field == 0 && mode == 5) k_mul_bench<fpk_t><<<blocks, threads>>>(sink.p, it);
            else if (field == 0) k_mul_bench<fp_t><<<blocks, threads>>>(sink.p, it);
            else if (mode == 5) k_mul_bench<frk_t><<<blocks, threads>>>(sink.p, it);
            else k_mul_bench<fr_t><<<blocks, threads>>>(sink.p, it);
        };
        launch(5);
        B200_CUDA_CHECK(cudaEventRecord(e0));
        launch(iters);
        B200_CUDA_CHECK(cudaEventRecord(e1));
        B200_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        B200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        B200_LAUNCH_CHECK();
        if (mul_per_s) *mul_per_s = (double)blocks * threads * iters * 2.0 / (ms * 1e-3);
    });
}

}  // extern "C"
