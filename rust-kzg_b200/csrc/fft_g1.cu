// fft_g1.cu -- FFTG1::fft_g1 (blst/src/fft_g1.rs:13-83) on the device: radix-2 DIT over G1 points,
// out[i] = sum_j w^(i*j) * P_j, natural order in and out; inverse uses the reversed roots and a final [n^-1].
// Every butterfly carries a full 255-bit scalar multiplication of a point by a root of unity (the reference does the
// same with blst_p1_mult), so the transform is (n/2) log n scalar multiplications, one lane quad each, stage by stage
// with the working set kept in XYZZ form in HBM (192 B per point).  Same group elements as the reference, hence
// byte-identical after compression.
#include "g1.cuh"
#include "g1_quad.cuh"
#include "ntt.cuh"
#include "util.cuh"

namespace b200 {

// bit-reversal permutation + Jacobian -> XYZZ
__global__ void k_g1_brp_in(const uint8_t* __restrict__ in_jac, uint8_t* __restrict__ work, size_t n, int log_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t base = (size_t)blockIdx.y * n;
    size_t r = log_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
    cc::xyzz_t p = cc::jac_to_xyzz(cc::load_jac(in_jac + (base + i) * 144));
    cc::store_xyzz(work + (base + r) * 192, p);
}
// one DIT stage: butterflies (i, i + 2^s) with twiddle w_n^(k * n / 2^(s+1)) (blst/src/fft_g1.rs:43-47).
// One QUAD of lanes per butterfly (g1_quad.cuh): the transform is a chain of log n full scalar multiplications and
// there are far fewer butterflies than lanes on the machine, so each multiplication is spread over four lanes.
__global__ void __launch_bounds__(32) k_g1_stage(uint8_t* __restrict__ work, size_t n, int log_n, int s, const uint8_t* __restrict__ roots,
                                                 size_t nmax, int inverse) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = q < n / 2;
    const size_t b = live ? q : 0;
    uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    const size_t half = (size_t)1 << s;
    // butterflies are numbered twiddle-major: all those with the same twiddle index lowk are consecutive, so the ones
    // multiplying by w^0 = 1 (a 2^-s fraction of the stage) fill whole warps, which then skip the scalar multiplication
    const int hi_bits = log_n - 1 - s;
    const size_t lowk = b >> hi_bits;
    const size_t i = ((b & (((size_t)1 << hi_bits) - 1)) << (s + 1)) | lowk;
    const int off = quad_store_offset();
    fp_t lo = load_field<fp_t>(w + i * 192 + off), t = load_field<fp_t>(w + (i + half) * 192 + off);
    if (__any_sync(kFullMask, lowk != 0)) {
        // butterflies with lowk == 0 multiply by roots[0] = 1
        size_t e = (lowk << (log_n - 1 - s)) * (nmax >> log_n);
        fr_t root = load_field_ro<fr_t>(roots + (inverse && e ? nmax - e : e) * 32).from_mont();
        t = quad_mul_scalar(t, root.v, table);
    }
    fp_t nt = (threadIdx.x & 3) == 1 ? t.neg() : t;
    fp_t sum = quad_add(lo, t), dif = quad_add(lo, nt);
    if (live) {
        store_field(w + i * 192 + off, sum);
        store_field(w + (i + half) * 192 + off, dif);
    }
}
// XYZZ -> Jacobian, with the [n^-1] scaling of the inverse transform (blst/src/fft_g1.rs:74-79); one quad per point
__global__ void __launch_bounds__(32) k_g1_out(const uint8_t* __restrict__ work, uint8_t* __restrict__ out_jac, size_t total,
                                               const uint8_t* __restrict__ scale) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = q < total;
    const size_t i = live ? q : 0;
    const int role = threadIdx.x & 3;
    fp_t p = load_field<fp_t>(work + i * 192 + quad_store_offset());
    if (scale) p = quad_mul_scalar(p, load_field_ro<fr_t>(scale).from_mont().v, table);
    // Jacobian (X*ZZ, Y*ZZZ, ZZ), see xyzz_to_jac
    fp_t t = p * shfl_xor_fp(p, 2);
    if (live && role < 2) store_field(out_jac + i * 144 + role * 48, t);
    if (live && role == 2) store_field(out_jac + i * 144 + 96, p);
}

void FFTSettingsDev::fft_g1(const void* in_jac_dev, void* out_jac_dev, size_t n, bool inverse, int batch, cudaStream_t st,
                            bool apply_scale) {
    // argument checks of FFTG1::fft_g1 (blst/src/fft_g1.rs:55-61)
    if (n > max_width_) throw CudaError(1, "Supplied list is longer than the available max width");
    if (n == 0 || (n & (n - 1))) throw CudaError(1, "A list with power-of-two length expected");
    int log_n = 0;
    while (((size_t)1 << log_n) < n) log_n++;
    size_t total = (size_t)batch * n;
    if (total * 6 > g1_work_elems_ * 6) {
        cudaFree(g1_work_);
        g1_work_ = dev_alloc<uint8_t>(total * 192);
        g1_work_elems_ = total;
    }
    launches_ = 0;
    k_g1_brp_in<<<dim3(div_up(n, 128), (unsigned)batch), 128, 0, st>>>((const uint8_t*)in_jac_dev, (uint8_t*)g1_work_, n, log_n);
    for (int s = 0; s < log_n; s++) {
        k_g1_stage<<<dim3(div_up(n / 2 * 4, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, n, log_n, s, (const uint8_t*)roots_,
                                                                            max_width_, inverse);
        launches_++;
    }
    const uint8_t* inv_n = (const uint8_t*)roots_ + (max_width_ + 1) * 32 + 33 * 32;
    k_g1_out<<<div_up(total * 4, 32), 32, 0, st>>>((const uint8_t*)g1_work_, (uint8_t*)out_jac_dev, total,
                                                 inverse && log_n && apply_scale ? inv_n + log_n * 32 : nullptr);
    launches_ += 2;
    B200_LAUNCH_CHECK();
}

}  // namespace b200
