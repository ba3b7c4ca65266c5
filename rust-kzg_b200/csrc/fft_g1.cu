// fft_g1.cu -- FFTG1::fft_g1 (blst/src/fft_g1.rs:13-83) on the device: radix-2 DIT over G1 points,
// out[i] = sum_j w^(i*j) * P_j, natural order in and out; inverse uses the reversed roots and a final [n^-1].
// Every butterfly carries a full 255-bit scalar multiplication of a point by a root of unity (the reference does the
// same with blst_p1_mult), so the transform is (n/2) log n scalar multiplications, one thread each, stage by stage
// with the working set kept in XYZZ form in HBM (192 B per point).  Same group elements as the reference, hence
// byte-identical after compression.
#include "g1.cuh"
#include "ntt.cuh"
#include "util.cuh"

namespace b200 {

// acc = [k] p, k a canonical 256-bit scalar (little-endian words), left-to-right double-and-add on XYZZ
__device__ __forceinline__ xyzz_t xyzz_mul_scalar(const xyzz_t& p, const fr_t& k_canon) {
    int top = -1;
#pragma unroll 1
    for (int i = 7; i >= 0; i--)
        if (k_canon.v[i]) { top = 32 * i + 31 - __clz(k_canon.v[i]); break; }
    if (top < 0 || p.is_inf()) return xyzz_t::inf();
    xyzz_t acc = p;
#pragma unroll 1
    for (int bit = top - 1; bit >= 0; bit--) {
        xyzz_dbl(acc);
        if ((k_canon.v[bit >> 5] >> (bit & 31)) & 1) xyzz_add(acc, p);
    }
    return acc;
}

// bit-reversal permutation + Jacobian -> XYZZ
__global__ void k_g1_brp_in(const uint8_t* __restrict__ in_jac, uint8_t* __restrict__ work, size_t n, int log_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t base = (size_t)blockIdx.y * n;
    size_t r = log_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
    cc::xyzz_t p = cc::jac_to_xyzz(cc::load_jac(in_jac + (base + i) * 144));
    cc::store_xyzz(work + (base + r) * 192, p);
}
// one DIT stage: butterflies (i, i + 2^s) with twiddle w_n^(k * n / 2^(s+1)) (blst/src/fft_g1.rs:43-47)
__global__ void __launch_bounds__(128) k_g1_stage(uint8_t* __restrict__ work, size_t n, int log_n, int s, const uint8_t* __restrict__ roots,
                                                  size_t nmax, int inverse) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n / 2) return;
    uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    size_t half = (size_t)1 << s;
    size_t lowk = b & (half - 1);
    size_t i = ((b >> s) << (s + 1)) | lowk;
    xyzz_t lo = load_xyzz(w + i * 192), hi = load_xyzz(w + (i + half) * 192);
    xyzz_t t = hi;
    if (lowk) {
        size_t e = (lowk << (log_n - 1 - s)) * (nmax >> log_n);
        fr_t root = load_field_ro<fr_t>(roots + (inverse ? nmax - e : e) * 32).from_mont();
        t = xyzz_mul_scalar(hi, root);
    }
    xyzz_t nt = t;
    nt.y = nt.y.neg();
    xyzz_t sum = lo, dif = lo;
    xyzz_add(sum, t);
    xyzz_add(dif, nt);
    store_xyzz(w + i * 192, sum);
    store_xyzz(w + (i + half) * 192, dif);
}
// XYZZ -> Jacobian, with the [n^-1] scaling of the inverse transform (blst/src/fft_g1.rs:74-79)
__global__ void __launch_bounds__(128) k_g1_out(const uint8_t* __restrict__ work, uint8_t* __restrict__ out_jac, size_t total,
                                                const uint8_t* __restrict__ scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    xyzz_t p = load_xyzz(work + i * 192);
    if (scale) p = xyzz_mul_scalar(p, load_field_ro<fr_t>(scale).from_mont());
    store_jac(out_jac + i * 144, xyzz_to_jac(p));
}

void FFTSettingsDev::fft_g1(const void* in_jac_dev, void* out_jac_dev, size_t n, bool inverse, int batch, cudaStream_t st) {
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
        k_g1_stage<<<dim3(div_up(n / 2, 128), (unsigned)batch), 128, 0, st>>>((uint8_t*)g1_work_, n, log_n, s, (const uint8_t*)roots_,
                                                                            max_width_, inverse);
        launches_++;
    }
    const uint8_t* inv_n = (const uint8_t*)roots_ + (max_width_ + 1) * 32 + 33 * 32;
    k_g1_out<<<div_up(total, 128), 128, 0, st>>>((const uint8_t*)g1_work_, (uint8_t*)out_jac_dev, total,
                                                 inverse && log_n ? inv_n + log_n * 32 : nullptr);
    launches_ += 2;
    B200_LAUNCH_CHECK();
}

}  // namespace b200
