// fft_g1.cu -- FFTG1::fft_g1 (blst/src/fft_g1.rs:13-83) on the device: radix-2 DIT over G1 points,
// out[i] = sum_j w^(i*j) * P_j, natural order in and out; inverse uses the reversed roots and a final [n^-1].
// Every butterfly carries a full 255-bit scalar multiplication of a point by a root of unity (the reference does the
// same with blst_p1_mult), so the transform is (n/2) log n scalar multiplications, one lane quad each, stage by stage
// with the working set kept in XYZZ form in HBM (192 B per point).  Same group elements as the reference, hence
// byte-identical after compression.
#include <cstdlib>

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
// Two DIT stages (s, s+1) at the latency of one.  The transform is a chain of log n dependent scalar multiplications
// (~0.8 ms each on a lane quad) with far fewer butterflies than the machine has lanes, so the chain, not the work, is the
// cost.  For the four points x0..x3 at j, j + h, j + 2h, j + 3h (h = 2^s) the two stages give
//     z0 = x0 + a x1 + b x2 + ab x3      z1 = x0 - a x1 + b' x2 - ab' x3
//     z2 = x0 + a x1 - b x2 - ab x3      z3 = x0 - a x1 - b' x2 + ab' x3
// with a = w_2h^low, b = w_4h^low, b' = w_4h^(low + h): FIVE independent products [a]x1, [b]x2, [ab]x3, [b']x2, [ab']x3 by
// roots of unity (one more than the two stages do, but all at the same depth), then eight additions.
// k_g1_stage2_mul: one lane quad per product; k_g1_stage2_comb: one lane quad per group of four points.
__global__ void __launch_bounds__(32) k_g1_stage2_mul(const uint8_t* __restrict__ work, uint8_t* __restrict__ tmp, size_t n, int log_n, int s,
                                                      const uint8_t* __restrict__ roots, size_t nmax, int inverse) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const size_t nprod = 5 * (n >> 2);
    const bool live = q < nprod;
    const size_t qq = live ? q : 0;
    const size_t g = qq / 5;
    const int which = (int)(qq - g * 5);
    const size_t h = (size_t)1 << s, low = g & (h - 1), j = ((g >> s) << (s + 2)) | low;
    // exponents in units of the n-th root
    const size_t ea = low << (log_n - 1 - s), eb = low << (log_n - 2 - s), ebp = eb + (n >> 2);
    const size_t src = which == 0 ? j + h : (which == 1 || which == 3) ? j + 2 * h : j + 3 * h;
    size_t e = which == 0 ? ea : which == 1 ? eb : which == 2 ? ea + eb : which == 3 ? ebp : ea + ebp;
    e &= n - 1;
    const uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    const int off = quad_store_offset();
    fp_t t = load_field<fp_t>(w + src * 192 + off);
    const size_t eu = e * (nmax >> log_n);
    fr_t root = load_field_ro<fr_t>(roots + (inverse && eu ? nmax - eu : eu) * 32).from_mont();
    t = quad_mul_scalar(t, root.v, table);
    if (live) store_field(tmp + ((size_t)blockIdx.y * nprod + q) * 192 + off, t);
}
__global__ void __launch_bounds__(32) k_g1_stage2_comb(uint8_t* __restrict__ work, const uint8_t* __restrict__ tmp, size_t n, int s) {
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = q < (n >> 2);
    const size_t g = live ? q : 0;
    const size_t h = (size_t)1 << s, low = g & (h - 1), j = ((g >> s) << (s + 2)) | low;
    uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    const uint8_t* pr = tmp + ((size_t)blockIdx.y * 5 * (n >> 2) + g * 5) * 192;
    const int off = quad_store_offset();
    const bool yrole = (threadIdx.x & 3) == 1;
    auto neg = [&](const fp_t& v) { return yrole ? v.neg() : v; };
    const fp_t x0 = load_field<fp_t>(w + j * 192 + off);
    const fp_t p0 = load_field<fp_t>(pr + off), p1 = load_field<fp_t>(pr + 192 + off), p2 = load_field<fp_t>(pr + 2 * 192 + off),
               p3 = load_field<fp_t>(pr + 3 * 192 + off), p4 = load_field<fp_t>(pr + 4 * 192 + off);
    const fp_t u = quad_add(x0, p0), v = quad_add(x0, neg(p0));
    const fp_t pp = quad_add(p1, p2), qd = quad_add(p3, neg(p4));
    const fp_t z0 = quad_add(u, pp), z2 = quad_add(u, neg(pp)), z1 = quad_add(v, qd), z3 = quad_add(v, neg(qd));
    if (live) {
        store_field(w + j * 192 + off, z0);
        store_field(w + (j + h) * 192 + off, z1);
        store_field(w + (j + 2 * h) * 192 + off, z2);
        store_field(w + (j + 3 * h) * 192 + off, z3);
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
    // stage 0 multiplies by w^0 only; after it the stages run in fused pairs (k_g1_stage2_*: two stages at the latency of
    // one scalar multiplication) while the launch is small: one warp per eight products with a 27 KiB table each, so
    // ~1100 warps are resident at a time -- up to 2^12 points per launch the fused products fit one wave (one blob's FK20
    // transforms: cells + proofs 12.0 -> 6.9 ms); at 64 blobs x 128 points they spill into a second wave and the plain
    // stages, 20 % less work, win (18.6 vs 21.7 ms per 64 blobs, scripts/fk20_timing.py)
    const int fuse_env = getenv("B200_FFT_G1_FUSE") ? atoi(getenv("B200_FFT_G1_FUSE")) : -1;   // per call: tests toggle it
    const bool fuse = fuse_env >= 0 ? fuse_env != 0 : total <= ((size_t)1 << 12);
    int s = 0;
    if (!fuse || (log_n & 1)) {
        if (log_n > 0) {
            k_g1_stage<<<dim3(div_up(n / 2 * 4, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, n, log_n, 0, (const uint8_t*)roots_,
                                                                                max_width_, inverse);
            launches_++;
        }
        s = 1;
    }
    if (fuse && log_n >= 2) {
        const size_t need = (size_t)batch * 5 * (n >> 2);
        if (need > g1_tmp_elems_) {
            cudaFree(g1_tmp_);
            g1_tmp_ = nullptr; g1_tmp_elems_ = 0;
            g1_tmp_ = dev_alloc<uint8_t>(need * 192);
            g1_tmp_elems_ = need;
        }
        for (; s + 2 <= log_n; s += 2) {
            k_g1_stage2_mul<<<dim3(div_up(5 * (n >> 2) * 4, 32), (unsigned)batch), 32, 0, st>>>((const uint8_t*)g1_work_, (uint8_t*)g1_tmp_, n, log_n, s,
                                                                                           (const uint8_t*)roots_, max_width_, inverse);
            k_g1_stage2_comb<<<dim3(div_up((n >> 2) * 4, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, (const uint8_t*)g1_tmp_, n, s);
            launches_ += 2;
        }
    }
    for (; s < log_n; s++) {
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
