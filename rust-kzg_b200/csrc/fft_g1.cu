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

#include <algorithm>
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
// SPLIT: two neighbouring quads per butterfly, one GLV half of the scalar multiplication each (quad_mul_scalar_half: a 20 %
// shorter chain on twice the lanes); afterwards the first quad forms the sum, the second the difference.
template <bool SPLIT>
__global__ void __launch_bounds__(32) k_g1_stage(uint8_t* __restrict__ work, size_t n, int log_n, int s, const uint8_t* __restrict__ roots,
                                                 size_t nmax, int inverse) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const size_t q2 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const size_t q = SPLIT ? q2 >> 1 : q2;
    const int half = SPLIT ? (int)(q2 & 1) : 0;
    const bool live = q < n / 2;
    const size_t b = live ? q : 0;
    uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    const size_t hstride = (size_t)1 << s;
    // butterflies are numbered twiddle-major: all those with the same twiddle index lowk are consecutive, so the ones
    // multiplying by w^0 = 1 (a 2^-s fraction of the stage) fill whole warps, which then skip the scalar multiplication
    const int hi_bits = log_n - 1 - s;
    const size_t lowk = b >> hi_bits;
    const size_t i = ((b & (((size_t)1 << hi_bits) - 1)) << (s + 1)) | lowk;
    const int off = quad_store_offset();
    fp_t lo = load_field<fp_t>(w + i * 192 + off), t = load_field<fp_t>(w + (i + hstride) * 192 + off);
    if (__any_sync(kFullMask, lowk != 0)) {
        // butterflies with lowk == 0 multiply by roots[0] = 1
        size_t e = (lowk << (log_n - 1 - s)) * (nmax >> log_n);
        fr_t root = load_field_ro<fr_t>(roots + (inverse && e ? nmax - e : e) * 32).from_mont();
        if (SPLIT) {
            t = quad_mul_scalar_half(t, root.v, table, half);
            t = quad_add(t, shfl_xor_fp(t, 4));                 // [k1] t + [k2] phi(t), on both quads
        } else {
            t = quad_mul_scalar(t, root.v, table);
        }
    }
    fp_t nt = (threadIdx.x & 3) == 1 ? t.neg() : t;
    if (SPLIT) {
        fp_t res = quad_add(lo, half ? nt : t);
        if (live) store_field(w + (i + (half ? hstride : 0)) * 192 + off, res);
    } else {
        fp_t sum = quad_add(lo, t), dif = quad_add(lo, nt);
        if (live) {
            store_field(w + i * 192 + off, sum);
            store_field(w + (i + hstride) * 192 + off, dif);
        }
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
template <bool SPLIT>
__global__ void __launch_bounds__(32) k_g1_stage2_mul(const uint8_t* __restrict__ work, uint8_t* __restrict__ tmp, size_t n, int log_n, int s,
                                                      const uint8_t* __restrict__ roots, size_t nmax, int inverse) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const size_t q2 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const size_t q = SPLIT ? q2 >> 1 : q2;
    const int ghalf = SPLIT ? (int)(q2 & 1) : 0;
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
    if (SPLIT) {
        t = quad_mul_scalar_half(t, root.v, table, ghalf);
        t = quad_add(t, shfl_xor_fp(t, 4));
    } else {
        t = quad_mul_scalar(t, root.v, table);
    }
    if (live && ghalf == 0) store_field(tmp + ((size_t)blockIdx.y * nprod + q) * 192 + off, t);
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
// Three DIT stages (s, s+1, s+2) at the latency of one, for transforms so small that even the fused pairs leave the machine
// idle (one blob's FK20 transforms: 128 points).  For the eight points x0..x7 at j + m h (h = 2^s) with a = w_2h^low,
// b / b' = w_4h^(low [+ h]), c_r = w_8h^(low + r h) and beta_r = b (r even) or b' (r odd):
//     z_r, z_(r+4) = u_r +- V_r,   u_r = (x0 +- a x1) +- beta_r (x2 +- a x3),   V_r = c_r [(x4 +- a x5) +- beta_r (x6 +- a x7)]
// (inner signs: - for odd r; outer signs: - for r >= 2).  Expanded, that is 21 independent products by roots of unity per
// group -- [a]x1, [b]x2, [b']x2, [ab]x3, [ab']x3 and [c_r]x4, [a c_r]x5, [c_r beta_r]x6, [a c_r beta_r]x7 for r = 0..3 --
// instead of the 12 of three plain stages, all at the same depth, then 8 additions per output pair.
// k_g1_stage3_mul: one lane quad per product; k_g1_stage3_comb: one lane quad per output pair (r, r + 4).
template <bool SPLIT>
__global__ void __launch_bounds__(32) k_g1_stage3_mul(const uint8_t* __restrict__ work, uint8_t* __restrict__ tmp, size_t n, int log_n, int s,
                                                      const uint8_t* __restrict__ roots, size_t nmax, int inverse) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const size_t q2 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const size_t q = SPLIT ? q2 >> 1 : q2;
    const int ghalf = SPLIT ? (int)(q2 & 1) : 0;
    const size_t nprod = 21 * (n >> 3);
    const bool live = q < nprod;
    const size_t qq = live ? q : 0;
    const size_t g = qq / 21;
    const int which = (int)(qq - g * 21);
    const size_t h = (size_t)1 << s, low = g & (h - 1), j = ((g >> s) << (s + 3)) | low;
    // exponents in units of the n-th root
    const size_t ea = low << (log_n - 1 - s), eb = low << (log_n - 2 - s), ec = low << (log_n - 3 - s);
    int m, r = 0;
    size_t e;
    if (which == 0) { m = 1; e = ea; }
    else if (which < 5) {                                  // 1: b x2   2: b' x2   3: ab x3   4: ab' x3
        const int odd = (which - 1) & 1;
        m = which < 3 ? 2 : 3;
        e = eb + (odd ? (n >> 2) : 0) + (m == 3 ? ea : 0);
    } else {                                               // 5 + 4 (m - 4) + r
        m = 4 + ((which - 5) >> 2);
        r = (which - 5) & 3;
        e = ec + (size_t)r * (n >> 3);
        if (m & 1) e += ea;
        if (m & 2) e += eb + ((r & 1) ? (n >> 2) : 0);
    }
    e &= n - 1;
    const uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    const int off = quad_store_offset();
    fp_t t = load_field<fp_t>(w + (j + (size_t)m * h) * 192 + off);
    const size_t eu = e * (nmax >> log_n);
    fr_t root = load_field_ro<fr_t>(roots + (inverse && eu ? nmax - eu : eu) * 32).from_mont();
    if (SPLIT) {
        t = quad_mul_scalar_half(t, root.v, table, ghalf);
        t = quad_add(t, shfl_xor_fp(t, 4));
    } else {
        t = quad_mul_scalar(t, root.v, table);
    }
    if (live && ghalf == 0) store_field(tmp + ((size_t)blockIdx.y * nprod + q) * 192 + off, t);
}
__global__ void __launch_bounds__(32) k_g1_stage3_comb(uint8_t* __restrict__ work, const uint8_t* __restrict__ tmp, size_t n, int s) {
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;   // four quads per group: both of a warp's groups in lockstep
    const bool live = q < (n >> 1);
    const size_t g = (live ? q : 0) >> 2;
    const int r = (int)(q & 3);
    const size_t h = (size_t)1 << s, low = g & (h - 1), j = ((g >> s) << (s + 3)) | low;
    uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    const uint8_t* pr = tmp + ((size_t)blockIdx.y * 21 * (n >> 3) + g * 21) * 192;
    const int off = quad_store_offset();
    const bool yrole = (threadIdx.x & 3) == 1;
    auto sgn = [&](const fp_t& v, bool minus) { return (minus && yrole) ? v.neg() : v; };
    auto P = [&](int k) { return load_field<fp_t>(pr + (size_t)k * 192 + off); };
    const bool in_minus = r & 1, out_minus = r >= 2;
    const fp_t x0 = load_field<fp_t>(w + j * 192 + off);
    const fp_t y = quad_add(x0, sgn(P(0), in_minus));
    const fp_t by = quad_add(P(1 + (r & 1)), sgn(P(3 + (r & 1)), in_minus));
    const fp_t u = quad_add(y, sgn(by, out_minus));
    const fp_t v45 = quad_add(P(5 + r), sgn(P(9 + r), in_minus));
    const fp_t v67 = quad_add(P(13 + r), sgn(P(17 + r), in_minus));
    const fp_t v = quad_add(v45, sgn(v67, out_minus));
    const fp_t z_lo = quad_add(u, v), z_hi = quad_add(u, sgn(v, true));
    // every quad of the warp has read x0 by now (the warp runs in lockstep through the additions above)
    __syncwarp();
    if (live) {
        store_field(w + (j + (size_t)r * h) * 192 + off, z_lo);
        store_field(w + (j + (size_t)(r + 4) * h) * 192 + off, z_hi);
    }
}
// R <= 6 DIT stages (s .. s+R-1) at the latency of one, for a LONE small transform (one blob's FK20 transforms), where
// even the fused triples leave the chain of scalar multiplications as the whole cost.  For the 2^R points x_m at j + m h:
//     z_r = x_0 + sum_{m >= 1} (-1)^popc(r & m) [ prod_{t in bits(m)} w_(2^(t+1) h)^(low + (r mod 2^t) h) ] x_m
// The bracket depends on r only through u = r mod 2^T, T = the top bit of m, so input m needs 2^T distinct products:
// (4^R - 1) / 3 per group in all (R = 6: 1365 instead of the 192 of six plain stages), every one at depth one.
// k_g1_stageR_mul: one lane quad (SPLIT: two) per product, product p -> level T, m = 2^T + (idx >> T), u = idx & (2^T - 1);
// the group's x_0 is copied behind its products so that the combination never reads what it overwrites.
// k_g1_stageR_comb: four quads per output r, a quarter of the 2^R terms each, then a two-level quad tree.
__device__ __forceinline__ size_t stageR_products(int R) { return (((size_t)1 << (2 * R)) - 1) / 3; }
template <bool SPLIT>
__global__ void __launch_bounds__(32) k_g1_stageR_mul(const uint8_t* __restrict__ work, uint8_t* __restrict__ tmp, size_t n, int log_n, int s, int R,
                                                      const uint8_t* __restrict__ roots, size_t nmax, int inverse) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const size_t q2 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const size_t q = SPLIT ? q2 >> 1 : q2;
    const int ghalf = SPLIT ? (int)(q2 & 1) : 0;
    const size_t NP = stageR_products(R), groups = n >> R, total = groups * (NP + 1);
    const bool live = q < total;
    const size_t qq = live ? q : 0;
    const size_t g = qq / (NP + 1), p = qq - g * (NP + 1);
    const size_t h = (size_t)1 << s, low = g & (h - 1), j = ((g >> s) << (s + R)) | low;
    const uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    uint8_t* dst = tmp + ((size_t)blockIdx.y * total + qq) * 192;
    const int off = quad_store_offset();
    if (__all_sync(kFullMask, p == NP)) {                      // (never a whole warp in practice; keeps the copy quads cheap)
        if (live && ghalf == 0) store_field(dst + off, load_field<fp_t>(w + j * 192 + off));
        return;
    }
    int T = 0;
    while (p < NP && p >= stageR_products(T + 1)) T++;
    const size_t idx = p < NP ? p - stageR_products(T) : 0;
    const size_t m = p < NP ? ((size_t)1 << T) + (idx >> T) : 0, u = idx & (((size_t)1 << T) - 1);
    size_t e = 0;
    for (int t = 0; t <= T; t++)
        if ((m >> t) & 1) e += (low + (u & (((size_t)1 << t) - 1)) * h) * (n >> (s + t + 1));
    e &= n - 1;
    fp_t t = load_field<fp_t>(w + (j + m * h) * 192 + off);
    const size_t eu = e * (nmax >> log_n);
    fr_t root = load_field_ro<fr_t>(roots + (inverse && eu ? nmax - eu : eu) * 32).from_mont();
    fp_t prod;
    if (SPLIT) {
        prod = quad_mul_scalar_half(t, root.v, table, ghalf);
        prod = quad_add(prod, shfl_xor_fp(prod, 4));
    } else {
        prod = quad_mul_scalar(t, root.v, table);
    }
    if (p == NP) prod = t;                                     // the x_0 copy (m = 0: t is x_0 itself)
    if (live && ghalf == 0) store_field(dst + off, prod);
}
__global__ void __launch_bounds__(32) k_g1_stageR_comb(uint8_t* __restrict__ work, const uint8_t* __restrict__ tmp, size_t n, int s, int R) {
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;   // four quads per output
    const size_t outs = n;                                                    // (n >> R) groups x 2^R outputs
    const bool live = (q >> 2) < outs;
    const size_t o = live ? q >> 2 : 0;
    const int part = (int)(q & 3);
    const size_t g = o >> R, r = o & (((size_t)1 << R) - 1);
    const size_t NP = stageR_products(R);
    const size_t h = (size_t)1 << s, low = g & (h - 1), j = ((g >> s) << (s + R)) | low;
    uint8_t* w = work + (size_t)blockIdx.y * n * 192;
    const uint8_t* pr = tmp + ((size_t)blockIdx.y * (n >> R) + g) * (NP + 1) * 192;
    const int off = quad_store_offset();
    const bool yrole = (threadIdx.x & 3) == 1;
    const size_t terms = (size_t)1 << R, per = terms >> 2 ? terms >> 2 : 1;
    fp_t acc = fp_t::zero();
#pragma unroll 1
    for (size_t m = part * per; m < (part + 1) * per && m < terms; m++) {
        fp_t v;
        if (m == 0) {
            v = load_field<fp_t>(pr + NP * 192 + off);
        } else {
            int T = 63 - __clzll((unsigned long long)m);
            const size_t p = stageR_products(T) + ((m - ((size_t)1 << T)) << T) + (r & (((size_t)1 << T) - 1));
            v = load_field<fp_t>(pr + p * 192 + off);
            if (yrole && (__popcll((unsigned long long)(r & m)) & 1)) v = v.neg();
        }
        acc = quad_add(acc, v);
    }
    acc = quad_tree(acc, 16);                                  // the four parts of an output are neighbouring quads
    if (live && part == 0) store_field(w + (j + r * h) * 192 + off, acc);
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

// One-time preparation for callers that will run many transforms of up to max_total points per launch (FK20): the work and
// product buffers at their final size -- growing them later means cudaFree, which synchronises the device under a running
// batch -- and every stage kernel variant loaded now instead of lazily inside a caller's first multi-blob pass.
void FFTSettingsDev::prepare_g1(size_t max_total) {
    if (max_total > g1_work_elems_) {
        cudaFree(g1_work_);
        g1_work_ = dev_alloc<uint8_t>(max_total * 192);
        g1_work_elems_ = max_total;
    }
    // the fused forms run up to 2^12 points per launch: 21/8 (triples, to 2^11) or 5/4 (pairs) products per point
    const size_t fused = std::min<size_t>(max_total, (size_t)1 << 12);
    size_t need = std::max<size_t>(21 * (std::min<size_t>(fused, (size_t)1 << 11) >> 3), 5 * (fused >> 2));
    need = std::max<size_t>(need, (std::min<size_t>(max_total, 256) >> 6) * 1366);   // six stages at once: 1365 products + x0 per 64 points
    if (need > g1_tmp_elems_) {
        cudaFree(g1_tmp_);
        g1_tmp_ = dev_alloc<uint8_t>(need * 192);
        g1_tmp_elems_ = need;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_g1_stage<true>);
    cudaFuncGetAttributes(&fa, k_g1_stage<false>);
    cudaFuncGetAttributes(&fa, k_g1_stage2_mul<true>);
    cudaFuncGetAttributes(&fa, k_g1_stage2_mul<false>);
    cudaFuncGetAttributes(&fa, k_g1_stage2_comb);
    cudaFuncGetAttributes(&fa, k_g1_stage3_mul<true>);
    cudaFuncGetAttributes(&fa, k_g1_stage3_mul<false>);
    cudaFuncGetAttributes(&fa, k_g1_stage3_comb);
    cudaFuncGetAttributes(&fa, k_g1_stageR_mul<true>);
    cudaFuncGetAttributes(&fa, k_g1_stageR_mul<false>);
    cudaFuncGetAttributes(&fa, k_g1_stageR_comb);
    cudaFuncGetAttributes(&fa, k_g1_brp_in);
    cudaFuncGetAttributes(&fa, k_g1_out);
    cudaGetLastError();
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
    // The transform is a chain of log n dependent scalar multiplications; a 128-point transform has only 64 butterflies per
    // stage.  While the launch is small the stages therefore run fused -- in triples (k_g1_stage3_*) or pairs (k_g1_stage2_*):
    // more products, all at the depth of one -- and stage by stage once the plain stages fill the machine (at 64 blobs x 128
    // points the fused products spill into a second wave of resident warps and the plain stages, 20 % less work, win:
    // 5.6 vs 6.5 ms, scripts/fft_g1_batch_timing.py).
    // B200_FFT_G1_FUSE (per call: tests toggle it): 0 plain stages, 2 fused pairs, 3 fused triples, unset = by launch size.
    // Triples (21 products per 8 points instead of 12) up to 2^11 points per launch -- 16 blobs' FK20 transforms, still one
    // wave of product quads: 128 points x 8 transforms 2.93 -> see scripts/fft_g1_batch_timing.py; pairs up to 2^12.
    const int fuse_env = getenv("B200_FFT_G1_FUSE") ? atoi(getenv("B200_FFT_G1_FUSE")) : -1;
    // 4 .. 6: R stages at once (k_g1_stageR_*), the default for launches of up to 256 points (one or two blobs' FK20 transforms: 1.01 / 1.28 ms per 128-point transform against 1.62 for split triples)
    const int fuse = fuse_env >= 0 ? (fuse_env == 1 ? 2 : fuse_env) : (total <= 256 && log_n >= 5) ? 6 : total <= ((size_t)1 << 11) ? 3
                                                                       : total <= ((size_t)1 << 12) ? 2 : 0;
    // B200_FFT_G1_SPLIT (per call): 1 / 0 force two quads / one quad per scalar multiplication; unset: two only while the
    // doubled launch leaves at most one warp per scheduler (592 x 8 quads) -- the halves do 1.6x the work of the whole, so they
    // pay off only where the chain, not the multiply pipe, is the limit (scripts/fft_g1_batch_timing.py: 128 points x 8
    // transforms, plain stages 5.57 -> 4.54 ms, pairs 2.93 -> 2.41; at 16 and more every split form loses to the unsplit one)
    const int split_env = getenv("B200_FFT_G1_SPLIT") ? atoi(getenv("B200_FFT_G1_SPLIT")) : -1;
    auto use_split = [&](size_t products) { return split_env >= 0 ? split_env != 0 : products * (size_t)batch * 2 <= (size_t)592 * 8; };
    auto plain = [&](int st_idx) {
        if (use_split(n / 2))
            k_g1_stage<true><<<dim3(div_up(n / 2 * 8, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, n, log_n, st_idx,
                                                                                      (const uint8_t*)roots_, max_width_, inverse);
        else
            k_g1_stage<false><<<dim3(div_up(n / 2 * 4, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, n, log_n, st_idx,
                                                                                       (const uint8_t*)roots_, max_width_, inverse);
        launches_++;
    };
    int s = 0;
    if (fuse >= 4 && fuse <= 6 && log_n >= 2) {
        const int R = std::min(fuse, std::max(2, log_n - 1));   // stage 0 multiplies by w^0 only: it stays a plain stage
        const size_t NP = (((size_t)1 << (2 * R)) - 1) / 3;
        const int lead = log_n % R;                            // leading stages one by one (stage 0 multiplies by w^0 only)
        for (; s < lead; s++) plain(s);
        const size_t per = (n >> R) * (NP + 1), need = (size_t)batch * per;
        if (need > g1_tmp_elems_) {
            cudaFree(g1_tmp_);
            g1_tmp_ = nullptr; g1_tmp_elems_ = 0;
            g1_tmp_ = dev_alloc<uint8_t>(need * 192);
            g1_tmp_elems_ = need;
        }
        for (; s + R <= log_n; s += R) {
            if (use_split(per))
                k_g1_stageR_mul<true><<<dim3(div_up(per * 8, 32), (unsigned)batch), 32, 0, st>>>((const uint8_t*)g1_work_, (uint8_t*)g1_tmp_, n, log_n, s,
                                                                                             R, (const uint8_t*)roots_, max_width_, inverse);
            else
                k_g1_stageR_mul<false><<<dim3(div_up(per * 4, 32), (unsigned)batch), 32, 0, st>>>((const uint8_t*)g1_work_, (uint8_t*)g1_tmp_, n, log_n, s,
                                                                                              R, (const uint8_t*)roots_, max_width_, inverse);
            k_g1_stageR_comb<<<dim3(div_up(n * 16, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, (const uint8_t*)g1_tmp_, n, s, R);
            launches_ += 2;
        }
    } else if (log_n > 0 && (fuse == 0 || (fuse == 2 && (log_n & 1)) || (fuse == 3 && log_n % 3 != 0))) {
        plain(0);                                              // stage 0 multiplies by w^0 only: cheap, and it fixes the parity
        s = 1;
    }
    if ((fuse == 2 || fuse == 3) && log_n >= 2) {
        const size_t need = (size_t)batch * (fuse == 3 ? 21 * (n >> 3) : 5 * (n >> 2));
        if (need > g1_tmp_elems_) {
            cudaFree(g1_tmp_);
            g1_tmp_ = nullptr; g1_tmp_elems_ = 0;
            g1_tmp_ = dev_alloc<uint8_t>(need * 192);
            g1_tmp_elems_ = need;
        }
        if (fuse == 3) {
            for (; s + 3 <= log_n; s += 3) {
                if (use_split(21 * (n >> 3)))
                    k_g1_stage3_mul<true><<<dim3(div_up(21 * (n >> 3) * 8, 32), (unsigned)batch), 32, 0, st>>>(
                        (const uint8_t*)g1_work_, (uint8_t*)g1_tmp_, n, log_n, s, (const uint8_t*)roots_, max_width_, inverse);
                else
                    k_g1_stage3_mul<false><<<dim3(div_up(21 * (n >> 3) * 4, 32), (unsigned)batch), 32, 0, st>>>(
                        (const uint8_t*)g1_work_, (uint8_t*)g1_tmp_, n, log_n, s, (const uint8_t*)roots_, max_width_, inverse);
                k_g1_stage3_comb<<<dim3(div_up((n >> 1) * 4, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, (const uint8_t*)g1_tmp_, n, s);
                launches_ += 2;
            }
        }
        for (; s + 2 <= log_n; s += 2) {
            if (use_split(5 * (n >> 2)))
                k_g1_stage2_mul<true><<<dim3(div_up(5 * (n >> 2) * 8, 32), (unsigned)batch), 32, 0, st>>>(
                    (const uint8_t*)g1_work_, (uint8_t*)g1_tmp_, n, log_n, s, (const uint8_t*)roots_, max_width_, inverse);
            else
                k_g1_stage2_mul<false><<<dim3(div_up(5 * (n >> 2) * 4, 32), (unsigned)batch), 32, 0, st>>>(
                    (const uint8_t*)g1_work_, (uint8_t*)g1_tmp_, n, log_n, s, (const uint8_t*)roots_, max_width_, inverse);
            k_g1_stage2_comb<<<dim3(div_up((n >> 2) * 4, 32), (unsigned)batch), 32, 0, st>>>((uint8_t*)g1_work_, (const uint8_t*)g1_tmp_, n, s);
            launches_ += 2;
        }
    }
    for (; s < log_n; s++) plain(s);
    const uint8_t* inv_n = (const uint8_t*)roots_ + (max_width_ + 1) * 32 + 33 * 32;
    k_g1_out<<<div_up(total * 4, 32), 32, 0, st>>>((const uint8_t*)g1_work_, (uint8_t*)out_jac_dev, total,
                                                 inverse && log_n && apply_scale ? inv_n + log_n * 32 : nullptr);
    launches_ += 2;
    B200_LAUNCH_CHECK();
}

}  // namespace b200
