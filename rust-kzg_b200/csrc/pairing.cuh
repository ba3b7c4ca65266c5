// pairing.cuh -- BLS12-381 extension tower and optimal-ate pairing check on the device.
//
// Replaces, for the verify_* functions, the reference's pairings_verify (blst/src/kzg_proofs.rs:74-100: two
// Miller loops aggregated by blst, final_exp, blst_fp12_is_one).  Only the boolean "prod e(P_i, Q_i) == 1" is
// observable, so the layout of the computation is free; it is chosen for the machine:
//
//  * every Q_i the KZG verifiers ever pair with is a FIXED point of the trusted setup ([1]G2, [s]G2, [s^64]G2), so the
//    G2 side of the Miller loop (63 doublings + 5 additions in E'(Fp2), eprint 2010/354 Alg. 26/27) is done once at
//    load time and kept as 68 line-coefficient triples per point (19.1 KiB); a check only evaluates the lines at P_i.
//  * a single pairing is a serial chain of ~500 Fp12 multiplications, each 54+ Fp multiplications: one thread would
//    need ~15 ms.  Fp12 is therefore held in shared memory in the flat basis Fp2[w]/(w^6 - xi), xi = 1 + u,
//    (w^2 = v, so tower coefficient c_i.c_j is w^(i + 2j)) and ONE CTA OF FOUR WARPS computes a product cooperatively:
//    thread (k, e, q) forms one of the six terms that land on component e (re / im) of output coefficient k, so the four
//    multiply pipes of the SM work on the same product; twelve threads sum the terms, wrapped ones (i + j >= 6)
//    separately so that xi is applied once.  Critical path: 2 Fp multiplications instead of 54 (1 for the sparse line
//    multiplication, 1 for a cyclotomic squaring).
//  * inversion, Frobenius and the final exponentiation (the addition chain of zkcrypto/bls12_381/src/pairings.rs:
//    138-171, with plain squarings) are built from the same warp primitives.
//
// All w12_* functions must be called by all kPairThreads threads of the CTA with uniform arguments.
#pragma once
#include "g1.cuh"
#include "mont.cuh"
#include "pairing_consts.cuh"

namespace b200 {

typedef Mont<FpParams, MONT_CALL> pf_t;  // unrolled multiplier behind a call: full speed, small code

template <class A, class B>
__device__ __forceinline__ A fp_cast(const B& b) {
    A a;
#pragma unroll
    for (int i = 0; i < 12; i++) a.v[i] = b.v[i];
    return a;
}
__device__ __forceinline__ pf_t pf_const(const uint32_t* w) {
    pf_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = w[i];
    return r;
}
__device__ __forceinline__ pf_t pf_shfl_xor(const pf_t& a, int m) {
    pf_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], m);
    return r;
}

// ---- Fp2 = Fp[u]/(u^2 + 1), one thread (zkcrypto/bls12_381/src/fp2.rs:86-245 for the formulas' provenance) ----------
struct fp2_t {
    pf_t re, im;
    static __device__ __forceinline__ fp2_t zero() { return fp2_t{pf_t::zero(), pf_t::zero()}; }
    static __device__ __forceinline__ fp2_t one() { return fp2_t{pf_t::one(), pf_t::zero()}; }
    __device__ __forceinline__ bool is_zero() const { return re.is_zero() && im.is_zero(); }
    __device__ __forceinline__ bool operator==(const fp2_t& o) const { return re == o.re && im == o.im; }
    __device__ __forceinline__ fp2_t neg() const { return fp2_t{re.neg(), im.neg()}; }
    __device__ __forceinline__ fp2_t dbl() const { return fp2_t{re.dbl(), im.dbl()}; }
    __device__ __forceinline__ fp2_t conj() const { return fp2_t{re, im.neg()}; }
    __device__ __forceinline__ fp2_t mul_xi() const { return fp2_t{re - im, re + im}; }  // (1 + u)
    __device__ __forceinline__ fp2_t scale(const pf_t& s) const { return fp2_t{re * s, im * s}; }
};
__device__ __forceinline__ fp2_t operator+(const fp2_t& a, const fp2_t& b) { return fp2_t{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ fp2_t operator-(const fp2_t& a, const fp2_t& b) { return fp2_t{a.re - b.re, a.im - b.im}; }
static __device__ __noinline__ fp2_t fp2_mul(const fp2_t& a, const fp2_t& b) {  // Karatsuba, 3 Fp products
    pf_t t0 = a.re * b.re, t1 = a.im * b.im, t2 = (a.re + a.im) * (b.re + b.im);
    return fp2_t{t0 - t1, t2 - t0 - t1};
}
static __device__ __noinline__ fp2_t fp2_sqr(const fp2_t& a) {
    pf_t m = a.re * a.im;
    return fp2_t{(a.re + a.im) * (a.re - a.im), m.dbl()};
}
__device__ __forceinline__ fp2_t operator*(const fp2_t& a, const fp2_t& b) { return fp2_mul(a, b); }
static __device__ __noinline__ fp2_t fp2_inverse(const fp2_t& a) {  // conj(a) / (re^2 + im^2); 0 -> 0
    pf_t t = (a.re * a.re + a.im * a.im).inverse();
    return fp2_t{a.re * t, (a.im * t).neg()};
}
// a^e, e = 12 little-endian words
static __device__ __noinline__ fp2_t fp2_pow(const fp2_t& a, const uint32_t* e) {
    fp2_t acc = fp2_t::one();
    bool started = false;
    for (int i = 383; i >= 0; i--) {
        if (started) acc = fp2_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) {
            acc = started ? fp2_mul(acc, a) : a;
            started = true;
        }
    }
    return acc;
}
// words of (p - sub) >> sh
__device__ __forceinline__ void p_minus_shift(uint32_t* e, uint32_t sub, int sh) {
    uint32_t borrow = sub;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t t = FpParams::mod(i);
        e[i] = t - borrow;
        borrow = t < borrow ? 1u : 0u;
    }
#pragma unroll
    for (int i = 0; i < 12; i++) e[i] = (e[i] >> sh) | (i < 11 ? e[i + 1] << (32 - sh) : 0u);
}
// square root in Fp2 for p = 3 mod 4 (eprint 2012/685 Alg. 9, as zkcrypto/bls12_381/src/fp2.rs:246-296 uses it)
static __device__ __noinline__ bool fp2_sqrt(const fp2_t& a, fp2_t& out) {
    if (a.is_zero()) { out = a; return true; }
    uint32_t e[12];
    p_minus_shift(e, 3, 2);  // (p - 3) / 4
    fp2_t a1 = fp2_pow(a, e);
    fp2_t alpha = fp2_mul(fp2_sqr(a1), a);
    fp2_t x0 = fp2_mul(a1, a);
    fp2_t r;
    if (alpha == fp2_t::one().neg()) {
        r = fp2_t{x0.im.neg(), x0.re};  // u * x0
    } else {
        p_minus_shift(e, 1, 1);  // (p - 1) / 2
        fp2_t b = fp2_pow(alpha + fp2_t::one(), e);
        r = fp2_mul(b, x0);
    }
    out = r;
    return fp2_sqr(r) == a;
}
__device__ __forceinline__ bool pf_lex_largest(const pf_t& y) { return cc::fp_is_lex_largest(fp_cast<fpc_t>(y)); }
__device__ __forceinline__ bool fp2_lex_largest(const fp2_t& a) {  // fp2.rs:172-181
    return pf_lex_largest(a.im) || (a.im.is_zero() && pf_lex_largest(a.re));
}
__device__ __forceinline__ fp2_t load_fp2(const void* p) {
    return fp2_t{load_field<pf_t>(p), load_field<pf_t>((const uint8_t*)p + 48)};
}
__device__ __forceinline__ void store_fp2(void* p, const fp2_t& a) {
    store_field((uint8_t*)p, a.re);
    store_field((uint8_t*)p + 48, a.im);
}

// 48 big-endian bytes -> canonical limbs; false if >= p.  mask: clear the three flag bits of the first byte
__device__ __forceinline__ bool pf_from_be48(const uint8_t* in, bool mask, pf_t& out) {
    pf_t x;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const uint8_t* p = in + 4 * (11 - k);
        x.v[k] = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
    }
    if (mask) x.v[11] &= 0x1fffffffu;
    bool lt = false;
#pragma unroll
    for (int i = 11; i >= 0; i--) {
        uint32_t m = FpParams::mod(i);
        if (x.v[i] != m) { lt = x.v[i] < m; break; }
    }
    out = x.to_mont();
    return lt;
}
// G2 compressed decoding (blst_p2_uncompress via FsG2::from_bytes, blst/src/types/g2.rs:50-72; encoding
// zkcrypto/bls12_381/src/notes/serialization.rs, g2.rs:402-465): c1 of x first, on-curve check, no subgroup check.
// Infinity decodes to x = y = 0.
static __device__ __noinline__ bool g2_uncompress(const uint8_t* in, fp2_t& x, fp2_t& y) {
    uint32_t b0 = in[0];
    uint32_t cflag = b0 >> 7, iflag = (b0 >> 6) & 1, sflag = (b0 >> 5) & 1;
    x = fp2_t::zero();
    y = fp2_t::zero();
    if (!cflag) return false;
    fp2_t xx;
    if (!pf_from_be48(in, true, xx.im) || !pf_from_be48(in + 48, false, xx.re)) return false;
    if (iflag) return !sflag && xx.is_zero();
    pf_t four = pf_t::one().dbl().dbl();
    fp2_t y2 = fp2_mul(fp2_sqr(xx), xx) + fp2_t{four, four};  // b' = 4(1 + u)
    fp2_t yy;
    if (!fp2_sqrt(y2, yy)) return false;
    if (fp2_lex_largest(yy) != (bool)sflag) yy = yy.neg();
    x = xx;
    y = yy;
    return true;
}

// ---- G2 side of the Miller loop: line coefficients of a fixed Q ------------------------------------------------------
static constexpr int kMillerLines = 68;             // 63 doublings + 5 additions for |x| = 0xd201000000010000
static constexpr int kLineBytes = 3 * 96;           // three Fp2 per line
static constexpr uint64_t kBlsXHalf = 0xd201000000010000ull >> 1;
struct g2proj_t { fp2_t x, y, z; };
// Alg. 26 of eprint 2010/354 (the form zkcrypto/bls12_381/src/pairings.rs:708-737 evaluates): doubles r, returns the
// tangent line as (l0, l1, l2) with  line(P) = l2 + (l1 * P.x) w^2 + (l0 * P.y) w^3
static __device__ __noinline__ void g2_line_dbl(g2proj_t& r, fp2_t* l) {
    fp2_t a = fp2_sqr(r.x), b = fp2_sqr(r.y), c = fp2_sqr(b);
    fp2_t d = (fp2_sqr(b + r.x) - a - c).dbl();
    fp2_t e = a.dbl() + a;
    fp2_t g = r.x + e;
    fp2_t f = fp2_sqr(e);
    fp2_t zz = fp2_sqr(r.z);
    fp2_t x3 = f - d - d;
    fp2_t z3 = fp2_sqr(r.z + r.y) - b - zz;
    fp2_t y3 = fp2_mul(d - x3, e) - c.dbl().dbl().dbl();
    l[1] = fp2_mul(e, zz).dbl().neg();
    l[2] = fp2_sqr(g) - a - f - b.dbl().dbl();
    l[0] = fp2_mul(z3, zz).dbl();
    r.x = x3; r.y = y3; r.z = z3;
}
// Alg. 27: r += q (q affine), returns the chord line in the same form (pairings.rs:739-770)
static __device__ __noinline__ void g2_line_add(g2proj_t& r, const fp2_t& qx, const fp2_t& qy, fp2_t* l) {
    fp2_t zz = fp2_sqr(r.z), yy = fp2_sqr(qy);
    fp2_t t0 = fp2_mul(zz, qx);
    fp2_t t1 = fp2_mul(fp2_sqr(qy + r.z) - yy - zz, zz);
    fp2_t t2 = t0 - r.x;
    fp2_t t3 = fp2_sqr(t2);
    fp2_t t4 = t3.dbl().dbl();
    fp2_t t5 = fp2_mul(t4, t2);
    fp2_t t6 = t1 - r.y - r.y;
    fp2_t t9 = fp2_mul(t6, qx);
    fp2_t t7 = fp2_mul(t4, r.x);
    fp2_t x3 = fp2_sqr(t6) - t5 - t7 - t7;
    fp2_t z3 = fp2_sqr(r.z + t2) - zz - t3;
    fp2_t t10 = qy + z3;
    fp2_t t8 = fp2_mul(t7 - x3, t6);
    fp2_t y3 = t8 - fp2_mul(r.y, t5).dbl();
    t10 = fp2_sqr(t10) - yy - fp2_sqr(z3);
    l[2] = t9.dbl() - t10;
    l[0] = z3.dbl();
    l[1] = t6.neg().dbl();
    r.x = x3; r.y = y3; r.z = z3;
}
// all 68 lines of the Miller loop of Q = (qx, qy), in evaluation order (pairings.rs:667-693)
static __device__ __noinline__ void g2_prepare_lines(const fp2_t& qx, const fp2_t& qy, uint8_t* out) {
    g2proj_t r{qx, qy, fp2_t::one()};
    fp2_t l[3];
    int idx = 0;
    for (int b = 61; b >= 0; b--) {
        g2_line_dbl(r, l);
        for (int k = 0; k < 3; k++) store_fp2(out + (size_t)idx * kLineBytes + k * 96, l[k]);
        idx++;
        if ((kBlsXHalf >> b) & 1) {
            g2_line_add(r, qx, qy, l);
            for (int k = 0; k < 3; k++) store_fp2(out + (size_t)idx * kLineBytes + k * 96, l[k]);
            idx++;
        }
    }
    g2_line_dbl(r, l);
    for (int k = 0; k < 3; k++) store_fp2(out + (size_t)idx * kLineBytes + k * 96, l[k]);
}

// ---- Fp12 on one CTA of four warps -----------------------------------------------------------------------------------
// An element is 12 pf_t in shared (or global) memory: c[2k + e] = component e (0 = re, 1 = im) of the coefficient of w^k.
// Every w12_* function is called by ALL kPairThreads threads of the CTA with uniform arguments and ends with a barrier.
// A product is split over 72 threads -- thread (k, e, q) forms the q-th of the six terms of component e of output
// coefficient k, two Fp multiplications for a full product, one for the sparse line -- one per SM sub-partition's worth
// of warps, so the four multiply pipes of the SM work on one Fp12 product at the same time; twelve threads then sum the
// terms (wrapped ones, i + j >= 6, separately, so that xi = 1 + u is applied once).
static constexpr int kPairThreads = 128;
static constexpr int kW12Bytes = 12 * 48;
static constexpr int kW12Terms = 72;   // scratch pf_t for the terms of one product

// dst[k, e] = S + xi-part of T, with S / T the sums of the unwrapped / wrapped terms; wrapmask[k] bit q = term q wraps.
// Thread (k, e) sums its own six terms; the other component's wrapped sum comes from the partner lane by shuffle.
__device__ __forceinline__ void w12_reduce_terms(pf_t* dst, const pf_t* prod, const uint32_t* wrapmask) {
    const int t = threadIdx.x;
    if (t < 12) {
        const int k = t >> 1, e = t & 1;
        const uint32_t wm = wrapmask[k];
        const pf_t* own = prod + t * 6;
        pf_t S = pf_t::zero(), T = pf_t::zero();
#pragma unroll 1
        for (int q = 0; q < 6; q++) {
            pf_t v = own[q];
            if ((wm >> q) & 1) T = T + v; else S = S + v;     // wm is the same for both lanes of a coefficient
        }
        pf_t To;
#pragma unroll
        for (int i = 0; i < 12; i++) To.v[i] = __shfl_xor_sync(0xfffu, T.v[i], 1);
        // xi T = (T.re - T.im) + (T.re + T.im) u
        pf_t ST = S + T;
        dst[t] = e ? ST + To : ST - To;
    }
    __syncthreads();
}
// dst = a * b.  dst may alias a and/or b.  prod: kW12Terms scratch elements.
static __device__ __noinline__ void w12_mul(pf_t* dst, const pf_t* a, const pf_t* b, pf_t* prod) {
    const int t = threadIdx.x;
    if (t < kW12Terms) {
        const int ke = t / 6, i = t - ke * 6, k = ke >> 1, e = ke & 1;
        int j = k - i;
        if (j < 0) j += 6;
        // e = 0: a.re b.re - a.im b.im        e = 1: a.re b.im + a.im b.re
        pf_t p0 = a[2 * i] * b[2 * j + e];
        pf_t p1 = a[2 * i + 1] * b[2 * j + (e ^ 1)];
        prod[t] = e ? p0 + p1 : p0 - p1;
    }
    __syncthreads();
    const uint32_t wrap[6] = {0x3eu, 0x3cu, 0x38u, 0x30u, 0x20u, 0x00u};   // term i wraps iff i > k
    w12_reduce_terms(dst, prod, wrap);
}
// dst = a * (b0 + b2 w^2 + b3 w^3), sp = {b0.re, b0.im, b2.re, b2.im, b3.re, b3.im}: the Miller-loop line.
// Term q of (k, e): slot = q / 2 (b0, b2, b3), part = q % 2 -- a single signed Fp product.
static __device__ __noinline__ void w12_mul_sparse(pf_t* dst, const pf_t* a, const pf_t* sp, pf_t* prod) {
    const int t = threadIdx.x;
    if (t < kW12Terms) {
        const int ke = t / 6, q = t - ke * 6, k = ke >> 1, e = ke & 1;
        const int slot = q >> 1, part = q & 1;
        const int j = slot == 0 ? 0 : slot + 1;
        int i = k - j;
        if (i < 0) i += 6;
        // e = 0: + a.re b.re - a.im b.im        e = 1: + a.re b.im + a.im b.re
        pf_t p = a[2 * i + part] * sp[2 * slot + (e ^ part)];
        prod[t] = (e == 0 && part) ? p.neg() : p;
    }
    __syncthreads();
    // slots 1 (j = 2) and 2 (j = 3) wrap when k < j: bits (2,3) for k < 2, bits (4,5) for k < 3
    const uint32_t wrap[6] = {0x3cu, 0x3cu, 0x30u, 0x00u, 0x00u, 0x00u};
    w12_reduce_terms(dst, prod, wrap);
}
// dst = a^2 for a in the cyclotomic subgroup (after the easy part of the final exponentiation): Granger-Scott
// (eprint 2009/565, the form zkcrypto/bls12_381/src/pairings.rs:66-113 uses) in the flat basis.  With
// F(x, y) = (xi y^2 + x^2, 2 x y):  (A0, A1) = F(a0, a3), (B0, B1) = F(a1, a4), (C0, C1) = F(a2, a5) and
//   a0' = 3 A0 - 2 a0   a3' = 3 A1 + 2 a3   a2' = 3 B0 - 2 a2   a5' = 3 B1 + 2 a5   a4' = 3 C0 - 2 a4   a1' = 3 xi C1 + 2 a1.
// The nine Fp2 squarings (x^2, y^2, (x + y)^2 per pair) are 18 Fp products, one per thread; sq: 18 scratch pf_t.
static __device__ __noinline__ void w12_cyclotomic_sqr(pf_t* dst, const pf_t* a, pf_t* sq) {
    const int t = threadIdx.x;
    // Both phases are written branch-free over the participating lanes (operands chosen by selects, absent terms are
    // zero): a lone warp pays ~100 cycles of carry-chain latency per field addition, so divergent paths that each add a
    // few elements cost more than the multiplication in the middle.
    if (t < 18) {
        const int p = t / 6, which = (t >> 1) % 3, part = t & 1;
        const pf_t zero = pf_t::zero();
        pf_t xr = a[2 * p], xi_ = a[2 * p + 1], yr = a[2 * (p + 3)], yi = a[2 * (p + 3) + 1];
        pf_t re = (which == 1 ? zero : xr) + (which == 0 ? zero : yr);
        pf_t im = (which == 1 ? zero : xi_) + (which == 0 ? zero : yi);
        // (re + im u)^2 = (re + im)(re - im) + (re + re) im u
        pf_t opa = re + (part ? re : im);
        pf_t opb = part ? im : re - im;
        sq[t] = opa * opb;
    }
    __syncthreads();
    pf_t r;
    if (t < 12) {
        const int k = t >> 1, e = t & 1;
        // which F feeds coefficient k: k = 0, 3 <- pair 0; k = 2, 5 <- pair 1; k = 4, 1 <- pair 2
        const int pr = (k == 0 || k == 3) ? 0 : (k == 2 || k == 5) ? 1 : 2;
        const bool first = (k == 0 || k == 2 || k == 4);       // F's first component (xi y^2 + x^2), combined with "- 2 a_k"
        const pf_t* q = sq + 6 * pr;                           // x^2 (re, im), y^2 (re, im), (x + y)^2 (re, im)
        // v = P1 + P2 + P3 - N1 - N2 - N3 with
        //   first,  e = 0:  y2.re + x2.re - y2.im             first,  e = 1:  y2.re + y2.im + x2.im
        //   second, k != 1: s.e - x2.e - y2.e                 (2 x y, s = (x + y)^2)
        //   k = 1,  e = 0:  (s.re - x2.re - y2.re) - (s.im - x2.im - y2.im)      k = 1, e = 1: the sum of the two
        const pf_t zero = pf_t::zero();
        const bool k1 = k == 1;
        pf_t P1 = first ? q[2] : q[4 + (k1 ? 0 : e)];
        pf_t P2 = first ? q[e] : (k1 ? (e ? q[5] : q[1]) : zero);
        pf_t P3 = first ? (e ? q[3] : zero) : (k1 && !e ? q[3] : zero);
        pf_t N1 = first ? (e ? zero : q[3]) : q[k1 ? 0 : e];
        pf_t N2 = first ? zero : q[2 + (k1 ? 0 : e)];
        pf_t N3 = first ? zero : (k1 ? (e ? q[1] : q[5]) : zero);
        pf_t N4 = (k1 && e) ? q[3] : zero;
        pf_t v = ((P1 + P2) + P3) - ((N1 + N2) + (N3 + N4));
        // 3 v -/+ 2 a_k = v + 2 (v -/+ a_k)
        pf_t ak = a[t];
        pf_t w = v + (first ? ak.neg() : ak);
        r = v + w.dbl();
    }
    __syncthreads();                                           // dst may alias a
    if (t < 12) dst[t] = r;
    __syncthreads();
}
__device__ __forceinline__ void w12_copy(pf_t* dst, const pf_t* a) {
    const int t = threadIdx.x;
    pf_t v;
    if (t < 12) v = a[t];
    __syncthreads();
    if (t < 12) dst[t] = v;
    __syncthreads();
}
__device__ __forceinline__ void w12_set_one(pf_t* dst) {
    const int t = threadIdx.x;
    if (t < 12) dst[t] = t == 0 ? pf_t::one() : pf_t::zero();
    __syncthreads();
}
__device__ __forceinline__ bool w12_is_one(const pf_t* a) {
    const int t = threadIdx.x;
    bool ok = true;
    if (t < 12) ok = a[t] == (t == 0 ? pf_t::one() : pf_t::zero());
    return __syncthreads_and(ok) != 0;
}
// conjugation over Fp6 = the p^6 Frobenius: odd powers of w change sign
__device__ __forceinline__ void w12_conj(pf_t* dst, const pf_t* a) {
    const int t = threadIdx.x;
    pf_t v;
    if (t < 12) {
        v = a[t];
        if ((t >> 1) & 1) v = v.neg();
    }
    __syncthreads();
    if (t < 12) dst[t] = v;
    __syncthreads();
}
// dst = a^p: coefficient k becomes conj(a_k) * FROB_GAMMA[k]
static __device__ __noinline__ void w12_frobenius(pf_t* dst, const pf_t* a) {
    const int t = threadIdx.x;
    pf_t c;
    if (t < 12) {
        const int k = t >> 1, e = t & 1;
        pf_t gre = pf_const(FROB_GAMMA[k][0]), gim = pf_const(FROB_GAMMA[k][1]);
        // (are - aim u)(gre + gim u) = are gre + aim gim + (are gim - aim gre) u
        pf_t p0 = a[2 * k] * (e ? gim : gre);
        pf_t p1 = a[2 * k + 1] * (e ? gre : gim);
        c = e ? p0 - p1 : p0 + p1;
    }
    __syncthreads();
    if (t < 12) dst[t] = c;
    __syncthreads();
}
// dst = 1 / f.  tmp: 4 scratch elements + kW12Terms.  With g = f conj(f) in Fp6 and N = g g^(p^2) g^(p^4) in Fp2:
// 1/f = conj(f) g^(p^2) g^(p^4) / N, one Fp inversion.
static __device__ __noinline__ void w12_inverse(pf_t* dst, const pf_t* f, pf_t* tmp) {
    pf_t *t1 = tmp, *t2 = tmp + 12, *t3 = tmp + 24, *t4 = tmp + 36, *prod = tmp + 48;
    w12_conj(t1, f);
    w12_mul(t2, f, t1, prod);                        // g
    w12_frobenius(t3, t2); w12_frobenius(t3, t3);    // g^(p^2)
    w12_frobenius(t4, t3); w12_frobenius(t4, t4);    // g^(p^4)
    w12_mul(t3, t3, t4, prod);                       // h
    w12_mul(t2, t2, t3, prod);                       // N: only the w^0 coefficient is non-zero
    if (threadIdx.x == 0) {
        fp2_t ni = fp2_inverse(fp2_t{t2[0], t2[1]});
        t2[0] = ni.re;
        t2[1] = ni.im;
    } else if (threadIdx.x < 12 && threadIdx.x >= 2) {
        t2[threadIdx.x] = pf_t::zero();
    }
    __syncthreads();
    w12_mul(t3, t3, t2, prod);
    w12_mul(dst, t1, t3, prod);
}
// dst = conj(f^|x|) = f^x for the (negative) BLS parameter; f in the cyclotomic subgroup.
// tmp: 1 scratch element + kW12Terms
static __device__ __noinline__ void w12_exp_x(pf_t* dst, const pf_t* f, pf_t* tmp) {
    const uint64_t X = 0xd201000000010000ull;
    pf_t* prod = tmp + 12;
    w12_copy(tmp, f);
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        w12_cyclotomic_sqr(tmp, tmp, prod);
        if ((X >> b) & 1) w12_mul(tmp, tmp, f, prod);
    }
    w12_conj(dst, tmp);
}
// f <- f^((p^12 - 1)/r) (up to the fixed cofactor of the chain in zkcrypto/bls12_381/src/pairings.rs:138-171).
// ws: 7 + 4 scratch elements + kW12Terms.
static constexpr int kW12FinalExpScratch = 11 * 12 + kW12Terms;
static __device__ __noinline__ void w12_final_exp(pf_t* f, pf_t* ws) {
    pf_t *t0 = ws, *t1 = ws + 12, *t2 = ws + 24, *t3 = ws + 36, *t4 = ws + 48, *t5 = ws + 60, *t6 = ws + 72, *sc = ws + 84;
    pf_t* prod = ws + 132;
    // easy part: f^((p^6 - 1)(p^2 + 1))
    w12_conj(t0, f);
    w12_inverse(t1, f, sc);
    w12_mul(t2, t0, t1, prod);
    w12_copy(t1, t2);
    w12_frobenius(t2, t2); w12_frobenius(t2, t2);
    w12_mul(t2, t2, t1, prod);
    // hard part
    w12_cyclotomic_sqr(t1, t2, prod); w12_conj(t1, t1);
    w12_exp_x(t3, t2, sc);
    w12_cyclotomic_sqr(t4, t3, prod);
    w12_mul(t5, t1, t3, prod);
    w12_exp_x(t1, t5, sc);
    w12_exp_x(t0, t1, sc);
    w12_exp_x(t6, t0, sc);
    w12_mul(t6, t6, t4, prod);
    w12_exp_x(t4, t6, sc);
    w12_conj(t5, t5);
    w12_mul(t5, t5, t2, prod); w12_mul(t4, t4, t5, prod);
    w12_conj(t5, t2);
    w12_mul(t1, t1, t2, prod);
    w12_frobenius(t1, t1); w12_frobenius(t1, t1); w12_frobenius(t1, t1);
    w12_mul(t6, t6, t5, prod);
    w12_frobenius(t6, t6);
    w12_mul(t3, t3, t0, prod);
    w12_frobenius(t3, t3); w12_frobenius(t3, t3);
    w12_mul(t3, t3, t1, prod);
    w12_mul(t3, t3, t6, prod);
    w12_mul(f, t3, t4, prod);
}

}  // namespace b200
