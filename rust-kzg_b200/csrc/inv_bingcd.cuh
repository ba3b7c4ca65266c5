// inv_bingcd.cuh -- modular inversion by the optimised binary GCD of T. Pornin (eprint 2020/972), variable time.
//
// The plain binary extended Euclid (mont.cuh: Mont::inverse) walks the operands one bit at a time with four full-width
// carry chains per step: ~79 k dependent instructions for a 381-bit modulus, ~160 us for a lone warp on B200.  Here
// every ROUND works on 64-bit approximations of (a, b) -- the low 31 bits and the top 33 bits of the pair -- for 31
// binary-GCD steps, collecting the steps into a 2x2 matrix of small integers (f0 g0 / f1 g1), and only then touches
// the full-width numbers: (a, b) <- (a f0 + b g0, a f1 + b g1) / 2^31 (exact), and the same matrix on the Bezout
// coefficients (u, v) modulo m with one Montgomery-style word reduction.  ceil((2 * bits - 1) / 31) = 25 rounds at most
// for Fp, each ~600 instructions: ~5x fewer, and far shorter dependency chains.  The approximate comparison inside a
// round can make a value come out negative; it is negated together with its matrix row, as in the paper.
//
// Invariants (mod m), with y the input:  a = y * 2^i * u,  b = y * 2^i * v  after round i  (each round divides (a, b) by
// 2^31 and (u, v) by 2^32).  At the end a = 0, b = gcd = 1, so  y^-1 = 2^rounds * v.
//
// Plain C++ on 32-bit limbs: the same source is compiled for the device (inside Mont<>) and, by tests, for the host.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define B200_INV_FN __device__ __forceinline__
#define B200_INV_CLZ(x) __clz((int)(x))
#else
#define B200_INV_FN inline
#define B200_INV_CLZ(x) ((x) ? __builtin_clz(x) : 32)
#endif

namespace b200 {

// P: N (limbs), mod(i) (limb i of the odd modulus m), INV (-m^-1 mod 2^32).
// out = y^-1 mod m for 0 < y < m (plain residues, not Montgomery form); y = 0 gives 0.
template <class P>
B200_INV_FN void inverse_bingcd(uint32_t* out, const uint32_t* y) {
    constexpr int N = P::N;
    uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = y[i]; b[i] = P::mod(i); u[i] = 0; v[i] = 0; }
    u[0] = 1;
    int rounds = 0;
    for (;; rounds++) {
        // a == 0: done (b = gcd = 1 for 0 < y < m, m prime)
        uint32_t nz = 0, top_or = 0;
        int top = 0;                                   // index of the highest limb in which a | b is non-zero
#pragma unroll
        for (int i = 0; i < N; i++) {
            nz |= a[i];
            const uint32_t ab = a[i] | b[i];
            if (ab) { top = i; top_or = ab; }
        }
        if (nz == 0) break;
        // ---- 64-bit approximations: low 31 bits exact, plus the 33 bits below bit n = max bit length -----------------
        uint64_t xa, xb;
        const int n = 32 * top + 32 - B200_INV_CLZ(top_or);
        if (n <= 64) {
            xa = ((uint64_t)a[1] << 32) | a[0];
            xb = ((uint64_t)b[1] << 32) | b[0];
        } else {
            const int sh = n - 33, word = sh >> 5, bit = sh & 31;   // window [sh, sh + 33) spans at most 3 limbs
            uint32_t a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (i == word) { a0 = a[i]; b0 = b[i]; }
                if (i == word + 1) { a1 = a[i]; b1 = b[i]; }
                if (i == word + 2) { a2 = a[i]; b2 = b[i]; }
            }
            uint64_t ha = (((uint64_t)a1 << 32) | a0) >> bit, hb = (((uint64_t)b1 << 32) | b0) >> bit;
            if (bit) { ha |= (uint64_t)a2 << (64 - bit); hb |= (uint64_t)b2 << (64 - bit); }
            ha &= 0x1ffffffffull; hb &= 0x1ffffffffull;
            xa = (ha << 31) | (a[0] & 0x7fffffffu);
            xb = (hb << 31) | (b[0] & 0x7fffffffu);
        }
        // ---- 31 binary-GCD steps on the approximations; |f|, |g| <= 2^31 ------------------------------------------
        int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
        for (int i = 0; i < 31; i++) {
            if (xa & 1) {
                if (xa < xb) {
                    uint64_t t = xa; xa = xb; xb = t;
                    int64_t s = f0; f0 = f1; f1 = s;
                    s = g0; g0 = g1; g1 = s;
                }
                xa -= xb; f0 -= f1; g0 -= g1;
            }
            xa >>= 1; f1 <<= 1; g1 <<= 1;
        }
        // ---- (a, b) <- (a f0 + b g0, a f1 + b g1) / 2^31, exact; a negative result is negated with its matrix row ------
        uint32_t na[N], nb[N];
        bool neg_a, neg_b;
        {
            // t = a * f + b * g in two's complement on N + 2 limbs; products of magnitudes, signs applied per term
            auto lin = [&](int64_t f, int64_t g, uint32_t* r) -> bool {
                const bool sf = f < 0, sg = g < 0;
                const uint64_t mf = (uint64_t)(sf ? -f : f), mg = (uint64_t)(sg ? -g : g);   // <= 2^31
                uint32_t t[N + 2];
                uint64_t cf = 0, cg = 0;
                int64_t carry = 0;                                   // running signed carry of the limb-wise sum
#pragma unroll
                for (int i = 0; i < N; i++) {
                    cf += (uint64_t)a[i] * mf;
                    cg += (uint64_t)b[i] * mg;
                    const uint32_t pf = (uint32_t)cf, pg = (uint32_t)cg;
                    cf >>= 32; cg >>= 32;
                    carry += sf ? -(int64_t)pf : (int64_t)pf;
                    carry += sg ? -(int64_t)pg : (int64_t)pg;
                    t[i] = (uint32_t)carry;
                    carry >>= 32;                                    // arithmetic shift: floor division
                }
                carry += sf ? -(int64_t)cf : (int64_t)cf;
                carry += sg ? -(int64_t)cg : (int64_t)cg;
                t[N] = (uint32_t)carry;
                carry >>= 32;
                t[N + 1] = (uint32_t)carry;
                const bool negative = (t[N + 1] >> 31) != 0;
                if (negative) {                                      // two's complement negation
                    uint32_t c = 1;
#pragma unroll
                    for (int i = 0; i < N + 2; i++) {
                        const uint32_t x = ~t[i] + c;
                        c = (c && x == 0) ? 1u : 0u;
                        t[i] = x;
                    }
                }
                // >> 31 (the low 31 bits are zero by construction)
#pragma unroll
                for (int i = 0; i < N; i++) r[i] = (t[i] >> 31) | (t[i + 1] << 1);
                return negative;
            };
            neg_a = lin(f0, g0, na);
            neg_b = lin(f1, g1, nb);
        }
        if (neg_a) { f0 = -f0; g0 = -g0; }
        if (neg_b) { f1 = -f1; g1 = -g1; }
        // ---- (u, v) <- (u f0 + v g0, u f1 + v g1) / 2^32 mod m -------------------------------------------------------------
        uint32_t nu[N], nv[N];
        {
            auto linmod = [&](int64_t f, int64_t g, uint32_t* r) {
                const bool sf = f < 0, sg = g < 0;
                const uint64_t mf = (uint64_t)(sf ? -f : f), mg = (uint64_t)(sg ? -g : g);
                // negative coefficient: use m - x instead of x (x in [0, m]; m - 0 = m is fine, it is 0 mod m)
                uint32_t us[N], vs[N];
                {
                    uint32_t bu = 0, bv = 0;
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        const uint64_t du = (uint64_t)P::mod(i) - u[i] - bu, dv = (uint64_t)P::mod(i) - v[i] - bv;
                        us[i] = sf ? (uint32_t)du : u[i];
                        vs[i] = sg ? (uint32_t)dv : v[i];
                        bu = (uint32_t)(du >> 63); bv = (uint32_t)(dv >> 63);
                    }
                }
                // t = us * mf + vs * mg  (< 2 m 2^31: N + 1 limbs), then one word of Montgomery reduction: (t + q m) / 2^32
                uint32_t t[N + 1];
                uint64_t c = 0, c2 = 0;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    c += (uint64_t)us[i] * mf;
                    c2 += (uint64_t)vs[i] * mg;
                    const uint64_t s = (c & 0xffffffffu) + (c2 & 0xffffffffu);
                    t[i] = (uint32_t)s;
                    c = (c >> 32) + (s >> 32);
                    c2 >>= 32;
                }
                t[N] = (uint32_t)(c + c2);                           // < 2^32: t < 2^(32 N + 32)
                const uint32_t q = t[0] * P::INV;
                uint64_t k = (uint64_t)q * P::mod(0) + t[0];         // low word becomes zero
                k >>= 32;
#pragma unroll
                for (int i = 1; i < N; i++) {
                    k += (uint64_t)q * P::mod(i) + t[i];
                    r[i - 1] = (uint32_t)k;
                    k >>= 32;
                }
                k += t[N];
                r[N - 1] = (uint32_t)k;
                const uint32_t hi = (uint32_t)(k >> 32);              // value < 2 m + ... : one conditional subtraction suffices
                // r >= m (or hi set): subtract m
                uint32_t d[N];
                uint32_t bw = 0;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    const uint64_t x = (uint64_t)r[i] - P::mod(i) - bw;
                    d[i] = (uint32_t)x;
                    bw = (uint32_t)(x >> 63);
                }
                if (hi || !bw) {
#pragma unroll
                    for (int i = 0; i < N; i++) r[i] = d[i];
                }
            };
            linmod(f0, g0, nu);
            linmod(f1, g1, nv);
        }
#pragma unroll
        for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
    }
    // y^-1 = 2^rounds * v mod m  (b = 1); y = 0: a was 0 from the start, rounds = 0, v = 0 -> 0
#pragma unroll 1
    for (int k = 0; k < rounds; k++) {
        uint32_t c = 0, d[N], bw = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t x = (v[i] << 1) | c;
            c = v[i] >> 31;
            v[i] = x;
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint64_t x = (uint64_t)v[i] - P::mod(i) - bw;
            d[i] = (uint32_t)x;
            bw = (uint32_t)(x >> 63);
        }
        if (c || !bw) {
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = d[i];
        }
    }
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = v[i];
}

}  // namespace b200
