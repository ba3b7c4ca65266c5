/*
 * kzg_oracle.c -- CPU oracle for the rust-kzg hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's CPU algorithm for the path this repo accelerates (Pippenger MSM over
 * BLS12-381 G1, radix-2 Fr NTT + DAS extension, and the EIP-4844 commitment / proof functions that call them).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and only
 * as the checker / reported CPU baseline.  The product library (rust-kzg_b200/csrc) never links or calls it.
 *
 * The reference (Rust) cannot be built in this environment and its arithmetic lives in the un-vendored crate
 * blst 0.3.16 (Cargo.lock:530-532).  What is restated: the published BLS12-381 field/curve arithmetic in
 * Montgomery form with blst's memory layouts (kzg/src/eth/c_bindings.rs:427-474) and the reference's own
 * algorithms; each function cites the file:line under /root/reference it follows.
 * Parity is PINNED: tests/test_oracle_golden.py checks this library against the reference's consensus-spec
 * vectors and KATs in tests/golden/ (all valid + invalid cases of blob_to_kzg_commitment, compute_kzg_proof,
 * compute_blob_kzg_proof, compute_challenge, compute_cells; fft/das/compress KAT tables).
 *
 * Not blst's assembly: expect 1.5-3x slower per core than the reference's published blst numbers.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -shared -fPIC -pthread).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <ctype.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* blst memory layouts (kzg/src/eth/c_bindings.rs:427-474): little-endian u64 limbs, Montgomery form */
typedef struct { u64 l[4]; } fr_t;
typedef struct { u64 l[6]; } fp_t;
typedef struct { fp_t x, y, z; } p1_t;              /* Jacobian, infinity <=> z == 0 */
typedef struct { fp_t x, y; } p1_affine_t;          /* infinity <=> all zero (blst/src/types/g1.rs:303-316) */
typedef struct { fp_t x, y, zzz, zz; } p1xyzz_t;    /* kzg/src/msm/pippenger_utils.rs:5-12 */

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------ */
/* moduli (zkcrypto/bls12_381/src/fp.rs:71, scalar.rs:77).  Every other constant is derived in ko_init().      */
static const u64 FP_P[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                            0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 FR_R[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                            0x73eda753299d7d48ULL};
static u64 FP_INV, FR_INV;
static fp_t FP_ONE, FP_RR, FP_B;    /* R mod p, R^2 mod p, curve b = 4 in Montgomery form */
static fr_t FR_ONE, FR_RR;
static int g_inited = 0;

/* ---- generic N-limb helpers ---- */
static inline int limbs_ge(const u64 *a, const u64 *b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline u64 limbs_sub(u64 *r, const u64 *a, const u64 *b, int n) {
    u64 borrow = 0;
    for (int i = 0; i < n; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
static inline u64 limbs_add(u64 *r, const u64 *a, const u64 *b, int n) {
    u64 carry = 0;
    for (int i = 0; i < n; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        r[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    return carry;
}
static inline int limbs_is_zero(const u64 *a, int n) {
    u64 acc = 0;
    for (int i = 0; i < n; i++) acc |= a[i];
    return acc == 0;
}
static inline void mod_add(u64 *r, const u64 *a, const u64 *b, const u64 *m, int n) {
    u64 t[6];
    u64 c = limbs_add(t, a, b, n);
    if (c || limbs_ge(t, m, n)) limbs_sub(t, t, m, n);
    memcpy(r, t, 8 * n);
}
static inline void mod_sub(u64 *r, const u64 *a, const u64 *b, const u64 *m, int n) {
    u64 t[6];
    if (limbs_sub(t, a, b, n)) limbs_add(t, t, m, n);
    memcpy(r, t, 8 * n);
}
/* Montgomery multiplication, coarsely integrated operand scanning (the textbook CIOS form) */
static inline __attribute__((always_inline)) void mont_mul(u64 *r, const u64 *a, const u64 *b, const u64 *m,
                                                           u64 inv, const int n) {
    u64 t[8] = {0};
    for (int i = 0; i < n; i++) {
        u128 c = 0;
        for (int j = 0; j < n; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[n];
        t[n] = (u64)c;
        t[n + 1] = (u64)(c >> 64);
        u64 q = t[0] * inv;
        c = (u128)q * m[0] + t[0];
        c >>= 64;
        for (int j = 1; j < n; j++) {
            c += (u128)q * m[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[n];
        t[n - 1] = (u64)c;
        t[n] = t[n + 1] + (u64)(c >> 64);
    }
    if (t[n] || limbs_ge(t, m, n)) limbs_sub(t, t, m, n);
    memcpy(r, t, 8 * n);
}

/* ---- Fp ---- */
static inline void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) { mont_mul(r->l, a->l, b->l, FP_P, FP_INV, 6); }
static inline void fp_sqr(fp_t *r, const fp_t *a) { mont_mul(r->l, a->l, a->l, FP_P, FP_INV, 6); }
static inline void fp_add(fp_t *r, const fp_t *a, const fp_t *b) { mod_add(r->l, a->l, b->l, FP_P, 6); }
static inline void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) { mod_sub(r->l, a->l, b->l, FP_P, 6); }
static inline int fp_is_zero(const fp_t *a) { return limbs_is_zero(a->l, 6); }
static inline int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a, b, sizeof(fp_t)) == 0; }
static inline void fp_neg(fp_t *r, const fp_t *a) {
    if (fp_is_zero(a)) { *r = *a; return; }
    limbs_sub(r->l, FP_P, a->l, 6);
}
static void fp_pow(fp_t *r, const fp_t *a, const u64 *e, int elimbs) {
    fp_t acc = FP_ONE, base = *a;
    int started = 0;
    for (int i = elimbs * 64 - 1; i >= 0; i--) {
        if (started) fp_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) {
            if (started) fp_mul(&acc, &acc, &base); else { acc = base; started = 1; }
        }
    }
    *r = acc;
}
static void fp_inv(fp_t *r, const fp_t *a) {  /* a^(p-2) */
    u64 e[6];
    u64 two[6] = {2, 0, 0, 0, 0, 0};
    limbs_sub(e, FP_P, two, 6);
    fp_pow(r, a, e, 6);
}
static int fp_sqrt(fp_t *r, const fp_t *a) {  /* p = 3 mod 4: a^((p+1)/4); returns 1 if a is a square */
    u64 e[6], one[6] = {1, 0, 0, 0, 0, 0};
    limbs_add(e, FP_P, one, 6);
    for (int i = 0; i < 6; i++) e[i] = (e[i] >> 2) | (i < 5 ? e[i + 1] << 62 : 0);
    fp_t s, chk;
    fp_pow(&s, a, e, 6);
    fp_sqr(&chk, &s);
    *r = s;
    return fp_eq(&chk, a);
}
static void fp_from_canon(fp_t *r, const u64 c[6]) { fp_t t; memcpy(t.l, c, 48); fp_mul(r, &t, &FP_RR); }
static void fp_to_canon(u64 c[6], const fp_t *a) {
    fp_t one = {{1, 0, 0, 0, 0, 0}}, t;
    fp_mul(&t, a, &one);
    memcpy(c, t.l, 48);
}

/* ---- Fr ---- */
static inline void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) { mont_mul(r->l, a->l, b->l, FR_R, FR_INV, 4); }
static inline void fr_add(fr_t *r, const fr_t *a, const fr_t *b) { mod_add(r->l, a->l, b->l, FR_R, 4); }
static inline void fr_sub(fr_t *r, const fr_t *a, const fr_t *b) { mod_sub(r->l, a->l, b->l, FR_R, 4); }
static inline int fr_is_zero(const fr_t *a) { return limbs_is_zero(a->l, 4); }
static inline int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a, b, sizeof(fr_t)) == 0; }
static void fr_pow(fr_t *r, const fr_t *a, const u64 *e, int elimbs) {
    fr_t acc = FR_ONE, base = *a;
    for (int i = elimbs * 64 - 1; i >= 0; i--) {
        fr_mul(&acc, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) fr_mul(&acc, &acc, &base);
    }
    *r = acc;
}
static void fr_inv(fr_t *r, const fr_t *a) {  /* a^(r-2); 0 -> 0 like blst_fr_eucl_inverse */
    u64 e[4], two[4] = {2, 0, 0, 0};
    limbs_sub(e, FR_R, two, 4);
    fr_pow(r, a, e, 4);
}
static void fr_from_canon(fr_t *r, const u64 c[4]) { fr_t t; memcpy(t.l, c, 32); fr_mul(r, &t, &FR_RR); }
static void fr_to_canon(u64 c[4], const fr_t *a) {
    fr_t one = {{1, 0, 0, 0}}, t;
    fr_mul(&t, a, &one);
    memcpy(c, t.l, 32);
}
static void fr_from_u64(fr_t *r, u64 v) { u64 c[4] = {v, 0, 0, 0}; fr_from_canon(r, c); }

static u64 neg_inv64(u64 m0) {  /* -m0^{-1} mod 2^64 by Newton iteration */
    u64 x = 1;
    for (int i = 0; i < 7; i++) x *= 2 - m0 * x;
    return (u64)0 - x;
}

API void ko_init(void) {
    if (g_inited) return;
    FP_INV = neg_inv64(FP_P[0]);
    FR_INV = neg_inv64(FR_R[0]);
    /* R mod p by doubling 1 (384 times), R^2 by 768 doublings; same for r */
    u64 t[6] = {1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 768; i++) {
        mod_add(t, t, t, FP_P, 6);
        if (i == 383) memcpy(FP_ONE.l, t, 48);
    }
    memcpy(FP_RR.l, t, 48);
    u64 s[4] = {1, 0, 0, 0};
    for (int i = 0; i < 512; i++) {
        mod_add(s, s, s, FR_R, 4);
        if (i == 255) memcpy(FR_ONE.l, s, 32);
    }
    memcpy(FR_RR.l, s, 32);
    u64 four[6] = {4, 0, 0, 0, 0, 0};
    fp_from_canon(&FP_B, four);
    g_inited = 1;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* wire formats */
/* Fr::from_bytes (blst/src/types/fr.rs:64-86): big-endian, must be < r.  returns 0 ok, 1 invalid */
API int ko_fr_from_bendian(fr_t *out, const uint8_t in[32]) {
    u64 c[4];
    for (int i = 0; i < 4; i++) {
        u64 v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | in[(3 - i) * 8 + j];
        c[i] = v;
    }
    if (limbs_ge(c, FR_R, 4)) return 1;
    fr_from_canon(out, c);
    return 0;
}
/* Fr::from_bytes_unchecked (fr.rs:88-107): reduces mod r */
API void ko_fr_from_bendian_unchecked(fr_t *out, const uint8_t in[32]) {
    u64 c[4];
    for (int i = 0; i < 4; i++) {
        u64 v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | in[(3 - i) * 8 + j];
        c[i] = v;
    }
    fr_from_canon(out, c); /* Montgomery mul by RR reduces any 256-bit input */
}
/* Fr::to_bytes (fr.rs:127-136) */
API void ko_fr_to_bendian(uint8_t out[32], const fr_t *a) {
    u64 c[4];
    fr_to_canon(c, a);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) out[(3 - i) * 8 + j] = (uint8_t)(c[i] >> (8 * (7 - j)));
}
/* Fr::to_scalar (fr.rs:271-277): canonical little-endian limbs */
API void ko_fr_to_scalar(u64 out[4], const fr_t *a) { fr_to_canon(out, a); }
API void ko_fr_from_u64_arr(fr_t *out, const u64 in[4]) { fr_from_canon(out, in); }

API void ko_fr_mul_batch(fr_t *r, const fr_t *a, const fr_t *b, size_t n) { for (size_t i = 0; i < n; i++) fr_mul(r + i, a + i, b + i); }
API void ko_fr_add_batch(fr_t *r, const fr_t *a, const fr_t *b, size_t n) { for (size_t i = 0; i < n; i++) fr_add(r + i, a + i, b + i); }
API void ko_fr_sub_batch(fr_t *r, const fr_t *a, const fr_t *b, size_t n) { for (size_t i = 0; i < n; i++) fr_sub(r + i, a + i, b + i); }
API void ko_fr_inv_batch(fr_t *r, const fr_t *a, size_t n) { for (size_t i = 0; i < n; i++) fr_inv(r + i, a + i); }
API void ko_fp_mul_batch(fp_t *r, const fp_t *a, const fp_t *b, size_t n) { for (size_t i = 0; i < n; i++) fp_mul(r + i, a + i, b + i); }
API void ko_fp_add_batch(fp_t *r, const fp_t *a, const fp_t *b, size_t n) { for (size_t i = 0; i < n; i++) fp_add(r + i, a + i, b + i); }
API void ko_fp_sub_batch(fp_t *r, const fp_t *a, const fp_t *b, size_t n) { for (size_t i = 0; i < n; i++) fp_sub(r + i, a + i, b + i); }
API void ko_fp_inv_batch(fp_t *r, const fp_t *a, size_t n) { for (size_t i = 0; i < n; i++) fp_inv(r + i, a + i); }
API void ko_fp_from_canon(fp_t *r, const u64 c[6]) { fp_from_canon(r, c); }
API void ko_fp_to_canon(u64 c[6], const fp_t *a) { fp_to_canon(c, a); }

/* ------------------------------------------------------------------------------------------------------------ */
/* G1, Jacobian (blst_p1_* call sites: blst/src/types/g1.rs:10-15, 102-195) */
static inline int p1_is_inf(const p1_t *p) { return fp_is_zero(&p->z); }
static void p1_set_inf(p1_t *p) { memset(p, 0, sizeof(*p)); }

static void p1_double(p1_t *r, const p1_t *p) {  /* dbl-2009-l, a = 0 */
    if (p1_is_inf(p) || fp_is_zero(&p->y)) { p1_set_inf(r); return; }
    fp_t A, B, C, D, E, F, t;
    fp_sqr(&A, &p->x);
    fp_sqr(&B, &p->y);
    fp_sqr(&C, &B);
    fp_add(&t, &p->x, &B); fp_sqr(&t, &t); fp_sub(&t, &t, &A); fp_sub(&t, &t, &C); fp_add(&D, &t, &t);
    fp_add(&E, &A, &A); fp_add(&E, &E, &A);
    fp_sqr(&F, &E);
    fp_t z3; fp_mul(&z3, &p->y, &p->z); fp_add(&z3, &z3, &z3);
    fp_sub(&r->x, &F, &D); fp_sub(&r->x, &r->x, &D);
    fp_sub(&t, &D, &r->x); fp_mul(&t, &E, &t);
    fp_add(&C, &C, &C); fp_add(&C, &C, &C); fp_add(&C, &C, &C);
    fp_sub(&r->y, &t, &C);
    r->z = z3;
}
/* blst_p1_add_or_double semantics: handles infinity operands, P+P and P+(-P) */
static void p1_add_or_double(p1_t *r, const p1_t *p, const p1_t *q) {
    if (p1_is_inf(p)) { *r = *q; return; }
    if (p1_is_inf(q)) { *r = *p; return; }
    fp_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    fp_sqr(&z1z1, &p->z); fp_sqr(&z2z2, &q->z);
    fp_mul(&u1, &p->x, &z2z2); fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&s1, &p->y, &q->z); fp_mul(&s1, &s1, &z2z2);
    fp_mul(&s2, &q->y, &p->z); fp_mul(&s2, &s2, &z1z1);
    if (fp_eq(&u1, &u2)) {
        if (fp_eq(&s1, &s2)) { p1_t d; p1_double(&d, p); *r = d; } else p1_set_inf(r);
        return;
    }
    fp_sub(&h, &u2, &u1);
    fp_add(&i, &h, &h); fp_sqr(&i, &i);
    fp_mul(&j, &h, &i);
    fp_sub(&rr, &s2, &s1); fp_add(&rr, &rr, &rr);
    fp_mul(&v, &u1, &i);
    p1_t o;
    fp_sqr(&o.x, &rr); fp_sub(&o.x, &o.x, &j); fp_sub(&o.x, &o.x, &v); fp_sub(&o.x, &o.x, &v);
    fp_sub(&t, &v, &o.x); fp_mul(&t, &rr, &t);
    fp_mul(&s1, &s1, &j); fp_add(&s1, &s1, &s1);
    fp_sub(&o.y, &t, &s1);
    fp_add(&t, &p->z, &q->z); fp_sqr(&t, &t); fp_sub(&t, &t, &z1z1); fp_sub(&t, &t, &z2z2);
    fp_mul(&o.z, &t, &h);
    *r = o;
}
static void p1_from_affine(p1_t *r, const p1_affine_t *a) {
    if (fp_is_zero(&a->x) && fp_is_zero(&a->y)) { p1_set_inf(r); return; }
    r->x = a->x; r->y = a->y; r->z = FP_ONE;
}
static void p1_to_affine(p1_affine_t *r, const p1_t *p) {
    if (p1_is_inf(p)) { memset(r, 0, sizeof(*r)); return; }
    fp_t zi, zi2;
    fp_inv(&zi, &p->z); fp_sqr(&zi2, &zi);
    fp_mul(&r->x, &p->x, &zi2);
    fp_mul(&zi2, &zi2, &zi);
    fp_mul(&r->y, &p->y, &zi2);
}
/* G1Mul::mul (blst/src/types/g1.rs:242-273) -- scalar is an Fr in Montgomery form */
static void p1_mult_canon(p1_t *r, const p1_t *p, const u64 *k, int nbits) {
    p1_t acc; p1_set_inf(&acc);
    for (int i = nbits - 1; i >= 0; i--) {
        p1_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) p1_add_or_double(&acc, &acc, p);
    }
    *r = acc;
}
API void ko_p1_mult(p1_t *r, const p1_t *p, const fr_t *s) {
    u64 k[4]; fr_to_canon(k, s);
    p1_mult_canon(r, p, k, 255);
}
API void ko_p1_add_or_double(p1_t *r, const p1_t *a, const p1_t *b) { p1_add_or_double(r, a, b); }
API void ko_p1_double(p1_t *r, const p1_t *a) { p1_double(r, a); }
API void ko_p1_from_affine(p1_t *r, const p1_affine_t *a) { p1_from_affine(r, a); }
API void ko_p1_to_affine(p1_affine_t *r, const p1_t *p) { p1_to_affine(r, p); }
API int ko_p1_is_inf(const p1_t *p) { return p1_is_inf(p); }
API int ko_p1_is_equal(const p1_t *a, const p1_t *b) {
    p1_affine_t x, y; p1_to_affine(&x, a); p1_to_affine(&y, b);
    return memcmp(&x, &y, sizeof(x)) == 0 && p1_is_inf(a) == p1_is_inf(b);
}
/* blst_p1_in_g1 (g1.rs:110-119): [r]P == inf */
API int ko_p1_in_g1(const p1_t *p) {
    if (p1_is_inf(p)) return 1;
    p1_t t; p1_mult_canon(&t, p, FR_R, 255);
    return p1_is_inf(&t);
}
/* batch Jacobian -> affine with one inversion (FsG1Affine::into_affines, g1.rs:326-345) */
API void ko_p1s_to_affine(p1_affine_t *out, const p1_t *in, size_t n) {
    if (n == 0) return;
    fp_t *pref = (fp_t *)malloc(n * sizeof(fp_t));
    fp_t acc = FP_ONE;
    for (size_t i = 0; i < n; i++) {
        pref[i] = acc;
        if (!p1_is_inf(&in[i])) fp_mul(&acc, &acc, &in[i].z);
    }
    fp_t inv; fp_inv(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (p1_is_inf(&in[i])) { memset(&out[i], 0, sizeof(out[i])); continue; }
        fp_t zi, zi2;
        fp_mul(&zi, &inv, &pref[i]);
        fp_mul(&inv, &inv, &in[i].z);
        fp_sqr(&zi2, &zi);
        fp_mul(&out[i].x, &in[i].x, &zi2);
        fp_mul(&zi2, &zi2, &zi);
        fp_mul(&out[i].y, &in[i].y, &zi2);
    }
    free(pref);
}

/* compressed encoding: zkcrypto/bls12_381/src/notes/serialization.rs:1-29; blst_p1_compress / _uncompress via
 * blst/src/types/g1.rs:65-100 */
static int fp_is_lex_largest(const fp_t *y) {  /* y > (p-1)/2 */
    u64 c[6], half[6];
    fp_to_canon(c, y);
    for (int i = 0; i < 6; i++) half[i] = (FP_P[i] >> 1) | (i < 5 ? FP_P[i + 1] << 63 : 0);   /* (p-1)/2 */
    for (int i = 5; i >= 0; i--) {
        if (c[i] > half[i]) return 1;
        if (c[i] < half[i]) return 0;
    }
    return 0;
}
API void ko_p1_compress(uint8_t out[48], const p1_t *p) {
    if (p1_is_inf(p)) { memset(out, 0, 48); out[0] = 0xC0; return; }
    p1_affine_t a; p1_to_affine(&a, p);
    u64 c[6]; fp_to_canon(c, &a.x);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 8; j++) out[(5 - i) * 8 + j] = (uint8_t)(c[i] >> (8 * (7 - j)));
    out[0] |= 0x80;
    if (fp_is_lex_largest(&a.y)) out[0] |= 0x20;
}
/* returns 0 ok; 1 = malformed (FsG1::from_bytes -> Err "Failed to uncompress").  No subgroup check. */
API int ko_p1_uncompress(p1_affine_t *out, const uint8_t in[48]) {
    int cf = in[0] >> 7 & 1, inf = in[0] >> 6 & 1, sf = in[0] >> 5 & 1;
    if (!cf) return 1;
    u64 c[6];
    for (int i = 0; i < 6; i++) {
        u64 v = 0;
        for (int j = 0; j < 8; j++) {
            uint8_t b = in[(5 - i) * 8 + j];
            if (i == 5 && j == 0) b &= 0x1F;
            v = (v << 8) | b;
        }
        c[i] = v;
    }
    if (inf) {
        if (sf || !limbs_is_zero(c, 6)) return 1;
        memset(out, 0, sizeof(*out));
        return 0;
    }
    if (limbs_ge(c, FP_P, 6)) return 1;
    fp_t x, y2, y;
    fp_from_canon(&x, c);
    fp_sqr(&y2, &x); fp_mul(&y2, &y2, &x); fp_add(&y2, &y2, &FP_B);
    if (!fp_sqrt(&y, &y2)) return 1;
    if (fp_is_lex_largest(&y) != sf) fp_neg(&y, &y);
    out->x = x; out->y = y;
    return 0;
}

#include "kzg_oracle_pairing.inc"

/* ------------------------------------------------------------------------------------------------------------ */
/* Pippenger MSM (kzg/src/msm) */
static inline int xyzz_is_inf(const p1xyzz_t *p) { return fp_is_zero(&p->zzz) && fp_is_zero(&p->zz); }

/* p1_dadd_affine (kzg/src/msm/pippenger_utils.rs:90-157) */
static void p1_dadd_affine(p1xyzz_t *out, const p1_affine_t *p2, int subtract) {
    if (fp_is_zero(&p2->x) && fp_is_zero(&p2->y)) return;
    if (xyzz_is_inf(out)) {
        out->x = p2->x; out->y = p2->y;
        out->zzz = FP_ONE;
        if (subtract) fp_neg(&out->zzz, &out->zzz);
        out->zz = FP_ONE;
        return;
    }
    fp_t p, r, pp, ppp, q, t;
    fp_mul(&p, &p2->x, &out->zz);
    fp_mul(&r, &p2->y, &out->zzz);
    if (subtract) fp_neg(&r, &r);
    fp_sub(&p, &p, &out->x);
    fp_sub(&r, &r, &out->y);
    if (!fp_is_zero(&p)) {
        fp_sqr(&pp, &p);
        fp_mul(&ppp, &pp, &p);
        fp_mul(&q, &out->x, &pp);
        fp_sqr(&out->x, &r);
        fp_add(&t, &q, &q);
        fp_sub(&out->x, &out->x, &ppp);
        fp_sub(&out->x, &out->x, &t);
        fp_sub(&q, &q, &out->x);
        fp_mul(&q, &q, &r);
        fp_mul(&out->y, &out->y, &ppp);
        fp_sub(&out->y, &q, &out->y);
        fp_mul(&out->zz, &out->zz, &pp);
        fp_mul(&out->zzz, &out->zzz, &ppp);
    } else if (fp_is_zero(&r)) {
        fp_t u, s, m;
        fp_add(&u, &p2->y, &p2->y);
        fp_sqr(&out->zz, &u);
        fp_mul(&out->zzz, &out->zz, &u);
        fp_mul(&s, &p2->x, &out->zz);
        fp_sqr(&m, &p2->x);
        fp_add(&t, &m, &m); fp_add(&m, &t, &m);
        fp_sqr(&out->x, &m);
        fp_add(&u, &s, &s);
        fp_sub(&out->x, &out->x, &u);
        fp_mul(&out->y, &out->zzz, &p2->y);
        fp_sub(&s, &s, &out->x);
        fp_mul(&s, &s, &m);
        fp_sub(&out->y, &s, &out->y);
        if (subtract) fp_neg(&out->zzz, &out->zzz);
    } else {
        memset(&out->zzz, 0, 2 * sizeof(fp_t));
    }
}
/* p1_dadd (pippenger_utils.rs:159-210) */
static void p1_dadd(p1xyzz_t *out, const p1xyzz_t *p2) {
    if (xyzz_is_inf(p2)) return;
    if (xyzz_is_inf(out)) { *out = *p2; return; }
    fp_t u, s, p, r, pp, ppp, q, t;
    fp_mul(&u, &out->x, &p2->zz);
    fp_mul(&s, &out->y, &p2->zzz);
    fp_mul(&p, &p2->x, &out->zz);
    fp_mul(&r, &p2->y, &out->zzz);
    fp_sub(&p, &p, &u);
    fp_sub(&r, &r, &s);
    if (!fp_is_zero(&p)) {
        fp_sqr(&pp, &p);
        fp_mul(&ppp, &pp, &p);
        fp_mul(&q, &u, &pp);
        fp_sqr(&out->x, &r);
        fp_add(&t, &q, &q);
        fp_sub(&out->x, &out->x, &ppp);
        fp_sub(&out->x, &out->x, &t);
        fp_sub(&q, &q, &out->x);
        fp_mul(&q, &q, &r);
        fp_mul(&out->y, &s, &ppp);
        fp_sub(&out->y, &q, &out->y);
        fp_mul(&out->zz, &out->zz, &p2->zz);
        fp_mul(&out->zzz, &out->zzz, &p2->zzz);
        fp_mul(&out->zz, &out->zz, &pp);
        fp_mul(&out->zzz, &out->zzz, &ppp);
    } else if (fp_is_zero(&r)) {
        fp_t v, w, m;
        fp_add(&u, &out->y, &out->y);
        fp_sqr(&v, &u);
        fp_mul(&w, &v, &u);
        fp_mul(&s, &out->x, &v);
        fp_sqr(&m, &out->x);
        fp_add(&t, &m, &m); fp_add(&m, &t, &m);
        fp_sqr(&out->x, &m);
        fp_add(&u, &s, &s);
        fp_sub(&out->x, &out->x, &u);
        fp_mul(&out->y, &w, &out->y);
        fp_sub(&s, &s, &out->x);
        fp_mul(&s, &s, &m);
        fp_sub(&out->y, &s, &out->y);
        fp_mul(&out->zz, &out->zz, &v);
        fp_mul(&out->zzz, &out->zzz, &w);
    } else {
        memset(&out->zzz, 0, 2 * sizeof(fp_t));
    }
}
/* p1_to_jacobian (pippenger_utils.rs:84-88) */
static void xyzz_to_jacobian(p1_t *out, const p1xyzz_t *in) {
    fp_mul(&out->x, &in->x, &in->zz);
    fp_mul(&out->y, &in->y, &in->zzz);
    out->z = in->zz;
}
/* p1_integrate_buckets (tiling_pippenger_ops.rs:21-45) */
static void p1_integrate_buckets(p1_t *out, p1xyzz_t *buckets, int wbits) {
    size_t n = ((size_t)1 << wbits) - 1;
    p1xyzz_t ret = buckets[n], acc = buckets[n];
    memset(&buckets[n], 0, sizeof(p1xyzz_t));
    while (n--) {
        if (!(fp_is_zero(&buckets[n].x) && fp_is_zero(&buckets[n].y) && xyzz_is_inf(&buckets[n]))) {
            p1_dadd(&acc, &buckets[n]);
            memset(&buckets[n], 0, sizeof(p1xyzz_t));
        }
        p1_dadd(&ret, &acc);
    }
    xyzz_to_jacobian(out, &ret);
}
/* get_wval_limb (pippenger_utils.rs:231-244): bits [off, off+bits) of the little-endian scalar, trash above */
static inline u64 get_wval_limb(const uint8_t *d, size_t off, size_t bits) {
    size_t top = (off + bits - 1) / 8;
    u64 ret = 0;
    size_t lo = off / 8;
    for (size_t i = 0; i < 4 && lo + i <= top && lo + i < 32; i++) ret |= (u64)d[lo + i] << (8 * i);
    return ret >> (off % 8);
}
static inline u64 booth_encode(u64 wval, size_t sz) {  /* pippenger_utils.rs:251-256 */
    u64 mask = (u64)0 - (wval >> sz);
    wval = (wval + 1) >> 1;
    return (wval ^ mask) - mask;
}
/* p1s_tile_pippenger (tiling_pippenger_ops.rs:68-104) */
static void p1s_tile_pippenger(p1_t *ret, const p1_affine_t *points, const u64 (*scalars)[4], size_t n,
                               p1xyzz_t *buckets, size_t bit0, size_t wbits, size_t cbits) {
    u64 wmask = ((u64)1 << (wbits + 1)) - 1;
    u64 z = bit0 == 0;
    bit0 -= (z ^ 1);
    wbits += (z ^ 1);
    for (size_t i = 0; i < n; i++) {
        u64 wval = (get_wval_limb((const uint8_t *)scalars[i], bit0, wbits) << z) & wmask;
        wval = booth_encode(wval, cbits);
        /* booth_decode (pippenger_utils.rs:270-281) */
        int sign = (wval >> cbits) & 1;
        u64 idx = wval & (((u64)1 << cbits) - 1);
        if (idx) p1_dadd_affine(&buckets[idx - 1], &points[i], sign);
    }
    p1_integrate_buckets(ret, buckets, (int)cbits - 1);
}
static void p1s_tile_pippenger_pub(p1_t *ret, const p1_affine_t *points, const u64 (*scalars)[4], size_t n,
                                   p1xyzz_t *buckets, size_t bit0, size_t window) {  /* :47-66 */
    size_t wbits = window, cbits = window;
    if (bit0 + window > 255) { wbits = 255 - bit0; cbits = wbits + 1; }
    p1s_tile_pippenger(ret, points, scalars, n, buckets, bit0, wbits, cbits);
}
static size_t num_bits(size_t l) { size_t n = 0; while (l) { n++; l >>= 1; } return n; }
API size_t ko_pippenger_window_size(size_t npoints) {  /* pippenger_utils.rs:300-317 */
    size_t wbits = num_bits(npoints);
    if (wbits > 13) return wbits - 4;
    if (wbits > 5) return wbits - 3;
    return 2;
}
/* tiling_pippenger (tiling_pippenger_ops.rs:106-138) on canonical scalars and affine points */
static void tiling_pippenger(p1_t *out, const p1_affine_t *points, const u64 (*scalars)[4], size_t n) {
    size_t window = ko_pippenger_window_size(n);
    p1xyzz_t *buckets = (p1xyzz_t *)calloc((size_t)1 << (window - 1), sizeof(p1xyzz_t));
    size_t wbits = 255 % window, cbits = wbits + 1, bit0 = 255;
    p1_t tile, ret;
    p1_set_inf(&ret);
    for (;;) {
        bit0 -= wbits;
        if (bit0 == 0) break;
        p1s_tile_pippenger(&tile, points, scalars, n, buckets, bit0, wbits, cbits);
        p1_add_or_double(&ret, &ret, &tile);
        for (size_t i = 0; i < window; i++) p1_double(&ret, &ret);
        cbits = window;
        wbits = window;
    }
    p1s_tile_pippenger(&tile, points, scalars, n, buckets, 0, wbits, cbits);
    p1_add_or_double(&ret, &ret, &tile);
    free(buckets);
    *out = ret;
}

/* breakdown (kzg/src/msm/parallel_pippenger_utils.rs:3-47) */
static size_t div_ceil(size_t a, size_t b) { return (a + b - 1) / b; }
static void breakdown(size_t window, size_t ncpus, size_t *pnx, size_t *pny, size_t *pwnd) {
    const size_t NBITS = 255;
    size_t nx, wnd;
    if (NBITS > window * ncpus) {
        nx = 1;
        wnd = num_bits(ncpus / 4);
        if (window + wnd > 18) {
            wnd = window - wnd;
        } else {
            wnd = div_ceil(NBITS / window, ncpus);
            if (div_ceil(NBITS / (window + 1), ncpus) < wnd) wnd = window + 1; else wnd = window;
        }
    } else {
        nx = 2;
        wnd = window - 2;
        while ((NBITS / wnd + 1) * nx < ncpus) {
            nx += 1;
            wnd = window - num_bits(3 * nx / 2);
        }
        nx -= 1;
        wnd = window - num_bits(3 * nx / 2);
    }
    size_t ny = NBITS / wnd + 1;
    wnd = NBITS / ny + 1;
    *pnx = nx; *pny = ny; *pwnd = wnd;
}

/* tiling_parallel_pippenger (kzg/src/msm/tiling_parallel_pippenger.rs:70-186): (point-slice x window) tile grid
 * pulled from an atomic counter by nthreads workers; rows combined MSB first. */
typedef struct { size_t x, dx, y, dy; p1_t res; } tile_t;
typedef struct {
    const p1_affine_t *points; const u64 (*scalars)[4];
    tile_t *grid; size_t total, window; size_t *counter;
} ppjob_t;
static void *pp_worker(void *arg) {
    ppjob_t *j = (ppjob_t *)arg;
    p1xyzz_t *buckets = (p1xyzz_t *)calloc((size_t)1 << (j->window - 1), sizeof(p1xyzz_t));
    for (;;) {
        size_t w = __atomic_fetch_add(j->counter, 1, __ATOMIC_RELAXED);
        if (w >= j->total) break;
        tile_t *t = &j->grid[w];
        p1s_tile_pippenger_pub(&t->res, j->points + t->x, j->scalars + t->x, t->dx, buckets, t->y, j->window);
    }
    free(buckets);
    return NULL;
}
static void tiling_parallel_pippenger(p1_t *out, const p1_affine_t *points, const u64 (*scalars)[4], size_t n,
                                      size_t ncpus) {
    if (ncpus < 2 || n < 32) { tiling_pippenger(out, points, scalars, n); return; }
    size_t nx, ny, window;
    breakdown(ko_pippenger_window_size(n), ncpus, &nx, &ny, &window);
    tile_t *grid = (tile_t *)calloc(nx * ny, sizeof(tile_t));
    size_t dx = n / nx, y = window * (ny - 1), total = 0;
    while (total < nx) {
        grid[total].x = total * dx; grid[total].dx = dx; grid[total].y = y; grid[total].dy = 255 - y;
        total++;
    }
    grid[total - 1].dx = n - grid[total - 1].x;
    while (y != 0) {
        y -= window;
        for (size_t i = 0; i < nx; i++) {
            grid[total].x = grid[i].x; grid[total].dx = grid[i].dx; grid[total].y = y; grid[total].dy = window;
            total++;
        }
    }
    size_t counter = 0;
    ppjob_t job = {points, scalars, grid, total, window, &counter};
    size_t nworkers = ncpus < total ? ncpus : total;
    pthread_t *th = (pthread_t *)malloc(nworkers * sizeof(pthread_t));
    for (size_t i = 0; i < nworkers; i++) pthread_create(&th[i], NULL, pp_worker, &job);
    for (size_t i = 0; i < nworkers; i++) pthread_join(th[i], NULL);
    free(th);
    /* combine rows MSB first (the reference overlaps this with the workers through a channel; same result) */
    p1_t ret; p1_set_inf(&ret);
    size_t row = 0;
    while (row < total) {
        size_t yy = grid[row].y;
        while (row < total && grid[row].y == yy) { p1_add_or_double(&ret, &ret, &grid[row].res); row++; }
        if (yy == 0) break;
        for (size_t i = 0; i < window; i++) p1_double(&ret, &ret);
    }
    free(grid);
    *out = ret;
}

/* msm() (kzg/src/msm/msm_impls.rs:114-148) with precomputation = None.
 * points: Jacobian (FsG1), scalars: Montgomery Fr.  nthreads <= 1 -> msm_sequential, else msm_parallel. */
API void ko_g1_lincomb(p1_t *out, const p1_t *points, const fr_t *scalars, size_t len, int nthreads) {
    if (len < 8) {
        p1_t acc; p1_set_inf(&acc);
        for (size_t i = 0; i < len; i++) {
            p1_t t; ko_p1_mult(&t, &points[i], &scalars[i]);
            p1_add_or_double(&acc, &acc, &t);
        }
        *out = acc;
        return;
    }
    /* pippenger(): filter infinity, batch_convert, to_scalar (msm_impls.rs:40-61) */
    p1_t *pts = (p1_t *)malloc(len * sizeof(p1_t));
    u64 (*scs)[4] = (u64 (*)[4])malloc(len * 32 + 8);
    size_t n = 0;
    for (size_t i = 0; i < len; i++) {
        if (p1_is_inf(&points[i])) continue;
        pts[n] = points[i];
        fr_to_canon(scs[n], &scalars[i]);
        n++;
    }
    p1_affine_t *aff = (p1_affine_t *)malloc((n + 1) * sizeof(p1_affine_t));
    ko_p1s_to_affine(aff, pts, n);
    if (n == 0) p1_set_inf(out);
    else if (nthreads > 1) tiling_parallel_pippenger(out, aff, (const u64 (*)[4])scs, n, (size_t)nthreads);
    else tiling_pippenger(out, aff, (const u64 (*)[4])scs, n);
    free(aff); free(scs); free(pts);
}
/* same, on affine points (the sppark-shaped FFI's input: blst_p1_affine + Montgomery blst_fr) */
API void ko_msm_affine(p1_t *out, const p1_affine_t *points, const fr_t *scalars, size_t len, int nthreads) {
    p1_affine_t *aff = (p1_affine_t *)malloc((len + 1) * sizeof(p1_affine_t));
    u64 (*scs)[4] = (u64 (*)[4])malloc(len * 32 + 8);
    size_t n = 0;
    for (size_t i = 0; i < len; i++) {
        if (fp_is_zero(&points[i].x) && fp_is_zero(&points[i].y)) continue;
        aff[n] = points[i];
        fr_to_canon(scs[n], &scalars[i]);
        n++;
    }
    if (n == 0) p1_set_inf(out);
    else if (n < 8) {
        p1_t acc; p1_set_inf(&acc);
        for (size_t i = 0; i < n; i++) {
            p1_t p, t; p1_from_affine(&p, &aff[i]);
            p1_mult_canon(&t, &p, scs[i], 255);
            p1_add_or_double(&acc, &acc, &t);
        }
        *out = acc;
    } else if (nthreads > 1) tiling_parallel_pippenger(out, aff, (const u64 (*)[4])scs, n, (size_t)nthreads);
    else tiling_pippenger(out, aff, (const u64 (*)[4])scs, n);
    free(aff); free(scs);
}
/* naive sum of scalar multiples: the reference tests' expected value (kzg-bench/src/tests/bls12_381.rs:238-243) */
API void ko_msm_naive(p1_t *out, const p1_t *points, const fr_t *scalars, size_t len) {
    p1_t acc; p1_set_inf(&acc);
    for (size_t i = 0; i < len; i++) {
        p1_t t; ko_p1_mult(&t, &points[i], &scalars[i]);
        p1_add_or_double(&acc, &acc, &t);
    }
    *out = acc;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* FFTSettings (blst/src/types/fft_settings.rs:28-58, 90-106) */
typedef struct {
    size_t max_width;
    fr_t *roots_of_unity;          /* max_width + 1 */
    fr_t *brp_roots_of_unity;      /* max_width */
    fr_t *reverse_roots_of_unity;  /* max_width + 1 */
} fft_settings_t;

static void fr_pow_u(fr_t *r, const fr_t *a, const u64 e[4]) { fr_pow(r, a, e, 4); }
/* SCALE2_ROOT_OF_UNITY[scale] = 7^((r-1)/2^scale) (blst/src/consts.rs:14-50; equality checked by the tests) */
API void ko_scale2_root_of_unity(fr_t *out, int scale) {
    u64 e[4], one[4] = {1, 0, 0, 0};
    limbs_sub(e, FR_R, one, 4);
    for (int s = 0; s < scale; s++)
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);
    fr_t seven; fr_from_u64(&seven, 7);
    fr_pow_u(out, &seven, e);
}
static size_t brp_index(size_t i, int bits) {
    size_t r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}
static int log2_pow2(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }

API void *ko_fft_settings_new(int scale) {
    ko_init();
    if (scale < 0 || scale >= 32) return NULL;
    fft_settings_t *fs = (fft_settings_t *)calloc(1, sizeof(*fs));
    size_t w = (size_t)1 << scale;
    fs->max_width = w;
    fs->roots_of_unity = (fr_t *)malloc((w + 1) * sizeof(fr_t));
    fs->brp_roots_of_unity = (fr_t *)malloc(w * sizeof(fr_t));
    fs->reverse_roots_of_unity = (fr_t *)malloc((w + 1) * sizeof(fr_t));
    fr_t root; ko_scale2_root_of_unity(&root, scale);
    fs->roots_of_unity[0] = FR_ONE;
    for (size_t i = 1; i <= w; i++) fr_mul(&fs->roots_of_unity[i], &fs->roots_of_unity[i - 1], &root);
    for (size_t i = 0; i <= w; i++) fs->reverse_roots_of_unity[i] = fs->roots_of_unity[w - i];
    for (size_t i = 0; i < w; i++) fs->brp_roots_of_unity[i] = fs->roots_of_unity[brp_index(i, scale)];
    return fs;
}
API void ko_fft_settings_free(void *h) {
    fft_settings_t *fs = (fft_settings_t *)h;
    if (!fs) return;
    free(fs->roots_of_unity); free(fs->brp_roots_of_unity); free(fs->reverse_roots_of_unity); free(fs);
}
API const fr_t *ko_fft_settings_roots(void *h, int which) {
    fft_settings_t *fs = (fft_settings_t *)h;
    return which == 0 ? fs->roots_of_unity : which == 1 ? fs->brp_roots_of_unity : fs->reverse_roots_of_unity;
}

/* fft_fr_fast_inner (blst/src/fft_fr.rs:49-108): recursive radix-2, output halves, input stride doubling.
 * par_depth > 0 forks the two halves onto two threads (the reference's rayon::join when half > 256). */
typedef struct {
    fr_t *ret; size_t n; const fr_t *data; size_t data_stride; const fr_t *roots; size_t roots_stride; int par_depth;
} fftjob_t;
static void fft_fr_fast(fr_t *ret, size_t n, const fr_t *data, size_t stride, const fr_t *roots, size_t roots_stride,
                        int par_depth);
static void *fft_thread(void *a) {
    fftjob_t *j = (fftjob_t *)a;
    fft_fr_fast(j->ret, j->n, j->data, j->data_stride, j->roots, j->roots_stride, j->par_depth);
    return NULL;
}
static void fft_fr_fast(fr_t *ret, size_t n, const fr_t *data, size_t stride, const fr_t *roots, size_t roots_stride,
                        int par_depth) {
    size_t half = n / 2;
    if (half == 0) { ret[0] = data[0]; return; }
    if (par_depth > 0 && half > 256) {
        fftjob_t j = {ret + half, half, data + stride, stride * 2, roots, roots_stride * 2, par_depth - 1};
        pthread_t th;
        pthread_create(&th, NULL, fft_thread, &j);
        fft_fr_fast(ret, half, data, stride * 2, roots, roots_stride * 2, par_depth - 1);
        pthread_join(th, NULL);
    } else {
        fft_fr_fast(ret, half, data, stride * 2, roots, roots_stride * 2, 0);
        fft_fr_fast(ret + half, half, data + stride, stride * 2, roots, roots_stride * 2, 0);
    }
    for (size_t i = 0; i < half; i++) {
        fr_t yr;
        fr_mul(&yr, &ret[i + half], &roots[i * roots_stride]);
        fr_sub(&ret[i + half], &ret[i], &yr);
        fr_add(&ret[i], &ret[i], &yr);
    }
}
static int par_depth_for(int nthreads) { int d = 0; while ((1 << d) < nthreads) d++; return nthreads > 1 ? d : 0; }
/* FFTFr::fft_fr / fft_fr_output (blst/src/fft_fr.rs:112-165). returns 0 ok, 1 Err */
API int ko_fft_fr(void *h, fr_t *out, const fr_t *data, size_t n, int inverse, int nthreads) {
    fft_settings_t *fs = (fft_settings_t *)h;
    if (n > fs->max_width) return 1;
    if (n == 0 || (n & (n - 1))) return 1;
    size_t stride = fs->max_width / n;
    const fr_t *roots = inverse ? fs->reverse_roots_of_unity : fs->roots_of_unity;
    fft_fr_fast(out, n, data, 1, roots, stride, par_depth_for(nthreads));
    if (inverse) {
        fr_t inv; fr_from_u64(&inv, (u64)n); fr_inv(&inv, &inv);
        for (size_t i = 0; i < n; i++) fr_mul(&out[i], &out[i], &inv);
    }
    return 0;
}
/* fft_fr_slow (blst/src/fft_fr.rs:168-186) */
API void ko_fft_fr_slow(void *h, fr_t *out, const fr_t *data, size_t n, int inverse) {
    fft_settings_t *fs = (fft_settings_t *)h;
    size_t rs = fs->max_width / n;
    const fr_t *roots = inverse ? fs->reverse_roots_of_unity : fs->roots_of_unity;
    for (size_t i = 0; i < n; i++) {
        fr_t acc; fr_mul(&acc, &data[0], &roots[0]);
        for (size_t j = 1; j < n; j++) {
            fr_t v; fr_mul(&v, &data[j], &roots[((i * j) % n) * rs]);
            fr_add(&acc, &acc, &v);
        }
        out[i] = acc;
    }
    if (inverse) {
        fr_t inv; fr_from_u64(&inv, (u64)n); fr_inv(&inv, &inv);
        for (size_t i = 0; i < n; i++) fr_mul(&out[i], &out[i], &inv);
    }
}
/* fft_g1_fast + FFTG1::fft_g1 (blst/src/fft_g1.rs:13-83): the same recursion over G1 with a full scalar
 * multiplication per butterfly.  returns 0 ok, 1 Err */
static void fft_g1_fast(p1_t *ret, size_t n, const p1_t *data, size_t stride, const fr_t *roots, size_t roots_stride) {
    size_t half = n / 2;
    if (half == 0) { ret[0] = data[0]; return; }
    fft_g1_fast(ret, half, data, stride * 2, roots, roots_stride * 2);
    fft_g1_fast(ret + half, half, data + stride, stride * 2, roots, roots_stride * 2);
    for (size_t i = 0; i < half; i++) {
        p1_t yr, neg;
        ko_p1_mult(&yr, &ret[i + half], &roots[i * roots_stride]);
        neg = yr; fp_neg(&neg.y, &neg.y);
        p1_add_or_double(&ret[i + half], &ret[i], &neg);
        p1_add_or_double(&ret[i], &ret[i], &yr);
    }
}
API int ko_fft_g1(void *h, p1_t *out, const p1_t *data, size_t n, int inverse) {
    fft_settings_t *fs = (fft_settings_t *)h;
    if (n > fs->max_width) return 1;
    if (n == 0 || (n & (n - 1))) return 1;
    size_t stride = fs->max_width / n;
    fft_g1_fast(out, n, data, 1, inverse ? fs->reverse_roots_of_unity : fs->roots_of_unity, stride);
    if (inverse) {
        fr_t inv; fr_from_u64(&inv, (u64)n); fr_inv(&inv, &inv);
        for (size_t i = 0; i < n; i++) ko_p1_mult(&out[i], &out[i], &inv);
    }
    return 0;
}
/* fft_g1_slow (blst/src/fft_g1.rs:86-104) */
API void ko_fft_g1_slow(void *h, p1_t *out, const p1_t *data, size_t n, int inverse) {
    fft_settings_t *fs = (fft_settings_t *)h;
    size_t rs = fs->max_width / n;
    const fr_t *roots = inverse ? fs->reverse_roots_of_unity : fs->roots_of_unity;
    for (size_t i = 0; i < n; i++) {
        p1_t acc; ko_p1_mult(&acc, &data[0], &roots[0]);
        for (size_t j = 1; j < n; j++) {
            p1_t v; ko_p1_mult(&v, &data[j], &roots[((i * j) % n) * rs]);
            p1_add_or_double(&acc, &acc, &v);
        }
        out[i] = acc;
    }
}

/* das_fft_extension_stride (blst/src/data_availability_sampling.rs:14-71) */
static void das_stride(const fft_settings_t *fs, fr_t *ev, size_t n, size_t stride) {
    if (n < 2) return;
    if (n == 2) {
        fr_t x, y, yr;
        fr_add(&x, &ev[0], &ev[1]); fr_sub(&y, &ev[0], &ev[1]);
        fr_mul(&yr, &y, &fs->roots_of_unity[stride]);
        fr_add(&ev[0], &x, &yr); fr_sub(&ev[1], &x, &yr);
        return;
    }
    size_t half = n / 2;
    for (size_t i = 0; i < half; i++) {
        fr_t t1, t2;
        fr_add(&t1, &ev[i], &ev[half + i]);
        fr_sub(&t2, &ev[i], &ev[half + i]);
        fr_mul(&ev[half + i], &t2, &fs->reverse_roots_of_unity[i * 2 * stride]);
        ev[i] = t1;
    }
    das_stride(fs, ev, half, stride * 2);
    das_stride(fs, ev + half, half, stride * 2);
    for (size_t i = 0; i < half; i++) {
        fr_t x = ev[i], y = ev[half + i], yr;
        fr_mul(&yr, &y, &fs->roots_of_unity[(1 + 2 * i) * stride]);
        fr_add(&ev[i], &x, &yr); fr_sub(&ev[half + i], &x, &yr);
    }
}
/* DASExtension::das_fft_extension (:78-100). returns 0 ok, 1 Err */
API int ko_das_fft_extension(void *h, fr_t *odds, const fr_t *evens, size_t n) {
    fft_settings_t *fs = (fft_settings_t *)h;
    if (n == 0 || (n & (n - 1)) || n * 2 > fs->max_width) return 1;
    size_t stride = fs->max_width / (n * 2);
    memmove(odds, evens, n * sizeof(fr_t));
    das_stride(fs, odds, n, stride);
    fr_t inv; fr_from_u64(&inv, (u64)n); fr_inv(&inv, &inv);
    for (size_t i = 0; i < n; i++) fr_mul(&odds[i], &odds[i], &inv);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* SHA-256 (FIPS 180-4) -- the reference uses the sha2 crate (kzg/src/eip_4844.rs:237-239) */
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha256_block(uint32_t st[8], const uint8_t *p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25), ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + SHA_K[i] + w[i];
        uint32_t S0 = ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
API void ko_sha256(uint8_t out[32], const uint8_t *msg, size_t len) {
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t i = 0;
    for (; i + 64 <= len; i += 64) sha256_block(st, msg + i);
    uint8_t tail[128] = {0};
    size_t rem = len - i;
    memcpy(tail, msg + i, rem);
    tail[rem] = 0x80;
    size_t tl = rem + 1 + 8 <= 64 ? 64 : 128;
    u64 bits = (u64)len * 8;
    for (int j = 0; j < 8; j++) tail[tl - 1 - j] = (uint8_t)(bits >> (8 * j));
    sha256_block(st, tail);
    if (tl == 128) sha256_block(st, tail + 64);
    for (int j = 0; j < 8; j++) { out[4 * j] = st[j] >> 24; out[4 * j + 1] = st[j] >> 16; out[4 * j + 2] = st[j] >> 8; out[4 * j + 3] = st[j]; }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* EIP-4844 (kzg/src/eip_4844.rs) */
#define FIELD_ELEMENTS_PER_BLOB 4096
#define BYTES_PER_BLOB 131072
typedef struct {
    fft_settings_t *fs;                 /* scale 13 (eip_4844.rs:1072-1077) */
    p1_t *g1_lagrange_brp;              /* 4096, bit-reversed (eip_4844.rs:1070) */
    p1_t *g1_monomial;                  /* 4096 */
    p1_t *x_ext_fft_columns;            /* [128 rows][64 offsets], built on first use (kzg_settings.rs:84-101) */
    p2_affine_t g2_monomial[65];        /* [s^i]G2 (eip_4844.rs:1050-1053) */
    int nthreads;
} kzg_settings_t;

static int hexval(int c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }

/* load_trusted_setup_string + load_trusted_setup_rust (eip_4844.rs:151-228, 1022-1086).
 * including the pairing sanity check (is_trusted_setup_in_lagrange_form, :1005-1020). */
API void *ko_load_trusted_setup_text(const char *text, size_t len) {
    ko_init();
    const char *p = text, *end = text + len;
    size_t counts[2];
    for (int k = 0; k < 2; k++) {
        while (p < end && isspace((unsigned char)*p)) p++;
        size_t v = 0; int any = 0;
        while (p < end && isdigit((unsigned char)*p)) { v = v * 10 + (size_t)(*p - '0'); p++; any = 1; }
        if (!any) return NULL;
        counts[k] = v;
    }
    if (counts[0] != FIELD_ELEMENTS_PER_BLOB || counts[1] != 65) return NULL;
    size_t nbytes = 4096 * 48 + 65 * 96 + 4096 * 48;
    uint8_t *raw = (uint8_t *)malloc(nbytes);
    for (size_t i = 0; i < nbytes; i++) {
        while (p < end && isspace((unsigned char)*p)) p++;
        if (p + 1 >= end + 1 || p >= end) { free(raw); return NULL; }
        int hi = hexval(p[0]), lo = (p + 1 < end) ? hexval(p[1]) : -1;
        if (hi < 0) { free(raw); return NULL; }
        if (lo >= 0) { raw[i] = (uint8_t)(hi << 4 | lo); p += 2; } else { raw[i] = (uint8_t)hi; p += 1; }
    }
    kzg_settings_t *s = (kzg_settings_t *)calloc(1, sizeof(*s));
    s->g1_lagrange_brp = (p1_t *)malloc(4096 * sizeof(p1_t));
    s->g1_monomial = (p1_t *)malloc(4096 * sizeof(p1_t));
    const uint8_t *lag = raw, *mono = raw + 4096 * 48 + 65 * 96, *g2b = raw + 4096 * 48;
    int bad = 0;
    for (size_t i = 0; i < 65 && !bad; i++)
        if (ko_p2_uncompress(&s->g2_monomial[i], g2b + 96 * i)) bad = 1;
    for (size_t i = 0; i < 4096 && !bad; i++) {
        p1_affine_t a;
        if (ko_p1_uncompress(&a, lag + 48 * i)) bad = 1;
        p1_from_affine(&s->g1_lagrange_brp[brp_index(i, 12)], &a);
        if (ko_p1_uncompress(&a, mono + 48 * i)) bad = 1;
        p1_from_affine(&s->g1_monomial[i], &a);
    }
    free(raw);
    /* is_trusted_setup_in_lagrange_form (eip_4844.rs:1005-1020, 1064-1068); lagrange[1] is at bit-reversed slot 2048 */
    if (!bad && pairings_verify(&s->g1_lagrange_brp[2048], &s->g2_monomial[0], &s->g1_lagrange_brp[0], &s->g2_monomial[1])) bad = 1;
    if (bad) { free(s->g1_lagrange_brp); free(s->g1_monomial); free(s); return NULL; }
    s->fs = (fft_settings_t *)ko_fft_settings_new(13);
    s->nthreads = 1;
    return s;
}
API void ko_settings_set_threads(void *h, int nthreads) { ((kzg_settings_t *)h)->nthreads = nthreads; }
API const p1_t *ko_settings_g1_lagrange_brp(void *h) { return ((kzg_settings_t *)h)->g1_lagrange_brp; }
API const p1_t *ko_settings_g1_monomial(void *h) { return ((kzg_settings_t *)h)->g1_monomial; }
API void *ko_settings_fft(void *h) { return ((kzg_settings_t *)h)->fs; }
API void ko_free_trusted_setup(void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    if (!s) return;
    ko_fft_settings_free(s->fs); free(s->g1_lagrange_brp); free(s->g1_monomial); free(s->x_ext_fft_columns); free(s);
}

/* bytes_to_blob (eip_4844.rs:867-880). returns 0 ok, 1 Err */
static int bytes_to_blob(fr_t *poly, const uint8_t *blob) {
    for (size_t i = 0; i < FIELD_ELEMENTS_PER_BLOB; i++)
        if (ko_fr_from_bendian(&poly[i], blob + 32 * i)) return 1;
    return 0;
}
/* blob_to_kzg_commitment_raw (eip_4844.rs:297-314) */
API int ko_blob_to_kzg_commitment(uint8_t out[48], const uint8_t *blob, void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    fr_t *poly = (fr_t *)malloc(FIELD_ELEMENTS_PER_BLOB * sizeof(fr_t));
    if (bytes_to_blob(poly, blob)) { free(poly); return 1; }
    p1_t c;
    ko_g1_lincomb(&c, s->g1_lagrange_brp, poly, FIELD_ELEMENTS_PER_BLOB, s->nthreads);
    ko_p1_compress(out, &c);
    free(poly);
    return 0;
}
/* compute_challenge_rust (eip_4844.rs:920-945) */
static void compute_challenge(fr_t *out, const fr_t *poly, const p1_t *commitment) {
    size_t len = 32 + BYTES_PER_BLOB + 48;
    uint8_t *buf = (uint8_t *)calloc(1, len);
    memcpy(buf, "FSBLOBVERIFY_V1_", 16);
    buf[30] = (FIELD_ELEMENTS_PER_BLOB >> 8) & 0xff; buf[31] = FIELD_ELEMENTS_PER_BLOB & 0xff;
    for (size_t i = 0; i < FIELD_ELEMENTS_PER_BLOB; i++) ko_fr_to_bendian(buf + 32 + 32 * i, &poly[i]);
    ko_p1_compress(buf + 32 + BYTES_PER_BLOB, commitment);
    uint8_t d[32];
    ko_sha256(d, buf, len);
    ko_fr_from_bendian_unchecked(out, d);
    free(buf);
}
API int ko_compute_challenge(uint8_t out[32], const uint8_t *blob, const uint8_t commitment[48]) {
    fr_t *poly = (fr_t *)malloc(FIELD_ELEMENTS_PER_BLOB * sizeof(fr_t));
    p1_affine_t a; p1_t c; fr_t z;
    if (bytes_to_blob(poly, blob) || ko_p1_uncompress(&a, commitment)) { free(poly); return 1; }
    p1_from_affine(&c, &a);
    compute_challenge(&z, poly, &c);
    ko_fr_to_bendian(out, &z);
    free(poly);
    return 0;
}
/* fr_batch_inv (eip_4844.rs:882-914). returns 0 ok, 1 Err("Zero input") */
static int fr_batch_inv(fr_t *out, const fr_t *a, size_t len) {
    fr_t acc = FR_ONE;
    for (size_t i = 0; i < len; i++) { out[i] = acc; fr_mul(&acc, &acc, &a[i]); }
    if (fr_is_zero(&acc)) return 1;
    fr_inv(&acc, &acc);
    for (size_t i = len; i-- > 0;) { fr_mul(&out[i], &out[i], &acc); fr_mul(&acc, &acc, &a[i]); }
    return 0;
}
/* evaluate_polynomial_in_evaluation_form (eip_4844.rs:954-1003) */
static int evaluate_polynomial_in_evaluation_form(fr_t *y, const fr_t *poly, const fr_t *x, const kzg_settings_t *s) {
    const size_t n = FIELD_ELEMENTS_PER_BLOB;
    const fr_t *roots = s->fs->brp_roots_of_unity;
    fr_t *inv_in = (fr_t *)malloc(2 * n * sizeof(fr_t)), *inv = inv_in + n;
    for (size_t i = 0; i < n; i++) {
        if (fr_eq(x, &roots[i])) { *y = poly[i]; free(inv_in); return 0; }
        fr_sub(&inv_in[i], x, &roots[i]);
    }
    if (fr_batch_inv(inv, inv_in, n)) { free(inv_in); return 1; }
    fr_t out; memset(&out, 0, sizeof(out));
    for (size_t i = 0; i < n; i++) {
        fr_t t;
        fr_mul(&t, &inv[i], &roots[i]);
        fr_mul(&t, &t, &poly[i]);
        fr_add(&out, &out, &t);
    }
    fr_t t; fr_from_u64(&t, n); fr_inv(&t, &t);
    fr_mul(&out, &out, &t);
    u64 e[4] = {n, 0, 0, 0};
    fr_pow(&t, x, e, 1);
    fr_sub(&t, &t, &FR_ONE);
    fr_mul(&out, &out, &t);
    *y = out;
    free(inv_in);
    return 0;
}
/* compute_kzg_proof_rust (eip_4844.rs:437-519): quotient polynomial q (returned for the parity tests) + MSM */
static int compute_quotient(fr_t *q, fr_t *y, const fr_t *poly, const fr_t *z, const kzg_settings_t *s) {
    const size_t n = FIELD_ELEMENTS_PER_BLOB;
    const fr_t *roots = s->fs->brp_roots_of_unity;
    if (evaluate_polynomial_in_evaluation_form(y, poly, z, s)) return 1;
    fr_t *inv_in = (fr_t *)malloc(2 * n * sizeof(fr_t)), *inv = inv_in + n;
    size_t m = 0;
    memset(q, 0, n * sizeof(fr_t));
    for (size_t i = 0; i < n; i++) {
        if (fr_eq(z, &roots[i])) { m = i + 1; inv_in[i] = FR_ONE; continue; }
        fr_sub(&q[i], &poly[i], y);
        fr_sub(&inv_in[i], &roots[i], z);
    }
    if (fr_batch_inv(inv, inv_in, n)) { free(inv_in); return 1; }
    for (size_t i = 0; i < n; i++) fr_mul(&q[i], &q[i], &inv[i]);
    if (m != 0) {
        m -= 1;
        memset(&q[m], 0, sizeof(fr_t));
        for (size_t i = 0; i < n; i++) {
            if (i == m) continue;
            fr_t t; fr_sub(&t, z, &roots[i]);
            fr_mul(&inv_in[i], &t, z);
        }
        if (fr_batch_inv(inv, inv_in, n)) { free(inv_in); return 1; }
        for (size_t i = 0; i < n; i++) {
            if (i == m) continue;
            fr_t t; fr_sub(&t, &poly[i], y);
            fr_mul(&t, &t, &roots[i]);
            fr_mul(&t, &t, &inv[i]);
            fr_add(&q[m], &q[m], &t);
        }
    }
    free(inv_in);
    return 0;
}
API int ko_compute_quotient(fr_t *q, fr_t *y, const fr_t *poly, const fr_t *z, void *h) {
    return compute_quotient(q, y, poly, z, (kzg_settings_t *)h);
}
/* compute_kzg_proof_raw (eip_4844.rs:521-539) */
API int ko_compute_kzg_proof(uint8_t proof[48], uint8_t y_out[32], const uint8_t *blob, const uint8_t z_bytes[32], void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    fr_t *poly = (fr_t *)malloc(2 * FIELD_ELEMENTS_PER_BLOB * sizeof(fr_t)), *q = poly + FIELD_ELEMENTS_PER_BLOB;
    fr_t z, y;
    int rc = 1;
    if (!bytes_to_blob(poly, blob) && !ko_fr_from_bendian(&z, z_bytes) && !compute_quotient(q, &y, poly, &z, s)) {
        p1_t pr;
        ko_g1_lincomb(&pr, s->g1_lagrange_brp, q, FIELD_ELEMENTS_PER_BLOB, s->nthreads);
        ko_p1_compress(proof, &pr);
        ko_fr_to_bendian(y_out, &y);
        rc = 0;
    }
    free(poly);
    return rc;
}
/* compute_blob_kzg_proof_raw (eip_4844.rs:541-584) */
API int ko_compute_blob_kzg_proof(uint8_t proof[48], const uint8_t *blob, const uint8_t commitment[48], void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    fr_t *poly = (fr_t *)malloc(2 * FIELD_ELEMENTS_PER_BLOB * sizeof(fr_t)), *q = poly + FIELD_ELEMENTS_PER_BLOB;
    p1_affine_t a; p1_t c; fr_t z, y;
    int rc = 1;
    if (!bytes_to_blob(poly, blob) && !ko_p1_uncompress(&a, commitment)) {
        p1_from_affine(&c, &a);
        if (p1_is_inf(&c) || ko_p1_in_g1(&c)) {
            compute_challenge(&z, poly, &c);
            if (!compute_quotient(q, &y, poly, &z, s)) {
                p1_t pr;
                ko_g1_lincomb(&pr, s->g1_lagrange_brp, q, FIELD_ELEMENTS_PER_BLOB, s->nthreads);
                ko_p1_compress(proof, &pr);
                rc = 0;
            }
        }
    }
    free(poly);
    return rc;
}
/* compute_cells (kzg/src/das.rs:244-275, 618-629): BRP + inverse fft_fr(4096), forward fft_fr(8192), BRP.
 * cells_out = 128 * 64 * 32 bytes.  The NTT golden vectors. */
API int ko_compute_cells(uint8_t *cells_out, const uint8_t *blob, void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    fr_t *poly = (fr_t *)malloc(4096 * sizeof(fr_t));
    if (bytes_to_blob(poly, blob)) { free(poly); return 1; }
    fr_t *brp = (fr_t *)calloc(8192, sizeof(fr_t)), *mono = (fr_t *)calloc(8192, sizeof(fr_t)), *ext = (fr_t *)malloc(8192 * sizeof(fr_t));
    for (size_t i = 0; i < 4096; i++) brp[brp_index(i, 12)] = poly[i];
    ko_fft_fr(s->fs, mono, brp, 4096, 1, 1);
    ko_fft_fr(s->fs, ext, mono, 8192, 0, 1);
    for (size_t i = 0; i < 8192; i++) ko_fr_to_bendian(cells_out + 32 * brp_index(i, 13), &ext[i]);
    free(poly); free(brp); free(mono); free(ext);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* EIP-7594 cell proofs by FK20 (kzg/src/das.rs:244-292, 626-696; setup in blst/src/types/kzg_settings.rs:38-101) */
#define CELL_SIZE 64
#define FK_K 64      /* n / cell_size */
#define FK_K2 128
typedef struct { kzg_settings_t *s; size_t next; } fkjob_t;
static void *fk_setup_worker(void *arg) {
    fkjob_t *j = (fkjob_t *)arg;
    kzg_settings_t *s = j->s;
    const size_t n = 4096;
    p1_t *x_ext = (p1_t *)calloc(FK_K2, sizeof(p1_t)), *points = (p1_t *)malloc(FK_K2 * sizeof(p1_t));
    for (;;) {
        size_t offset = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (offset >= CELL_SIZE) break;
        size_t start = n - CELL_SIZE - 1 - offset;
        memset(x_ext, 0, FK_K2 * sizeof(p1_t));
        for (size_t i = 0; i + 1 < FK_K; i++) x_ext[i] = s->g1_monomial[start - i * CELL_SIZE];
        /* toeplitz_part_1: fft_g1_fast over the zero-extended vector, roots stride = 8192 / 128 */
        fft_g1_fast(points, FK_K2, x_ext, 1, s->fs->roots_of_unity, s->fs->max_width / FK_K2);
        for (size_t row = 0; row < FK_K2; row++) s->x_ext_fft_columns[row * CELL_SIZE + offset] = points[row];
    }
    free(x_ext); free(points);
    return NULL;
}
static void fk20_setup(kzg_settings_t *s) {
    if (s->x_ext_fft_columns) return;
    s->x_ext_fft_columns = (p1_t *)malloc(FK_K2 * CELL_SIZE * sizeof(p1_t));
    fkjob_t job = {s, 0};
    int nt = s->nthreads > 1 ? s->nthreads : 1;
    if (nt > 64) nt = 64;
    pthread_t th[64];
    for (int i = 0; i < nt; i++) pthread_create(&th[i], NULL, fk_setup_worker, &job);
    for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
}
API const p1_t *ko_settings_x_ext_fft_columns(void *h) { fk20_setup((kzg_settings_t *)h); return ((kzg_settings_t *)h)->x_ext_fft_columns; }

/* toeplitz_coeffs_stride (kzg/src/das.rs:626-658) */
static void toeplitz_coeffs_stride(fr_t *out, const fr_t *in, size_t n, size_t offset, size_t stride) {
    size_t r = n / stride, l = stride, d = n - 1, d_minus_i = d - offset;
    memset(out, 0, 2 * r * sizeof(fr_t));
    out[0] = in[d_minus_i];
    for (size_t j = 1; j < r - 1; j++) out[2 * r - j] = in[d_minus_i - j * l];
}
/* compute_fk20_proofs (kzg/src/das.rs:660-696) + reverse_bit_order of the proofs (:287): mono = the first 4096
 * monomial coefficients; proofs_out = 128 x 48 compressed bytes in cell order */
static void fk20_proofs_from_mono(uint8_t *proofs_out, const fr_t *mono, kzg_settings_t *s) {
    const size_t n = 4096;
    fk20_setup(s);
    fr_t *coeffs = (fr_t *)malloc(FK_K2 * CELL_SIZE * sizeof(fr_t));   /* [row j][offset i] */
    fr_t tc[FK_K2], tf[FK_K2];
    for (size_t i = 0; i < CELL_SIZE; i++) {
        toeplitz_coeffs_stride(tc, mono, n, i, CELL_SIZE);
        ko_fft_fr(s->fs, tf, tc, FK_K2, 0, 1);
        for (size_t j = 0; j < FK_K2; j++) coeffs[j * CELL_SIZE + i] = tf[j];
    }
    p1_t hext[FK_K2], hh[FK_K2], pr[FK_K2];
    for (size_t j = 0; j < FK_K2; j++)   /* g1_lincomb_batch: one 64-term lincomb per row */
        ko_g1_lincomb(&hext[j], s->x_ext_fft_columns + j * CELL_SIZE, coeffs + j * CELL_SIZE, CELL_SIZE, 1);
    ko_fft_g1(s->fs, hh, hext, FK_K2, 1);
    for (size_t j = FK_K; j < FK_K2; j++) p1_set_inf(&hh[j]);
    ko_fft_g1(s->fs, pr, hh, FK_K2, 0);
    for (size_t j = 0; j < FK_K2; j++) ko_p1_compress(proofs_out + 48 * brp_index(j, 7), &pr[j]);
    free(coeffs);
}
/* compute_cells_and_kzg_proofs (kzg/src/das.rs:244-292) through the byte-level wrapper; cells_out (128*2048 B) and
 * proofs_out (128*48 B) may each be NULL but not both.  returns 0 ok, 1 Err */
API int ko_compute_cells_and_kzg_proofs(uint8_t *cells_out, uint8_t *proofs_out, const uint8_t *blob, void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    if (!cells_out && !proofs_out) return 1;
    const size_t n = 4096;
    fr_t *poly = (fr_t *)malloc(n * sizeof(fr_t));
    if (bytes_to_blob(poly, blob)) { free(poly); return 1; }
    fr_t *brp = (fr_t *)calloc(2 * n, sizeof(fr_t)), *mono = (fr_t *)calloc(2 * n, sizeof(fr_t));
    for (size_t i = 0; i < n; i++) brp[brp_index(i, 12)] = poly[i];
    ko_fft_fr(s->fs, mono, brp, n, 1, 1);                       /* poly_lagrange_to_monomial; upper half stays zero */
    if (cells_out) {
        fr_t *ext = (fr_t *)malloc(2 * n * sizeof(fr_t));
        ko_fft_fr(s->fs, ext, mono, 2 * n, 0, 1);
        for (size_t i = 0; i < 2 * n; i++) ko_fr_to_bendian(cells_out + 32 * brp_index(i, 13), &ext[i]);
        free(ext);
    }
    if (proofs_out) fk20_proofs_from_mono(proofs_out, mono, s);
    free(poly); free(brp); free(mono);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* EIP-4844 verification (kzg/src/eip_4844.rs:328-435, 586-866).  Return 0 = Ok(*ok), 1 = Err (C ABI: BadArgs). */
static const uint8_t G1_GEN_COMPRESSED[48] = {  /* compress(G1 generator), zkcrypto/bls12_381/src/g1.rs:179-200 */
    0x97, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f, 0xc3, 0x68, 0x8c, 0x4f,
    0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58, 0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef,
    0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};
static void g1_generator(p1_t *g) { p1_affine_t a; ko_p1_uncompress(&a, G1_GEN_COMPRESSED); p1_from_affine(g, &a); }
static void p1_neg(p1_t *r, const p1_t *p) { *r = *p; fp_neg(&r->y, &r->y); }
static void p1_sub(p1_t *r, const p1_t *a, const p1_t *b) { p1_t nb; p1_neg(&nb, b); p1_add_or_double(r, a, &nb); }
/* G1::from_bytes + the "!is_inf && !is_valid" check of verify_kzg_proof_rust (eip_4844.rs:601-606) */
static int g1_from_bytes_valid(p1_t *out, const uint8_t in[48]) {
    p1_affine_t a;
    if (ko_p1_uncompress(&a, in)) return 1;
    p1_from_affine(out, &a);
    if (!p1_is_inf(out) && !ko_p1_in_g1(out)) return 1;
    return 0;
}
/* check_proof_single (blst/src/types/kzg_settings.rs:178-196): e(C - [y]G1, G2) == e(proof, [s]G2 - [z]G2) */
static int check_proof_single(const p1_t *com, const p1_t *proof, const fr_t *z, const fr_t *y, const kzg_settings_t *s) {
    g2_init();
    u64 zc[4];
    fr_to_canon(zc, z);
    p2_t g2, zg2, sg2, smz;
    p2_from_affine(&g2, &G2_GEN);
    p2_mult_canon(&zg2, &g2, zc, 255);
    fp2_neg(&zg2.y, &zg2.y);
    p2_from_affine(&sg2, &s->g2_monomial[1]);
    p2_add_or_double(&smz, &sg2, &zg2);
    p2_affine_t smz_a; p2_to_affine(&smz_a, &smz);
    p1_t g1, yg1, cmy;
    g1_generator(&g1);
    ko_p1_mult(&yg1, &g1, y);
    p1_sub(&cmy, com, &yg1);
    return pairings_verify(&cmy, &G2_GEN, proof, &smz_a);
}
/* verify_kzg_proof_raw (eip_4844.rs:612-636) */
API int ko_verify_kzg_proof(int *ok, const uint8_t commitment[48], const uint8_t z_bytes[32], const uint8_t y_bytes[32],
                            const uint8_t proof[48], void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    p1_affine_t ca, pa; p1_t c, p; fr_t z, y;
    if (ko_p1_uncompress(&ca, commitment) || ko_fr_from_bendian(&z, z_bytes) || ko_fr_from_bendian(&y, y_bytes) ||
        ko_p1_uncompress(&pa, proof)) return 1;
    p1_from_affine(&c, &ca); p1_from_affine(&p, &pa);
    if (!p1_is_inf(&c) && !ko_p1_in_g1(&c)) return 1;
    if (!p1_is_inf(&p) && !ko_p1_in_g1(&p)) return 1;
    *ok = check_proof_single(&c, &p, &z, &y, s);
    return 0;
}
/* verify_blob_kzg_proof_raw (eip_4844.rs:638-698) */
API int ko_verify_blob_kzg_proof(int *ok, const uint8_t *blob, const uint8_t commitment[48], const uint8_t proof[48], void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    fr_t *poly = (fr_t *)malloc(FIELD_ELEMENTS_PER_BLOB * sizeof(fr_t));
    p1_affine_t ca, pa; p1_t c, p; fr_t z, y;
    int rc = 1;
    if (!bytes_to_blob(poly, blob) && !ko_p1_uncompress(&ca, commitment) && !ko_p1_uncompress(&pa, proof)) {
        p1_from_affine(&c, &ca); p1_from_affine(&p, &pa);
        if ((p1_is_inf(&c) || ko_p1_in_g1(&c)) && (p1_is_inf(&p) || ko_p1_in_g1(&p))) {
            compute_challenge(&z, poly, &c);
            if (!evaluate_polynomial_in_evaluation_form(&y, poly, &z, s)) {
                *ok = check_proof_single(&c, &p, &z, &y, s);
                rc = 0;
            }
        }
    }
    free(poly);
    return rc;
}
/* compute_r_powers + verify_kzg_proof_batch (eip_4844.rs:328-435) */
static int verify_kzg_proof_batch(const p1_t *cs, const fr_t *zs, const fr_t *ys, const p1_t *ps, size_t n, const kzg_settings_t *s) {
    g2_init();
    size_t len = 32 + n * (48 + 32 + 32 + 48);
    uint8_t *buf = (uint8_t *)calloc(1, len);
    memcpy(buf, "RCKZGBATCH___V1_", 16);
    for (int i = 0; i < 8; i++) { buf[16 + i] = (uint8_t)((u64)FIELD_ELEMENTS_PER_BLOB >> (8 * (7 - i))); buf[24 + i] = (uint8_t)((u64)n >> (8 * (7 - i))); }
    for (size_t i = 0; i < n; i++) {
        uint8_t *o = buf + 32 + i * 160;
        ko_p1_compress(o, &cs[i]); ko_fr_to_bendian(o + 48, &zs[i]); ko_fr_to_bendian(o + 80, &ys[i]); ko_p1_compress(o + 112, &ps[i]);
    }
    uint8_t d[32]; fr_t r;
    ko_sha256(d, buf, len);
    ko_fr_from_bendian_unchecked(&r, d);
    free(buf);
    fr_t *rp = (fr_t *)malloc(2 * n * sizeof(fr_t)), *rz = rp + n;
    p1_t *cmy = (p1_t *)malloc(n * sizeof(p1_t));
    rp[0] = FR_ONE;
    for (size_t i = 1; i < n; i++) fr_mul(&rp[i], &rp[i - 1], &r);
    p1_t g1, proof_lincomb, proof_z_lincomb, cmy_lincomb, rhs;
    g1_generator(&g1);
    ko_msm_naive(&proof_lincomb, ps, rp, n);
    for (size_t i = 0; i < n; i++) {
        p1_t yg; ko_p1_mult(&yg, &g1, &ys[i]);
        p1_sub(&cmy[i], &cs[i], &yg);
        fr_mul(&rz[i], &rp[i], &zs[i]);
    }
    ko_msm_naive(&proof_z_lincomb, ps, rz, n);
    ko_msm_naive(&cmy_lincomb, cmy, rp, n);
    p1_add_or_double(&rhs, &cmy_lincomb, &proof_z_lincomb);
    int ok = pairings_verify(&proof_lincomb, &s->g2_monomial[1], &rhs, &G2_GEN);
    free(rp); free(cmy);
    return ok;
}
/* verify_blob_kzg_proof_batch_raw (eip_4844.rs:728-866), the non-"parallel" branch */
API int ko_verify_blob_kzg_proof_batch(int *ok, const uint8_t *blobs, const uint8_t *commitments, const uint8_t *proofs,
                                       size_t n, void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    if (n == 0) { *ok = 1; return 0; }
    if (n == 1) return ko_verify_blob_kzg_proof(ok, blobs, commitments, proofs, h);
    fr_t *polys = (fr_t *)malloc(n * FIELD_ELEMENTS_PER_BLOB * sizeof(fr_t));
    p1_t *cs = (p1_t *)malloc(2 * n * sizeof(p1_t)), *ps = cs + n;
    fr_t *zs = (fr_t *)malloc(2 * n * sizeof(fr_t)), *ys = zs + n;
    int rc = 1;
    for (size_t i = 0; i < n; i++) if (bytes_to_blob(polys + i * FIELD_ELEMENTS_PER_BLOB, blobs + i * BYTES_PER_BLOB)) goto done;
    for (size_t i = 0; i < n; i++) { p1_affine_t a; if (ko_p1_uncompress(&a, commitments + 48 * i)) goto done; p1_from_affine(&cs[i], &a); }
    for (size_t i = 0; i < n; i++) { p1_affine_t a; if (ko_p1_uncompress(&a, proofs + 48 * i)) goto done; p1_from_affine(&ps[i], &a); }
    for (size_t i = 0; i < n; i++) if (!p1_is_inf(&cs[i]) && !ko_p1_in_g1(&cs[i])) goto done;
    for (size_t i = 0; i < n; i++) if (!p1_is_inf(&ps[i]) && !ko_p1_in_g1(&ps[i])) goto done;
    for (size_t i = 0; i < n; i++) {
        const fr_t *poly = polys + i * FIELD_ELEMENTS_PER_BLOB;
        compute_challenge(&zs[i], poly, &cs[i]);
        if (evaluate_polynomial_in_evaluation_form(&ys[i], poly, &zs[i], s)) goto done;
    }
    *ok = verify_kzg_proof_batch(cs, zs, ys, ps, n, s);
    rc = 0;
done:
    free(polys); free(cs); free(zs);
    return rc;
}
API const p2_affine_t *ko_settings_g2_monomial(void *h) { return ((kzg_settings_t *)h)->g2_monomial; }

/* ------------------------------------------------------------------------------------------------------------ */
/* EIP-7594 recovery and cell verification (kzg/src/das.rs:101-207, 294-452, 454-616, 697-900).
 * Return 0 = Ok, 1 = Err (C ABI: BadArgs). */
#define CELLS_PER_EXT_BLOB 128
#define BYTES_PER_CELL 2048
static int cells_to_fr(fr_t *out, const uint8_t *cells, size_t ncells) {
    for (size_t i = 0; i < ncells * CELL_SIZE; i++)
        if (ko_fr_from_bendian(&out[i], cells + 32 * i)) return 1;
    return 0;
}
static void shift_poly(fr_t *poly, size_t n, const fr_t *factor) {  /* das.rs:454-460 */
    fr_t fp = FR_ONE;
    for (size_t i = 1; i < n; i++) { fr_mul(&fp, &fp, factor); fr_mul(&poly[i], &poly[i], &fp); }
}
/* recover_cells (das.rs:549-616); present[i] marks provided cells, data = 8192 evaluations in cell order */
static int recover_cells(fr_t *data, const uint8_t *present, const kzg_settings_t *s) {
    const size_t N = 8192;
    fr_t *brp = (fr_t *)malloc(N * sizeof(fr_t)), *van = (fr_t *)calloc(N, sizeof(fr_t)), *van_eval = (fr_t *)malloc(N * sizeof(fr_t));
    fr_t *ez = (fr_t *)malloc(N * sizeof(fr_t)), *t = (fr_t *)malloc(N * sizeof(fr_t)), *vc = (fr_t *)malloc(N * sizeof(fr_t));
    uint8_t *present_brp = (uint8_t *)malloc(N);
    for (size_t i = 0; i < N; i++) { size_t r = brp_index(i, 13); brp[r] = data[i]; present_brp[r] = present[i / CELL_SIZE]; }
    size_t missing[CELLS_PER_EXT_BLOB], nm = 0;
    for (size_t i = 0; i < CELLS_PER_EXT_BLOB; i++) if (!present[i]) missing[nm++] = brp_index(i, 7);
    int rc = 1;
    if (nm > CELLS_PER_EXT_BLOB / 2 || nm == 0) goto done;
    {
        /* vanishing_polynomial_for_missing_cells (:516-547) with compute_vanishing_polynomial_from_roots (:488-514) */
        fr_t shortp[CELLS_PER_EXT_BLOB + 1], neg, zero;
        memset(&zero, 0, sizeof(zero));
        const size_t stride = N / CELLS_PER_EXT_BLOB;
        fr_sub(&shortp[0], &zero, &s->fs->roots_of_unity[missing[0] * stride]);
        for (size_t i = 1; i < nm; i++) {
            fr_sub(&neg, &zero, &s->fs->roots_of_unity[missing[i] * stride]);
            fr_add(&shortp[i], &neg, &shortp[i - 1]);
            for (size_t j = i - 1; j >= 1; j--) { fr_mul(&shortp[j], &shortp[j], &neg); fr_add(&shortp[j], &shortp[j], &shortp[j - 1]); }
            fr_mul(&shortp[0], &shortp[0], &neg);
        }
        shortp[nm] = FR_ONE;
        for (size_t i = 0; i <= nm; i++) van[i * CELL_SIZE] = shortp[i];
    }
    ko_fft_fr(s->fs, van_eval, van, N, 0, 1);
    for (size_t i = 0; i < N; i++) {
        if (!present_brp[i]) memset(&ez[i], 0, sizeof(fr_t)); else fr_mul(&ez[i], &brp[i], &van_eval[i]);
    }
    ko_fft_fr(s->fs, t, ez, N, 1, 1);
    {
        fr_t seven, inv7;
        fr_from_u64(&seven, 7);
        fr_inv(&inv7, &seven);
        shift_poly(t, N, &seven);                       /* coset_fft (:462-473) */
        ko_fft_fr(s->fs, ez, t, N, 0, 1);
        memcpy(vc, van, N * sizeof(fr_t));
        shift_poly(vc, N, &seven);
        ko_fft_fr(s->fs, t, vc, N, 0, 1);
        for (size_t i = 0; i < N; i++) { fr_inv(&t[i], &t[i]); fr_mul(&ez[i], &ez[i], &t[i]); }
        ko_fft_fr(s->fs, t, ez, N, 1, 1);               /* coset_ifft (:475-486) */
        shift_poly(t, N, &inv7);
    }
    ko_fft_fr(s->fs, ez, t, N, 0, 1);
    for (size_t i = 0; i < N; i++) data[brp_index(i, 13)] = ez[i];
    rc = 0;
done:
    free(brp); free(van); free(van_eval); free(ez); free(t); free(vc); free(present_brp);
    return rc;
}
/* recover_cells_and_kzg_proofs (das.rs:101-207) behind the byte-level wrapper (eth/c_bindings.rs:201-286) */
API int ko_recover_cells_and_kzg_proofs(uint8_t *cells_out, uint8_t *proofs_out, const u64 *cell_indices, const uint8_t *cells,
                                        size_t n, void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    const size_t N = 8192;
    if (n > CELLS_PER_EXT_BLOB) return 1;                       /* "Cell length cannot be larger than CELLS_PER_EXT_BLOB" */
    fr_t *in = (fr_t *)malloc((n ? n : 1) * CELL_SIZE * sizeof(fr_t)), *data = (fr_t *)calloc(N, sizeof(fr_t));
    uint8_t present[CELLS_PER_EXT_BLOB] = {0};
    int rc = 1;
    if (cells_to_fr(in, cells, n)) goto done;
    if (n < CELLS_PER_EXT_BLOB / 2) goto done;                  /* "Impossible to recover" */
    for (size_t i = 0; i < n; i++) {
        if (cell_indices[i] >= CELLS_PER_EXT_BLOB) goto done;
        if (i + 1 < n && cell_indices[i + 1] <= cell_indices[i]) goto done;   /* strictly ascending */
        present[cell_indices[i]] = 1;
        memcpy(data + cell_indices[i] * CELL_SIZE, in + i * CELL_SIZE, CELL_SIZE * sizeof(fr_t));
    }
    if (n != CELLS_PER_EXT_BLOB && recover_cells(data, present, s)) goto done;
    for (size_t i = 0; i < N; i++) ko_fr_to_bendian(cells_out + 32 * i, &data[i]);
    if (proofs_out) {
        fr_t *brp = (fr_t *)malloc(N * sizeof(fr_t)), *mono = (fr_t *)malloc(N * sizeof(fr_t));
        for (size_t i = 0; i < N; i++) brp[brp_index(i, 13)] = data[i];
        ko_fft_fr(s->fs, mono, brp, N, 1, 1);                   /* poly_lagrange_to_monomial on all 8192 (:186-188) */
        fk20_proofs_from_mono(proofs_out, mono, s);
        free(brp); free(mono);
    }
    rc = 0;
done:
    free(in); free(data);
    return rc;
}
/* compute_verify_cell_kzg_proof_batch_challenge (das.rs:390-452); inputs already validated */
static void cell_batch_challenge(fr_t *out, const uint8_t *commitments48, size_t ncomm, const u64 *commitment_indices,
                                 const u64 *cell_indices, const uint8_t *cells, const uint8_t *proofs48, size_t n) {
    size_t len = 16 + 32 + ncomm * 48 + n * (16 + BYTES_PER_CELL + 48);
    uint8_t *buf = (uint8_t *)malloc(len), *p = buf;
    memcpy(p, "RCKZGCBATCH__V1_", 16); p += 16;
    u64 head[4] = {FIELD_ELEMENTS_PER_BLOB, CELL_SIZE, ncomm, n};
    for (int k = 0; k < 4; k++) for (int i = 0; i < 8; i++) *p++ = (uint8_t)(head[k] >> (8 * (7 - i)));
    memcpy(p, commitments48, ncomm * 48); p += ncomm * 48;
    for (size_t c = 0; c < n; c++) {
        for (int i = 0; i < 8; i++) *p++ = (uint8_t)(commitment_indices[c] >> (8 * (7 - i)));
        for (int i = 0; i < 8; i++) *p++ = (uint8_t)(cell_indices[c] >> (8 * (7 - i)));
        memcpy(p, cells + c * BYTES_PER_CELL, BYTES_PER_CELL); p += BYTES_PER_CELL;
        memcpy(p, proofs48 + 48 * c, 48); p += 48;
    }
    uint8_t d[32];
    ko_sha256(d, buf, len);
    ko_fr_from_bendian_unchecked(out, d);
    free(buf);
}
/* blst/src/eip_7594.rs:35-97: every commitment, cell element and proof must decode */
API int ko_compute_verify_cell_kzg_proof_batch_challenge(uint8_t out32[32], const uint8_t *commitments48, size_t ncomm,
                                                         const u64 *commitment_indices, const u64 *cell_indices,
                                                         const uint8_t *cells, const uint8_t *proofs48, size_t n) {
    p1_affine_t a; fr_t f, r;
    for (size_t i = 0; i < ncomm; i++) if (ko_p1_uncompress(&a, commitments48 + 48 * i)) return 1;
    for (size_t i = 0; i < n * CELL_SIZE; i++) if (ko_fr_from_bendian(&f, cells + 32 * i)) return 1;
    for (size_t i = 0; i < n; i++) if (ko_p1_uncompress(&a, proofs48 + 48 * i)) return 1;
    cell_batch_challenge(&r, commitments48, ncomm, commitment_indices, cell_indices, cells, proofs48, n);
    ko_fr_to_bendian(out32, &r);
    return 0;
}
/* verify_cell_kzg_proof_batch (das.rs:294-388) behind the byte-level wrapper (eth/c_bindings.rs:288-352) */
API int ko_verify_cell_kzg_proof_batch(int *ok, const uint8_t *commitments48, const u64 *cell_indices, const uint8_t *cells,
                                       const uint8_t *proofs48, size_t n, void *h) {
    kzg_settings_t *s = (kzg_settings_t *)h;
    g2_init();
    size_t nn = n ? n : 1;
    p1_t *cs = (p1_t *)malloc(nn * sizeof(p1_t)), *ps = (p1_t *)malloc(nn * sizeof(p1_t)), *uc = (p1_t *)malloc(nn * sizeof(p1_t));
    fr_t *fr = (fr_t *)malloc(nn * CELL_SIZE * sizeof(fr_t)), *rp = (fr_t *)malloc(nn * sizeof(fr_t)), *w = (fr_t *)calloc(nn, sizeof(fr_t));
    u64 *cidx = (u64 *)malloc(nn * sizeof(u64));
    uint8_t *ucb = (uint8_t *)malloc(nn * 48);
    fr_t *agg = (fr_t *)calloc(CELLS_PER_EXT_BLOB * CELL_SIZE, sizeof(fr_t));
    int rc = 1;
    for (size_t i = 0; i < n; i++) { p1_affine_t a; if (ko_p1_uncompress(&a, commitments48 + 48 * i)) goto done; p1_from_affine(&cs[i], &a); }
    if (cells_to_fr(fr, cells, n)) goto done;
    for (size_t i = 0; i < n; i++) { p1_affine_t a; if (ko_p1_uncompress(&a, proofs48 + 48 * i)) goto done; p1_from_affine(&ps[i], &a); }
    if (n == 0) { *ok = 1; rc = 0; goto done; }
    for (size_t i = 0; i < n; i++) if (cell_indices[i] >= CELLS_PER_EXT_BLOB) goto done;
    for (size_t i = 0; i < n; i++) if (!p1_is_inf(&ps[i]) && !ko_p1_in_g1(&ps[i])) goto done;
    /* deduplicate_with_indices (:57-76) */
    size_t nu = 0;
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (; j < nu; j++) if (memcmp(ucb + 48 * j, commitments48 + 48 * i, 48) == 0) break;
        if (j == nu) { memcpy(ucb + 48 * nu, commitments48 + 48 * i, 48); uc[nu] = cs[i]; nu++; }
        cidx[i] = j;
    }
    for (size_t j = 0; j < nu; j++) if (!p1_is_inf(&uc[j]) && !ko_p1_in_g1(&uc[j])) goto done;
    {
        fr_t r;
        cell_batch_challenge(&r, ucb, nu, cidx, cell_indices, cells, proofs48, n);
        rp[0] = FR_ONE;
        for (size_t i = 1; i < n; i++) fr_mul(&rp[i], &rp[i - 1], &r);
        p1_t proof_lincomb, final_sum, interp, wsum;
        ko_msm_naive(&proof_lincomb, ps, rp, n);
        /* compute_weighted_sum_of_commitments (:697-741) */
        for (size_t i = 0; i < n; i++) fr_add(&w[cidx[i]], &w[cidx[i]], &rp[i]);
        ko_msm_naive(&final_sum, uc, w, nu);
        /* compute_commitment_to_aggregated_interpolation_poly (:780-842) */
        uint8_t used[CELLS_PER_EXT_BLOB] = {0};
        for (size_t i = 0; i < n; i++) {
            used[cell_indices[i]] = 1;
            for (size_t k = 0; k < CELL_SIZE; k++) {
                fr_t t; fr_mul(&t, &fr[i * CELL_SIZE + k], &rp[i]);
                fr_add(&agg[cell_indices[i] * CELL_SIZE + k], &agg[cell_indices[i] * CELL_SIZE + k], &t);
            }
        }
        fr_t poly[CELL_SIZE], col[CELL_SIZE], colp[CELL_SIZE];
        memset(poly, 0, sizeof(poly));
        for (size_t c = 0; c < CELLS_PER_EXT_BLOB; c++) {
            if (!used[c]) continue;
            for (size_t k = 0; k < CELL_SIZE; k++) col[brp_index(k, 6)] = agg[c * CELL_SIZE + k];
            ko_fft_fr(s->fs, colp, col, CELL_SIZE, 1, 1);
            /* get_inv_coset_shift_for_cell (:743-778): roots_of_unity[8192 - brp7(c)] */
            shift_poly(colp, CELL_SIZE, &s->fs->roots_of_unity[8192 - brp_index(c, 7)]);
            for (size_t k = 0; k < CELL_SIZE; k++) fr_add(&poly[k], &poly[k], &colp[k]);
        }
        ko_msm_naive(&interp, s->g1_monomial, poly, CELL_SIZE);
        p1_sub(&final_sum, &final_sum, &interp);
        /* computed_weighted_sum_of_proofs (:844-880): weights r^i * h_k^64 = roots_of_unity[brp7(cell) * 64] */
        for (size_t i = 0; i < n; i++) fr_mul(&w[i], &rp[i], &s->fs->roots_of_unity[brp_index(cell_indices[i], 7) * CELL_SIZE]);
        ko_msm_naive(&wsum, ps, w, n);
        p1_add_or_double(&final_sum, &final_sum, &wsum);
        *ok = pairings_verify(&final_sum, &G2_GEN, &proof_lincomb, &s->g2_monomial[CELL_SIZE]);
        rc = 0;
    }
done:
    free(cs); free(ps); free(uc); free(fr); free(rp); free(w); free(cidx); free(ucb); free(agg);
    return rc;
}
