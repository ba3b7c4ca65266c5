// g1.cuh -- BLS12-381 G1 point arithmetic on the device (y^2 = x^3 + 4 over Fp).
//
// Three representations, all Montgomery-form coordinates with blst's memory layouts so they cross the C ABI as is:
//   affine_t  (x, y)            96 B   blst_p1_affine; infinity <=> all zero   (blst/src/types/g1.rs:303-316)
//   jac_t     (x, y, z)        144 B   blst_p1 Jacobian; infinity <=> z == 0   (blst/src/types/g1.rs:151-177)
//   xyzz_t    (x, y, zzz, zz)  192 B   bucket type: X/ZZ, Y/ZZZ with ZZ^3 = ZZZ^2; infinity <=> zz == 0.
//                                      Same field order as the reference's P1XYZZ (kzg/src/msm/pippenger_utils.rs:5-12).
// The exceptional cases of the addition law (either operand infinity, P + P, P + (-P)) are all handled, as the
// reference's p1_dadd_affine / p1_dadd do (kzg/src/msm/pippenger_utils.rs:90-210): blobs are adversarial inputs.
#pragma once
#include "mont.cuh"

namespace b200 {

struct affine_t {
    fp_t x, y;
    __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
struct jac_t {
    fp_t x, y, z;
    __device__ __forceinline__ bool is_inf() const { return z.is_zero(); }
    static __device__ __forceinline__ jac_t inf() { return jac_t{fp_t::zero(), fp_t::zero(), fp_t::zero()}; }
};
struct xyzz_t {
    fp_t x, y, zzz, zz;
    __device__ __forceinline__ bool is_inf() const { return zz.is_zero(); }
    static __device__ __forceinline__ xyzz_t inf() { return xyzz_t{fp_t::zero(), fp_t::zero(), fp_t::zero(), fp_t::zero()}; }
};

__device__ __forceinline__ affine_t load_affine(const void* p) {
    affine_t a;
    a.x = load_field_ro<fp_t>(p);
    a.y = load_field_ro<fp_t>(reinterpret_cast<const char*>(p) + 48);
    return a;
}
__device__ __forceinline__ void store_affine(void* p, const affine_t& a) {
    store_field(p, a.x);
    store_field(reinterpret_cast<char*>(p) + 48, a.y);
}
__device__ __forceinline__ xyzz_t load_xyzz(const void* p) {
    const char* q = reinterpret_cast<const char*>(p);
    xyzz_t r;
    r.x = load_field<fp_t>(q); r.y = load_field<fp_t>(q + 48); r.zzz = load_field<fp_t>(q + 96); r.zz = load_field<fp_t>(q + 144);
    return r;
}
__device__ __forceinline__ void store_xyzz(void* p, const xyzz_t& a) {
    char* q = reinterpret_cast<char*>(p);
    store_field(q, a.x); store_field(q + 48, a.y); store_field(q + 96, a.zzz); store_field(q + 144, a.zz);
}
__device__ __forceinline__ jac_t load_jac(const void* p) {
    const char* q = reinterpret_cast<const char*>(p);
    jac_t r;
    r.x = load_field<fp_t>(q); r.y = load_field<fp_t>(q + 48); r.z = load_field<fp_t>(q + 96);
    return r;
}
__device__ __forceinline__ void store_jac(void* p, const jac_t& a) {
    char* q = reinterpret_cast<char*>(p);
    store_field(q, a.x); store_field(q + 48, a.y); store_field(q + 96, a.z);
}

// XYZZ <- 2 * affine   (mdbl-2008-s-1)
__device__ __forceinline__ xyzz_t xyzz_dbl_affine(const affine_t& p) {
    xyzz_t r;
    fp_t u = p.y.dbl();
    r.zz = u.sqr();
    r.zzz = r.zz * u;
    fp_t s = p.x * r.zz;
    fp_t m = p.x.sqr();
    m = m.dbl() + m;
    r.x = m.sqr() - s.dbl();
    r.y = m * (s - r.x) - r.zzz * p.y;
    return r;
}

// acc += p  (p affine, non-infinity handled by caller for speed; full exceptional-case handling inside)
// madd-2008-s: 8M + 2S
__device__ __forceinline__ void xyzz_add_affine(xyzz_t& acc, const affine_t& p) {
    if (p.is_inf()) return;
    if (acc.is_inf()) {
        acc.x = p.x; acc.y = p.y; acc.zz = fp_t::one(); acc.zzz = fp_t::one();
        return;
    }
    fp_t P = p.x * acc.zz - acc.x;
    fp_t R = p.y * acc.zzz - acc.y;
    if (!P.is_zero()) {
        fp_t PP = P.sqr();
        fp_t PPP = PP * P;
        fp_t Q = acc.x * PP;
        fp_t x3 = R.sqr() - PPP - Q.dbl();
        acc.y = R * (Q - x3) - acc.y * PPP;
        acc.x = x3;
        acc.zz = acc.zz * PP;
        acc.zzz = acc.zzz * PPP;
    } else if (R.is_zero()) {
        acc = xyzz_dbl_affine(p);
    } else {
        acc = xyzz_t::inf();
    }
}

// acc += q  (both XYZZ), add-2008-s: 12M + 2S; doubling via dbl-2008-s-1
__device__ __forceinline__ void xyzz_add(xyzz_t& acc, const xyzz_t& q) {
    if (q.is_inf()) return;
    if (acc.is_inf()) { acc = q; return; }
    fp_t U1 = acc.x * q.zz;
    fp_t S1 = acc.y * q.zzz;
    fp_t P = q.x * acc.zz - U1;
    fp_t R = q.y * acc.zzz - S1;
    if (!P.is_zero()) {
        fp_t PP = P.sqr();
        fp_t PPP = PP * P;
        fp_t Q = U1 * PP;
        fp_t x3 = R.sqr() - PPP - Q.dbl();
        acc.y = R * (Q - x3) - S1 * PPP;
        acc.x = x3;
        acc.zz = acc.zz * q.zz * PP;
        acc.zzz = acc.zzz * q.zzz * PPP;
    } else if (R.is_zero()) {
        // acc == q: double (dbl-2008-s-1)
        fp_t U = acc.y.dbl();
        fp_t V = U.sqr();
        fp_t W = U * V;
        fp_t S = acc.x * V;
        fp_t M = acc.x.sqr();
        M = M.dbl() + M;
        fp_t x3 = M.sqr() - S.dbl();
        acc.y = M * (S - x3) - W * acc.y;
        acc.x = x3;
        acc.zz = V * acc.zz;
        acc.zzz = W * acc.zzz;
    } else {
        acc = xyzz_t::inf();
    }
}
// out-of-line copies for cold paths (keeps the hot kernels' code size and register allocation focused)
static __device__ __noinline__ void xyzz_add_noinline(xyzz_t& acc, const xyzz_t& q) { xyzz_add(acc, q); }

// acc = 2 * acc
__device__ __forceinline__ void xyzz_dbl(xyzz_t& acc) {
    if (acc.is_inf()) return;
    fp_t U = acc.y.dbl();
    if (U.is_zero()) { acc = xyzz_t::inf(); return; }  // order-2 point: cannot occur in the prime-order subgroup
    fp_t V = U.sqr();
    fp_t W = U * V;
    fp_t S = acc.x * V;
    fp_t M = acc.x.sqr();
    M = M.dbl() + M;
    fp_t x3 = M.sqr() - S.dbl();
    acc.y = M * (S - x3) - W * acc.y;
    acc.x = x3;
    acc.zz = V * acc.zz;
    acc.zzz = W * acc.zzz;
}

// XYZZ -> Jacobian with Z = ZZZ/ZZ?  No division needed: (X*ZZ, Y*ZZZ, ZZ) is the same point in Jacobian
// coordinates with Z := ZZ (the reference's p1_to_jacobian, kzg/src/msm/pippenger_utils.rs:84-88):
//   X_j / Z^2 = X*ZZ / ZZ^2 = X/ZZ,   Y_j / Z^3 = Y*ZZZ / ZZ^3 = Y*ZZZ/ZZZ^2 = Y/ZZZ.
__device__ __forceinline__ jac_t xyzz_to_jac(const xyzz_t& p) {
    if (p.is_inf()) return jac_t::inf();
    return jac_t{p.x * p.zz, p.y * p.zzz, p.zz};
}
__device__ __forceinline__ xyzz_t jac_to_xyzz(const jac_t& p) {
    if (p.is_inf()) return xyzz_t::inf();
    fp_t zz = p.z.sqr();
    return xyzz_t{p.x, p.y, zz * p.z, zz};
}
__device__ __forceinline__ xyzz_t affine_to_xyzz(const affine_t& p) {
    if (p.is_inf()) return xyzz_t::inf();
    return xyzz_t{p.x, p.y, fp_t::one(), fp_t::one()};
}
// XYZZ -> affine (one field inversion)
__device__ __forceinline__ affine_t xyzz_to_affine(const xyzz_t& p) {
    if (p.is_inf()) return affine_t{fp_t::zero(), fp_t::zero()};
    // 1/ZZZ; then 1/ZZ = ZZZ^-2 * ZZ^2  (ZZ^3 = ZZZ^2)
    fp_t izzz = p.zzz.inverse();
    fp_t izz = izzz.sqr() * p.zz.sqr();
    return affine_t{p.x * izz, p.y * izzz};
}
__device__ __forceinline__ affine_t jac_to_affine(const jac_t& p) {
    if (p.is_inf()) return affine_t{fp_t::zero(), fp_t::zero()};
    fp_t zi = p.z.inverse();
    fp_t zi2 = zi.sqr();
    return affine_t{p.x * zi2, p.y * zi2 * zi};
}

// y > (p-1)/2 on the canonical value: the "lexicographically largest" flag of the compressed encoding
// (zkcrypto/bls12_381/src/notes/serialization.rs:18-29)
__device__ __forceinline__ bool fp_is_lex_largest(const fp_t& y_mont) {
    fp_t c = y_mont.from_mont();
    // (p-1)/2, little-endian u32 limbs
    const uint32_t H[12] = {0xffffd555u, 0xdcff7fffu, 0x58a9ffffu, 0x0f55ffffu, 0x7b587b12u, 0xb3986950u,
                            0x79c2895fu, 0xb23ba5c2u, 0x21a5d66bu, 0x258dd3dbu, 0x1cbff34du, 0x0d0088f5u};
#pragma unroll
    for (int i = 11; i >= 0; i--) {
        if (c.v[i] > H[i]) return true;
        if (c.v[i] < H[i]) return false;
    }
    return false;
}

// 48-byte compressed encoding of an affine point (blst_p1_compress, blst/src/types/g1.rs:94-100)
__device__ __forceinline__ void affine_compress(uint8_t* out, const affine_t& a) {
    if (a.is_inf()) {
        out[0] = 0xC0;
        for (int i = 1; i < 48; i++) out[i] = 0;
        return;
    }
    fp_t c = a.x.from_mont();
    for (int i = 0; i < 12; i++) {
        uint32_t w = c.v[11 - i];
        out[4 * i] = w >> 24; out[4 * i + 1] = w >> 16; out[4 * i + 2] = w >> 8; out[4 * i + 3] = w;
    }
    out[0] |= 0x80;
    if (fp_is_lex_largest(a.y)) out[0] |= 0x20;
}

}  // namespace b200
