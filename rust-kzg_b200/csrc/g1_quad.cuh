// g1_quad.cuh -- lane-parallel XYZZ arithmetic for the latency-bound reduction tails.
//
// Measured on B200 (scripts/ubench/lat.cu): ONE warp needs 13.4 us for an XYZZ addition and 8.7 us for a doubling --
// an Fp multiplication is ~290 IMAD.WIDE, which issue at 8 lanes/clk per SM sub-partition, so one warp alone is
// already bound by its scheduler's FMA-heavy pipe and the 14 (9) multiplications of the formula run back to back.
// The tails (bucket combine, marginal sums, group finish) are trees in which most lanes hold nothing useful, so the
// idle lanes are put to work: a point is spread over a QUAD of lanes,
//
//      lane 4k+0: X      lane 4k+1: Y      lane 4k+2: ZZ      lane 4k+3: ZZZ        ("role" = lane & 3)
//
// and the formulas are scheduled so that every multiplication LEVEL is one SIMT multiplication in which the four
// lanes compute four different products.  add-2008-s (12M + 2S) becomes 4 levels, dbl-2008-s-1 (6M + 3S) becomes 3;
// operands move between the lanes of a quad with warp shuffles (8 and 6 field-element shuffles).
//
// All functions here must be called by ALL 32 lanes of a warp (full-mask shuffles); quads whose inputs are
// meaningless just compute garbage.  Exceptional cases follow g1_body.inc's xyzz_add / xyzz_dbl exactly: infinity
// operands, P == Q (falls back to the quad doubling), P == -Q (infinity).
#pragma once
#include "g1.cuh"

namespace b200 {

static constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ fp_t shfl_idx_fp(const fp_t& a, int src) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = __shfl_sync(kFullMask, a.v[i], src);
    return r;
}
__device__ __forceinline__ fp_t shfl_xor_fp(const fp_t& a, int m) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = __shfl_xor_sync(kFullMask, a.v[i], m);
    return r;
}
__device__ __forceinline__ fp_t shfl_dn_fp(const fp_t& a, int d) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = __shfl_down_sync(kFullMask, a.v[i], d);
    return r;
}
__device__ __forceinline__ xyzz_t shfl_down_xyzz(const xyzz_t& v, int d) {
    xyzz_t o;
    o.x = shfl_dn_fp(v.x, d); o.y = shfl_dn_fp(v.y, d); o.zzz = shfl_dn_fp(v.zzz, d); o.zz = shfl_dn_fp(v.zz, d);
    return o;
}

// this lane's component of the full point held by lane `src` (src may differ per quad)
__device__ __forceinline__ fp_t quad_scatter(const xyzz_t& p, int src) {
    const int role = threadIdx.x & 3;
    fp_t x = shfl_idx_fp(p.x, src), y = shfl_idx_fp(p.y, src), zz = shfl_idx_fp(p.zz, src), zzz = shfl_idx_fp(p.zzz, src);
    return role == 0 ? x : role == 1 ? y : role == 2 ? zz : zzz;
}
// the full point of this lane's quad, valid on the quad's first lane
__device__ __forceinline__ xyzz_t quad_gather(const fp_t& comp) {
    xyzz_t r;
    r.x = comp;
    r.y = shfl_dn_fp(comp, 1);
    r.zz = shfl_dn_fp(comp, 2);
    r.zzz = shfl_dn_fp(comp, 3);
    return r;
}
// byte offset of this lane's component inside a stored xyzz_t (x, y, zzz, zz -- see store_xyzz)
__device__ __forceinline__ int quad_store_offset() {
    const int role = threadIdx.x & 3;
    return role == 0 ? 0 : role == 1 ? 48 : role == 2 ? 144 : 96;
}

// The quad formulas run on one or a few warps per SM: their multiplications go through ONE out-of-line copy of the
// multiplier (B200_QUAD_INLINE_MUL restores inlining) so that a doubling + an addition are ~2 KiB of code instead of
// ~45 KiB -- with a lone warp per scheduler every instruction-cache miss is exposed latency.
#ifdef B200_QUAD_INLINE_MUL
__device__ __forceinline__ fp_t qmul(const fp_t& a, const fp_t& b) { return a * b; }
#else
static __device__ __noinline__ fp_t qmul(fp_t a, fp_t b) { return a * b; }
#endif

// 2 * a   (dbl-2008-s-1 in three multiplication levels)
static __device__ __noinline__ fp_t quad_dbl(fp_t a) {
    const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
    const bool a_inf = __shfl_sync(kFullMask, (int)a.is_zero(), base | 2);
    fp_t U = shfl_idx_fp(a.dbl(), base | 1);                       // U = 2Y on every lane
    // level 1: lane0 X^2 ; lanes 1..3 V = U^2
    fp_t m = role == 0 ? a : U;
    fp_t l1 = qmul(m, m);
    fp_t M = l1.dbl() + l1;                                        // lane0: M = 3 X^2
    fp_t V = shfl_idx_fp(l1, base | 1);
    fp_t Mb = shfl_idx_fp(M, base);
    // level 2: lane0 S = X V ; lane1 W = U V ; lane2 ZZ3 = ZZ V ; lane3 M^2
    fp_t m1 = role == 3 ? Mb : (role == 1 ? U : a);
    fp_t m2 = role == 3 ? Mb : V;
    fp_t l2 = qmul(m1, m2);
    fp_t MM = shfl_idx_fp(l2, base | 3);
    fp_t Wb = shfl_idx_fp(l2, base | 1);
    fp_t x3 = MM - l2.dbl();                                       // lane0: M^2 - 2S
    // level 3: lane0 M (S - X3) ; lane1 W Y ; lane3 ZZZ3 = W ZZZ ; lane2 idle
    m1 = role == 0 ? M : Wb;
    m2 = role == 0 ? (l2 - x3) : a;
    fp_t l3 = qmul(m1, m2);
    fp_t t = shfl_idx_fp(l3, base);
    fp_t y3 = t - l3;                                              // lane1
    fp_t r = role == 0 ? x3 : role == 1 ? y3 : role == 2 ? l2 : l3;
    if (U.is_zero()) r = fp_t::zero();                             // order-2 point: cannot occur in the subgroup
    return a_inf ? a : r;
}

// a + b   (add-2008-s in four multiplication levels)
static __device__ __noinline__ fp_t quad_add(fp_t a, fp_t b) {
    const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
    const bool a_inf = __shfl_sync(kFullMask, (int)a.is_zero(), base | 2);
    const bool b_inf = __shfl_sync(kFullMask, (int)b.is_zero(), base | 2);
    // level 1: own component of a times the opposite component of b: lane0 U1 = X1 ZZ2, lane1 S1 = Y1 ZZZ2,
    //          lane2 U2 = ZZ1 X2, lane3 S2 = ZZZ1 Y2
    fp_t l1 = qmul(a, shfl_xor_fp(b, 2));
    fp_t l1x = shfl_xor_fp(l1, 2);
    fp_t pr = role < 2 ? (l1x - l1) : (l1 - l1x);                  // lanes 0,2: P = U2 - U1 ; lanes 1,3: R = S2 - S1
    const bool p_zero = __shfl_sync(kFullMask, (int)pr.is_zero(), base);
    const bool r_zero = __shfl_sync(kFullMask, (int)pr.is_zero(), base | 1);
    // level 2: lane0 PP = P^2 ; lane1 RR = R^2 ; lane2 ZZ1 ZZ2 ; lane3 ZZZ1 ZZZ2
    fp_t m1 = role < 2 ? pr : a;
    fp_t m2 = role < 2 ? pr : b;
    fp_t l2 = qmul(m1, m2);
    fp_t pp = shfl_idx_fp(l2, base);
    fp_t u1 = shfl_idx_fp(l1, base);
    // level 3: lane0 PPP = P PP ; lane1 Q = U1 PP ; lane2 ZZ3 = (ZZ1 ZZ2) PP ; lane3 idle
    m1 = role == 0 ? pr : role == 1 ? u1 : l2;
    fp_t l3 = qmul(m1, pp);
    fp_t ppp = shfl_idx_fp(l3, base);
    fp_t s1 = shfl_idx_fp(l1, base | 1);
    fp_t x3 = l2 - ppp - l3.dbl();                                 // lane1: RR - PPP - 2Q
    // level 4: lane0 S1 PPP ; lane1 R (Q - X3) ; lane3 ZZZ3 = (ZZZ1 ZZZ2) PPP ; lane2 idle
    m1 = role == 0 ? s1 : role == 1 ? pr : l2;
    m2 = role == 1 ? (l3 - x3) : ppp;
    fp_t l4 = qmul(m1, m2);
    fp_t t0 = shfl_idx_fp(l4, base);
    fp_t y3 = l4 - t0;                                             // lane1
    fp_t x3_0 = shfl_idx_fp(x3, base | 1);
    fp_t r = role == 0 ? x3_0 : role == 1 ? y3 : role == 2 ? l3 : l4;
    // exceptional cases, in xyzz_add's order
    const bool same_x = !a_inf && !b_inf && p_zero;
    const bool need_dbl = same_x && r_zero;
    if (__any_sync(kFullMask, need_dbl)) {                         // warp-uniform: P == Q somewhere, rare
        fp_t d = quad_dbl(a);
        if (need_dbl) r = d;
    }
    if (same_x && !r_zero) r = fp_t::zero();                       // P == -Q
    if (a_inf) r = b;
    if (b_inf) r = a;
    return r;
}

// Sum over aligned segments of S lanes (S = 1, 2, 4, 8, 16, 32), each lane holding a full point; the result is valid
// on the first lane of every segment.  Level 1 runs as 16 plain additions (all lanes would be needed for 16 quads);
// from level 2 on there are at most 8 additions per warp and each gets a quad.
// seg_sum_quad (S >= 4 only) returns the quad-distributed sum: it lives in the first quad of every segment.
__device__ __forceinline__ fp_t seg_sum_quad(xyzz_t acc, int S) {
    const int lane = threadIdx.x & 31, base = lane & ~3;
    {
        xyzz_t o = shfl_down_xyzz(acc, 1);
        if (lane & 1) o = xyzz_t::inf();
        xyzz_add(acc, o);
    }
    fp_t a = quad_scatter(acc, base), b = quad_scatter(acc, base + 2);
    a = quad_add(a, b);
#pragma unroll 1
    for (int step = 4; step < S; step <<= 1) {
        b = shfl_dn_fp(a, step);
        if ((lane & (2 * step - 1)) >= step) b = fp_t::zero();     // quads that are not segment leaders: add infinity
        a = quad_add(a, b);
    }
    return a;
}

// tree over quad-distributed points: quad j += quad j + step/4 for step = 4, 8, ... < lanes; the total ends up in quad 0
// of every aligned block of `lanes` lanes
__device__ __forceinline__ fp_t quad_tree(fp_t a, int lanes) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int step = 4; step < lanes; step <<= 1) {
        fp_t b = shfl_dn_fp(a, step);
        if ((lane & (2 * step - 1)) >= step) b = fp_t::zero();
        a = quad_add(a, b);
    }
    return a;
}
// k = k1 + k2 * z^2 with 0 <= k1 < z^2 (128 bits) and k2 < 2^128, for canonical k < r: plain long division, because
// z^2 = 0xac45a4010001a4020000000100000000 (z = the BLS parameter) is the GLV eigenvalue up to sign -- r = z^4 - z^2 + 1, so
// (-z^2)^2 + (-z^2) + 1 = 0 mod r, and the curve endomorphism (x, y) -> (beta x, y) acts on G1 as multiplication by -z^2
// (the same relation the subgroup test uses, eprint 2021/1130 sec. 6).  Hence [k]P = [k1]P + [k2](beta x, -y): two
// half-length scalars over one shared doubling chain.
__device__ __forceinline__ void glv_split(const uint32_t (&k)[8], uint32_t (&k1)[4], uint32_t (&k2)[4]) {
    const uint32_t Z2[4] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u};
    uint32_t rem[5] = {0, 0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int bit = 255; bit >= 0; bit--) {
        // rem = (rem << 1) | bit of k
#pragma unroll
        for (int i = 4; i > 0; i--) rem[i] = (rem[i] << 1) | (rem[i - 1] >> 31);
        rem[0] = (rem[0] << 1) | ((k[bit >> 5] >> (bit & 31)) & 1u);
        // rem >= Z2 ?
        bool ge = rem[4] != 0;
        if (!ge) {
            ge = true;
#pragma unroll
            for (int i = 3; i >= 0; i--) {
                if (rem[i] != Z2[i]) { ge = rem[i] > Z2[i]; break; }
            }
        }
        if (ge) {
            uint32_t borrow = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint64_t d = (uint64_t)rem[i] - Z2[i] - borrow;
                rem[i] = (uint32_t)d;
                borrow = (uint32_t)(d >> 63);
            }
            rem[4] -= borrow;
            if (bit < 128) q[bit >> 5] |= 1u << (bit & 31);   // the quotient of k < r fits 128 bits
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { k1[i] = rem[i]; k2[i] = q[i]; }
}
// signed 4-bit digits of a 128-bit number, least significant first: d_j in [-7, 8], 33 of them (the last is the carry).
// mag: 33 nibbles |d_j|, sgn: bit j set when d_j < 0.
__device__ __forceinline__ void booth4(const uint32_t (&k)[4], uint32_t (&mag)[5], uint32_t (&sgn)[2]) {
#pragma unroll
    for (int i = 0; i < 5; i++) mag[i] = 0;
    sgn[0] = sgn[1] = 0;
    uint32_t carry = 0;
#pragma unroll 1
    for (int j = 0; j < 32; j++) {
        uint32_t d = ((k[j >> 3] >> ((j & 7) * 4)) & 15u) + carry;
        carry = d > 8u;
        if (carry) {
            d = 16u - d;
            sgn[j >> 5] |= 1u << (j & 31);
        }
        mag[j >> 3] |= d << ((j & 7) * 4);
    }
    mag[4] = carry;
}
// [k] p for quad-distributed p; k canonical little-endian words (< r), the same value on the four lanes of a quad.
// GLV split (above), then signed 4-bit windows over the 128-bit halves, most significant first: a table of the multiples
// d * P, d = 0..8, per quad in shared memory plus the X coordinates of d * (beta x, -y) (Y is negated on the fly, ZZ / ZZZ
// are shared) -- kQuadTableBytes per warp (17 KiB: signed digits halve the table of an unsigned window, which doubles the
// warps an SM can hold -- these kernels are latency chains and live on occupancy -- and shortens the table build from 14
// to 7 point operations); every lane only ever touches its own component slots, so no synchronisation -- then
// 33 x (4 quad doublings + 2 quad additions).  The instruction stream is uniform across the warp's quads -- with a
// bit-serial double-and-add every quad would pay for an addition whenever any of the eight needs one.
static constexpr int kQuadTableEntries = 9;
static constexpr int kQuadTableStride = kQuadTableEntries * 192 + kQuadTableEntries * 48;
static constexpr int kQuadTableBytes = 8 * kQuadTableStride;
__device__ __forceinline__ fp_t quad_mul_scalar(const fp_t& p, const uint32_t (&k)[8], uint8_t* warp_table) {
    const int lane = threadIdx.x & 31, role = lane & 3;
    uint8_t* tab = warp_table + (lane >> 2) * kQuadTableStride + quad_store_offset();
    uint8_t* tabx = warp_table + (lane >> 2) * kQuadTableStride + kQuadTableEntries * 192;   // beta * X_d, 48 B each
    uint32_t k1[4], k2[4];
    glv_split(k, k1, k2);
    uint32_t m1[5], s1[2], m2[5], s2[2];
    booth4(k1, m1, s1);
    booth4(k2, m2, s2);
    fp_t beta;
    {
        const uint32_t Bm[12] = {0x798a64e8u, 0x30f1361bu, 0x7ece5a2au, 0xf3b8ddabu, 0xc61577f7u, 0x16a8ca3au,
                                 0x74fd029bu, 0xc26a2ff8u, 0x60701c6eu, 0x3636b766u, 0x241b6160u, 0x051ba4abu};
#pragma unroll
        for (int i = 0; i < 12; i++) beta.v[i] = Bm[i];
    }
    store_field(tab, fp_t::zero());
    store_field(tab + 192, p);
#pragma unroll 1
    for (int d = 2; d < kQuadTableEntries; d++) {
        fp_t t = (d & 1) ? quad_add(load_field<fp_t>(tab + (d - 1) * 192), p) : quad_dbl(load_field<fp_t>(tab + (d >> 1) * 192));
        store_field(tab + d * 192, t);
    }
#pragma unroll 1
    for (int d = 0; d < kQuadTableEntries; d++) {
        fp_t bx = load_field<fp_t>(tab + d * 192) * beta;      // only the X lane's product is kept
        if (role == 0) store_field(tabx + d * 48, bx);
    }
    fp_t acc = fp_t::zero();
#pragma unroll 1
    for (int j = 32; j >= 0; j--) {
        if (j != 32) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) acc = quad_dbl(acc);
        }
        const uint32_t d1 = (m1[j >> 3] >> ((j & 7) * 4)) & 15u, d2 = (m2[j >> 3] >> ((j & 7) * 4)) & 15u;
        const bool n1 = j < 32 && ((s1[j >> 5] >> (j & 31)) & 1u), n2 = j < 32 && ((s2[j >> 5] >> (j & 31)) & 1u);
        fp_t e = load_field<fp_t>(tab + d1 * 192);             // +-d1 * P
        if (role == 1 && n1) e = e.neg();
        acc = quad_add(acc, e);
        e = load_field<fp_t>(tab + d2 * 192);                  // +-d2 * (beta x, -y) = (beta X_d2, -+Y_d2, ZZ_d2, ZZZ_d2)
        if (role == 0) e = load_field<fp_t>(tabx + d2 * 48);
        if (role == 1 && !n2) e = e.neg();
        acc = quad_add(acc, e);
    }
    return acc;
}

// One GLV half of [k] p: half = 0 gives [k1] p, half = 1 gives [k2] (beta x, -y), so that TWO quads (neighbours in a warp)
// share one scalar multiplication -- each runs 33 x (4 doublings + 1 addition) instead of 33 x (4 + 2), a 20 % shorter
// dependent chain, and the caller adds the two halves (shfl_xor 4).  The instruction stream is the same for both halves:
// only the digit, the sign rule and the source of the X component differ, all per-lane selects.
__device__ __forceinline__ fp_t quad_mul_scalar_half(const fp_t& p, const uint32_t (&k)[8], uint8_t* warp_table, int half) {
    const int lane = threadIdx.x & 31, role = lane & 3;
    uint8_t* tab = warp_table + (lane >> 2) * kQuadTableStride + quad_store_offset();
    uint8_t* tabx = warp_table + (lane >> 2) * kQuadTableStride + kQuadTableEntries * 192;   // beta * X_d, 48 B each
    uint32_t k1[4], k2[4];
    glv_split(k, k1, k2);
    uint32_t m[5], sg[2];
    if (half) booth4(k2, m, sg); else booth4(k1, m, sg);
    fp_t beta;
    {
        const uint32_t Bm[12] = {0x798a64e8u, 0x30f1361bu, 0x7ece5a2au, 0xf3b8ddabu, 0xc61577f7u, 0x16a8ca3au,
                                 0x74fd029bu, 0xc26a2ff8u, 0x60701c6eu, 0x3636b766u, 0x241b6160u, 0x051ba4abu};
#pragma unroll
        for (int i = 0; i < 12; i++) beta.v[i] = Bm[i];
    }
    store_field(tab, fp_t::zero());
    store_field(tab + 192, p);
#pragma unroll 1
    for (int d = 2; d < kQuadTableEntries; d++) {
        fp_t t = (d & 1) ? quad_add(load_field<fp_t>(tab + (d - 1) * 192), p) : quad_dbl(load_field<fp_t>(tab + (d >> 1) * 192));
        store_field(tab + d * 192, t);
    }
#pragma unroll 1
    for (int d = 0; d < kQuadTableEntries; d++) {
        fp_t bx = load_field<fp_t>(tab + d * 192) * beta;      // only the X lane's product is kept
        if (role == 0) store_field(tabx + d * 48, bx);
    }
    const bool use_bx = half && role == 0;
    fp_t acc = fp_t::zero();
#pragma unroll 1
    for (int j = 32; j >= 0; j--) {
        if (j != 32) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) acc = quad_dbl(acc);
        }
        const uint32_t d = (m[j >> 3] >> ((j & 7) * 4)) & 15u;
        const bool neg = j < 32 && ((sg[j >> 5] >> (j & 31)) & 1u);
        fp_t e = use_bx ? load_field<fp_t>(tabx + d * 48) : load_field<fp_t>(tab + d * 192);
        if (role == 1 && (neg != (half != 0))) e = e.neg();      // half 1 carries the endomorphism's -y
        acc = quad_add(acc, e);
    }
    return acc;
}

// [|z|] p for quad-distributed p, z = -0xd201000000010000 the BLS parameter: the ladder of the subgroup test
// (beta x, y) == -[z^2] P  (eprint 2021/1130 sec. 6, zkcrypto/bls12_381/src/g1.rs:401-410); the scalar is a constant, so
// every quad of a warp follows the same instruction stream
__device__ __forceinline__ fp_t quad_mul_by_abs_z(const fp_t& base) {
    const uint64_t Z = 0xd201000000010000ull;
    fp_t acc = base;
#pragma unroll 1
    for (int bit = 62; bit >= 0; bit--) {
        acc = quad_dbl(acc);
        if ((Z >> bit) & 1) acc = quad_add(acc, base);
    }
    return acc;
}
// p (affine, non-infinity, on the curve) is in the prime-order subgroup; comp = p quad-distributed (X, Y, 1, 1).
// The verdict is valid on the first lane of the quad.  All 32 lanes must call.
__device__ __forceinline__ bool quad_in_subgroup(const affine_t& a, const fp_t& comp) {
    const int base = threadIdx.x & 28;
    fp_t t = quad_mul_by_abs_z(comp);                      // |z| P
    const bool t_inf = __shfl_sync(kFullMask, (int)t.is_zero(), base | 2);
    fp_t u = quad_mul_by_abs_z(t);                         // z^2 P
    xyzz_t U = quad_gather(u);                             // valid on the quad's first lane
    fp_t beta;
    const uint32_t Bm[12] = {0x798a64e8u, 0x30f1361bu, 0x7ece5a2au, 0xf3b8ddabu, 0xc61577f7u, 0x16a8ca3au,
                             0x74fd029bu, 0xc26a2ff8u, 0x60701c6eu, 0x3636b766u, 0x241b6160u, 0x051ba4abu};
#pragma unroll
    for (int k = 0; k < 12; k++) beta.v[k] = Bm[k];
    // (beta x, y) == -(X/ZZ, Y/ZZZ)  <=>  beta x ZZ == X  and  y ZZZ == -Y
    return !t_inf && !U.is_inf() && (beta * a.x * U.zz == U.x) && (a.y * U.zzz == U.y.neg());
}

// lane 0 of the warp ends up with the sum over the warp's 32 points
__device__ __forceinline__ xyzz_t warp_sum_xyzz(const xyzz_t& v) { return quad_gather(seg_sum_quad(v, 32)); }

}  // namespace b200
