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

// 2 * a   (dbl-2008-s-1 in three multiplication levels)
static __device__ __noinline__ fp_t quad_dbl(fp_t a) {
    const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
    const bool a_inf = __shfl_sync(kFullMask, (int)a.is_zero(), base | 2);
    fp_t U = shfl_idx_fp(a.dbl(), base | 1);                       // U = 2Y on every lane
    // level 1: lane0 X^2 ; lanes 1..3 V = U^2
    fp_t m = role == 0 ? a : U;
    fp_t l1 = m.sqr();
    fp_t M = l1.dbl() + l1;                                        // lane0: M = 3 X^2
    fp_t V = shfl_idx_fp(l1, base | 1);
    fp_t Mb = shfl_idx_fp(M, base);
    // level 2: lane0 S = X V ; lane1 W = U V ; lane2 ZZ3 = ZZ V ; lane3 M^2
    fp_t m1 = role == 3 ? Mb : (role == 1 ? U : a);
    fp_t m2 = role == 3 ? Mb : V;
    fp_t l2 = m1 * m2;
    fp_t MM = shfl_idx_fp(l2, base | 3);
    fp_t Wb = shfl_idx_fp(l2, base | 1);
    fp_t x3 = MM - l2.dbl();                                       // lane0: M^2 - 2S
    // level 3: lane0 M (S - X3) ; lane1 W Y ; lane3 ZZZ3 = W ZZZ ; lane2 idle
    m1 = role == 0 ? M : Wb;
    m2 = role == 0 ? (l2 - x3) : a;
    fp_t l3 = m1 * m2;
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
    fp_t l1 = a * shfl_xor_fp(b, 2);
    fp_t l1x = shfl_xor_fp(l1, 2);
    fp_t pr = role < 2 ? (l1x - l1) : (l1 - l1x);                  // lanes 0,2: P = U2 - U1 ; lanes 1,3: R = S2 - S1
    const bool p_zero = __shfl_sync(kFullMask, (int)pr.is_zero(), base);
    const bool r_zero = __shfl_sync(kFullMask, (int)pr.is_zero(), base | 1);
    // level 2: lane0 PP = P^2 ; lane1 RR = R^2 ; lane2 ZZ1 ZZ2 ; lane3 ZZZ1 ZZZ2
    fp_t m1 = role < 2 ? pr : a;
    fp_t m2 = role < 2 ? pr : b;
    fp_t l2 = m1 * m2;
    fp_t pp = shfl_idx_fp(l2, base);
    fp_t u1 = shfl_idx_fp(l1, base);
    // level 3: lane0 PPP = P PP ; lane1 Q = U1 PP ; lane2 ZZ3 = (ZZ1 ZZ2) PP ; lane3 idle
    m1 = role == 0 ? pr : role == 1 ? u1 : l2;
    fp_t l3 = m1 * pp;
    fp_t ppp = shfl_idx_fp(l3, base);
    fp_t s1 = shfl_idx_fp(l1, base | 1);
    fp_t x3 = l2 - ppp - l3.dbl();                                 // lane1: RR - PPP - 2Q
    // level 4: lane0 S1 PPP ; lane1 R (Q - X3) ; lane3 ZZZ3 = (ZZZ1 ZZZ2) PPP ; lane2 idle
    m1 = role == 0 ? s1 : role == 1 ? pr : l2;
    m2 = role == 1 ? (l3 - x3) : ppp;
    fp_t l4 = m1 * m2;
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
// [k] p for quad-distributed p; k canonical little-endian words, the same value on the four lanes of a quad.
// Fixed 4-bit windows, most significant first: 16-entry table of multiples per quad in shared memory
// (kQuadTableBytes per warp; every lane only ever touches its own component slots, so no synchronisation), then 64 x
// (4 quad doublings + 1 quad addition).  The instruction stream is uniform across the warp's quads -- with a
// bit-serial double-and-add every quad would pay for an addition whenever any of the eight needs one.
static constexpr int kQuadTableBytes = 8 * 16 * 192;
__device__ __forceinline__ fp_t quad_mul_scalar(const fp_t& p, const uint32_t (&k)[8], uint8_t* warp_table) {
    const int lane = threadIdx.x & 31;
    uint8_t* tab = warp_table + (lane >> 2) * (16 * 192) + quad_store_offset();
    store_field(tab, fp_t::zero());
    store_field(tab + 192, p);
#pragma unroll 1
    for (int d = 2; d < 16; d++) {
        fp_t t = (d & 1) ? quad_add(load_field<fp_t>(tab + (d - 1) * 192), p) : quad_dbl(load_field<fp_t>(tab + (d >> 1) * 192));
        store_field(tab + d * 192, t);
    }
    fp_t acc = load_field<fp_t>(tab + ((k[7] >> 28) & 15) * 192);
#pragma unroll 1
    for (int j = 62; j >= 0; j--) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) acc = quad_dbl(acc);
        uint32_t d = (k[j >> 3] >> ((j & 7) * 4)) & 15;
        acc = quad_add(acc, load_field<fp_t>(tab + d * 192));
    }
    return acc;
}

// lane 0 of the warp ends up with the sum over the warp's 32 points
__device__ __forceinline__ xyzz_t warp_sum_xyzz(const xyzz_t& v) { return quad_gather(seg_sum_quad(v, 32)); }

}  // namespace b200
