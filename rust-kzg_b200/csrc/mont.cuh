// mont.cuh -- Montgomery arithmetic for BLS12-381 Fp (381 bit, 12 x u32) and Fr (255 bit, 8 x u32) on sm_100a.
//
// Layout: little-endian 32-bit limbs in registers, Montgomery form with R = 2^384 (Fp) / 2^256 (Fr) -- bit-identical
// to blst's blst_fp / blst_fr (kzg/src/eth/c_bindings.rs:427-474), so values cross the C ABI without conversion.
// Values are always fully reduced to [0, mod) between operations (bit-exact byte output, cheap zero tests).
//
// Multiplication is word-serial Montgomery (CIOS) on the integer pipe.  The running sum T is held as two
// staggered arrays of 64-bit columns,
//        T = E + (O << 32),     E column k = (E[2k], E[2k+1]),   O column k = (O[2k], O[2k+1]),
// so that every 32x32->64 product lands on a 64-bit aligned column and a whole row a[even]*w (resp. a[odd]*w) is one
// carry chain of mad.lo.cc/madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32(.X) -- one SASS instruction per
// 32x32 product.  Dividing by 2^32 after each row swaps the roles of E and O (a register renaming in the fully
// unrolled code); the one stray word (old E[1]) is added to new E[0] and its carry is absorbed as the carry-in of
// the next O-chain, which sits exactly one limb higher.  Per row: 2N wide multiply-adds + 4 scalar ops.
#pragma once
#include "inv_bingcd.cuh"
#include <stdint.h>

namespace b200 {

// ---- carry-chain rows: acc(2*PAIRS limbs) += a[0], a[2], ... (stride 2) * w ------------------------------------
// CIN: take the incoming carry flag; the outgoing carry flag is left live for the caller's next asm statement
// (the statements are adjacent and compiler-generated PTX never touches CC).
template <int PAIRS, bool CIN>
struct Chain;

template <bool CIN>
struct Chain<6, CIN> {
    static __device__ __forceinline__ void mad(uint32_t* acc, const uint32_t* a, uint32_t w) {
        if (CIN)
            asm volatile(
                "madc.lo.cc.u32 %0, %12, %18, %0;  madc.hi.cc.u32 %1, %12, %18, %1;\n\t"
                "madc.lo.cc.u32 %2, %13, %18, %2;  madc.hi.cc.u32 %3, %13, %18, %3;\n\t"
                "madc.lo.cc.u32 %4, %14, %18, %4;  madc.hi.cc.u32 %5, %14, %18, %5;\n\t"
                "madc.lo.cc.u32 %6, %15, %18, %6;  madc.hi.cc.u32 %7, %15, %18, %7;\n\t"
                "madc.lo.cc.u32 %8, %16, %18, %8;  madc.hi.cc.u32 %9, %16, %18, %9;\n\t"
                "madc.lo.cc.u32 %10, %17, %18, %10; madc.hi.cc.u32 %11, %17, %18, %11;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
                  "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(w));
        else
            asm volatile(
                "mad.lo.cc.u32 %0, %12, %18, %0;   madc.hi.cc.u32 %1, %12, %18, %1;\n\t"
                "madc.lo.cc.u32 %2, %13, %18, %2;  madc.hi.cc.u32 %3, %13, %18, %3;\n\t"
                "madc.lo.cc.u32 %4, %14, %18, %4;  madc.hi.cc.u32 %5, %14, %18, %5;\n\t"
                "madc.lo.cc.u32 %6, %15, %18, %6;  madc.hi.cc.u32 %7, %15, %18, %7;\n\t"
                "madc.lo.cc.u32 %8, %16, %18, %8;  madc.hi.cc.u32 %9, %16, %18, %9;\n\t"
                "madc.lo.cc.u32 %10, %17, %18, %10; madc.hi.cc.u32 %11, %17, %18, %11;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
                  "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(w));
    }
    // acc = a * w (no accumulate), no carries can occur
    static __device__ __forceinline__ void mul(uint32_t* acc, const uint32_t* a, uint32_t w) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            uint64_t t;                                   // one IMAD.WIDE.U32 instead of IMAD + IMAD.HI
            asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[2 * k]), "r"(w));
            acc[2 * k] = (uint32_t)t;
            acc[2 * k + 1] = (uint32_t)(t >> 32);
        }
    }
};

template <bool CIN>
struct Chain<4, CIN> {
    static __device__ __forceinline__ void mad(uint32_t* acc, const uint32_t* a, uint32_t w) {
        if (CIN)
            asm volatile(
                "madc.lo.cc.u32 %0, %8, %12, %0;  madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                "madc.lo.cc.u32 %2, %9, %12, %2;  madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                "madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.cc.u32 %7, %11, %12, %7;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
                  "+r"(acc[7])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
        else
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %12, %0;   madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                "madc.lo.cc.u32 %2, %9, %12, %2;  madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                "madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.cc.u32 %7, %11, %12, %7;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
                  "+r"(acc[7])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
    }
    static __device__ __forceinline__ void mul(uint32_t* acc, const uint32_t* a, uint32_t w) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t t;                                   // one IMAD.WIDE.U32 instead of IMAD + IMAD.HI
            asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[2 * k]), "r"(w));
            acc[2 * k] = (uint32_t)t;
            acc[2 * k + 1] = (uint32_t)(t >> 32);
        }
    }
};

template <bool CIN>
struct Chain<3, CIN> {
    static __device__ __forceinline__ void mad(uint32_t* acc, const uint32_t* a, uint32_t w) {
        if (CIN)
            asm volatile(
                "madc.lo.cc.u32 %0, %6, %9, %0;  madc.hi.cc.u32 %1, %6, %9, %1;\n\t"
                "madc.lo.cc.u32 %2, %7, %9, %2;  madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
                "madc.lo.cc.u32 %4, %8, %9, %4;  madc.hi.cc.u32 %5, %8, %9, %5;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(w));
        else
            asm volatile(
                "mad.lo.cc.u32 %0, %6, %9, %0;   madc.hi.cc.u32 %1, %6, %9, %1;\n\t"
                "madc.lo.cc.u32 %2, %7, %9, %2;  madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
                "madc.lo.cc.u32 %4, %8, %9, %4;  madc.hi.cc.u32 %5, %8, %9, %5;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(w));
    }
    static __device__ __forceinline__ void mul(uint32_t* acc, const uint32_t* a, uint32_t w) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint64_t t;                                   // one IMAD.WIDE.U32 instead of IMAD + IMAD.HI
            asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[2 * k]), "r"(w));
            acc[2 * k] = (uint32_t)t;
            acc[2 * k + 1] = (uint32_t)(t >> 32);
        }
    }
};
template <bool CIN>
struct Chain<2, CIN> {
    static __device__ __forceinline__ void mad(uint32_t* acc, const uint32_t* a, uint32_t w) {
        if (CIN)
            asm volatile(
                "madc.lo.cc.u32 %0, %4, %6, %0;  madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
                "madc.lo.cc.u32 %2, %5, %6, %2;  madc.hi.cc.u32 %3, %5, %6, %3;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
                : "r"(a[0]), "r"(a[2]), "r"(w));
        else
            asm volatile(
                "mad.lo.cc.u32 %0, %4, %6, %0;   madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
                "madc.lo.cc.u32 %2, %5, %6, %2;  madc.hi.cc.u32 %3, %5, %6, %3;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
                : "r"(a[0]), "r"(a[2]), "r"(w));
    }
    static __device__ __forceinline__ void mul(uint32_t* acc, const uint32_t* a, uint32_t w) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            uint64_t t;                                   // one IMAD.WIDE.U32 instead of IMAD + IMAD.HI
            asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[2 * k]), "r"(w));
            acc[2 * k] = (uint32_t)t;
            acc[2 * k + 1] = (uint32_t)(t >> 32);
        }
    }
};

// Plain product of two H-limb numbers (H even), 2H limbs out, with the same staggered E / O carry-chain rows as the
// Montgomery multiplier below but no reduction: H * H wide multiply-adds.  Row i leaves limb i of the product in E[0].
template <int H>
__device__ __forceinline__ void wide_mul(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    constexpr int HP = H / 2;
    uint32_t E[H + 2], O[H + 2];
    Chain<HP, false>::mul(E, a, b[0]);
    Chain<HP, false>::mul(O, a + 1, b[0]);
    E[H] = 0;
    out[0] = E[0];
#pragma unroll
    for (int i = 1; i < H; i++) {
        // T >>= 32:  E' = O (+ stray E[1]),  O'[k] = E[k+2]
        uint32_t stray = E[1];
        uint32_t nE[H + 2], nO[H + 2];
#pragma unroll
        for (int k = 0; k < H; k++) nE[k] = O[k];
#pragma unroll
        for (int k = 0; k < H - 1; k++) nO[k] = E[k + 2];
        nO[H - 1] = 0;
#pragma unroll
        for (int k = 0; k < H; k++) { E[k] = nE[k]; O[k] = nO[k]; }
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(stray));
        Chain<HP, true>::mad(O, a + 1, b[i]);
        Chain<HP, false>::mad(E, a, b[i]);
        asm volatile("addc.u32 %0, 0, 0;" : "=r"(E[H]));
        out[i] = E[0];
    }
    // the high half: (E >> 32) + O
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(out[H]) : "r"(O[0]), "r"(E[1]));
#pragma unroll
    for (int k = 1; k < H - 1; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(out[H + k]) : "r"(O[k]), "r"(E[k + 1]));
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(out[2 * H - 1]) : "r"(O[H - 1]), "r"(E[H]));
}

// ---- field parameters -------------------------------------------------------------------------------------------
// moduli: zkcrypto/bls12_381/src/fp.rs:71, scalar.rs:77 (cross-checked with arkworks3-sppark-wlc/sppark/ff/bls12-381.hpp:10-39)
struct FpParams {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^-1 mod 2^32
    static __device__ __forceinline__ constexpr uint32_t mod(int i) {
        constexpr uint32_t M[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return M[i];
    }
    static __device__ __forceinline__ constexpr uint32_t one(int i) {  // R mod p
        constexpr uint32_t M[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return M[i];
    }
    static __device__ __forceinline__ constexpr uint32_t rr(int i) {  // R^2 mod p
        constexpr uint32_t M[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return M[i];
    }
};
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;  // -r^-1 mod 2^32
    static __device__ __forceinline__ constexpr uint32_t mod(int i) {
        constexpr uint32_t M[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return M[i];
    }
    static __device__ __forceinline__ constexpr uint32_t one(int i) {  // R mod r
        constexpr uint32_t M[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return M[i];
    }
    static __device__ __forceinline__ constexpr uint32_t rr(int i) {  // R^2 mod r
        constexpr uint32_t M[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return M[i];
    }
};

// ---- the field element ------------------------------------------------------------------------------------------
// COMPACT = false: fully unrolled multiplier (register renaming, ~370 SASS instructions) for the throughput kernels.
// COMPACT = true: the same row recurrence as a rolled loop (~1 KiB of code).  Cold, low-occupancy phases (bucket
// reduction, table build, compression) are instruction-fetch bound with the unrolled form: a point addition is
// ~80 KiB of straight-line code, far beyond the 32 KiB instruction cache, and runs at tens of cycles per
// instruction when every line misses.
//
// MODE 2 ("r28"): the same Montgomery product (same R, same packed operands and result) computed in radix 2^28:
// with 28-bit limbs a 64-bit column absorbs all 2L partial products (< 2^56 each) of a CIOS sweep without overflow,
// so every product is a carry-free IMAD.WIDE and the carries become shifts / adds on the ALU pipe.  R = 2^(32N) is
// kept by making the last reduction step 32N - 28(L-1) bits wide (20 for Fp, 4 for Fr); the result is re-packed to
// 32-bit limbs from that bit offset.  An experiment that did NOT pay off on B200 (IMAD.WIDE costs the same with or
// without carries, and this form needs 406 of them instead of 302); kept as an independently derived cross-check.
enum { MONT_UNROLLED = 0, MONT_COMPACT = 1, MONT_R28 = 2, MONT_CALL = 3, MONT_DFMA = 4, MONT_KARA = 5 };
template <class P, int MODE = MONT_UNROLLED>
struct __align__(16) Mont {
    static constexpr bool COMPACT = MODE == MONT_COMPACT;
    typedef P params_t;
    static constexpr int N = P::N;
    uint32_t v[N];

    static __device__ __forceinline__ Mont zero() {
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    static __device__ __forceinline__ Mont one() {
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::one(i);
        return r;
    }
    static __device__ __forceinline__ Mont rr() {
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::rr(i);
        return r;
    }
    __device__ __forceinline__ bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i];
        return acc == 0;
    }
    __device__ __forceinline__ bool operator==(const Mont& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i] ^ o.v[i];
        return acc == 0;
    }
    __device__ __forceinline__ bool operator!=(const Mont& o) const { return !(*this == o); }

    // r = r - mod if r >= mod   (r < 2*mod on entry, top carry bit passed in `hi`)
    __device__ __forceinline__ void final_sub(uint32_t hi) {
        uint32_t t[N];
        uint32_t borrow;
        asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(v[0]), "r"(P::mod(0)));
#pragma unroll
        for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[i]) : "r"(v[i]), "r"(P::mod(i)));
        asm volatile("subc.u32 %0, %1, 0;" : "=r"(borrow) : "r"(hi));
        // borrow == 0  <=>  value >= mod  -> keep t
        if (borrow == 0) {
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = t[i];
        }
    }

    friend __device__ __forceinline__ Mont operator+(const Mont& a, const Mont& b) {
        Mont r;
        uint32_t hi;
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
        for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
        asm volatile("addc.u32 %0, 0, 0;" : "=r"(hi));
        r.final_sub(hi);
        return r;
    }
    friend __device__ __forceinline__ Mont operator-(const Mont& a, const Mont& b) {
        Mont r;
        uint32_t borrow;
        asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
        for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
        asm volatile("subc.u32 %0, 0, 0;" : "=r"(borrow));
        // borrow is 0 or 0xffffffff: add (mod & borrow)
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r.v[0]) : "r"(P::mod(0) & borrow));
#pragma unroll
        for (int i = 1; i < N - 1; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r.v[i]) : "r"(P::mod(i) & borrow));
        asm volatile("addc.u32 %0, %0, %1;" : "+r"(r.v[N - 1]) : "r"(P::mod(N - 1) & borrow));
        return r;
    }
    __device__ __forceinline__ Mont neg() const {  // -a mod m  (0 -> 0)
        Mont r;
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < N; i++) nz |= v[i];
        asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(P::mod(0)), "r"(v[0]));
#pragma unroll
        for (int i = 1; i < N - 1; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(P::mod(i)), "r"(v[i]));
        asm volatile("subc.u32 %0, %1, %2;" : "=r"(r.v[N - 1]) : "r"(P::mod(N - 1)), "r"(v[N - 1]));
        if (nz == 0) r = *this;
        return r;
    }
    __device__ __forceinline__ Mont cneg(bool flag) const {
        Mont n = neg();
        return flag ? n : *this;
    }
    __device__ __forceinline__ Mont dbl() const { return *this + *this; }

    // Montgomery product a*b*R^-1 mod m, fully reduced.
    friend __device__ __forceinline__ Mont operator*(const Mont& a, const Mont& b) {
        if (MODE == MONT_COMPACT) return mul_compact(a, b);
        if (MODE == MONT_R28) return mul_r28(a, b, false);
        if (MODE == MONT_CALL) return mul_call(a, b);
        if (MODE == MONT_DFMA) return mul_dfma(a, b);
        if (MODE == MONT_KARA) return mul_kara(a, b);
        return mul_unrolled(a, b);
    }
    // ---- MODE 4: the a*b half of the product on the FP64 pipe ---------------------------------------------------
    // B200 issues DFMA at 64/clk/SM on a pipe of its own, while every IMAD.WIDE costs 4 cycles of the FMA-heavy pipe
    // and the carry-chain multiplier is bound by exactly that pipe.  The 2N x 32-bit product a*b is therefore formed
    // from 48-bit limbs held as doubles: for limbs A, B < 2^48 (A*B < 2^96)
    //      s1 = fma_rz(A, B, 2^100)          = 2^100 + H * 2^48,   H = floor(A*B / 2^48)   (ulp(2^100) = 2^48)
    //      s2 = fma_rz(A, B, 2^100 + 2^52 - s1) = 2^52 + L,         L = A*B - H * 2^48      (exact)
    // so the IEEE mantissa fields of s1, s2 ARE H and L.  The raw 64-bit patterns are summed per 48-bit column with
    // integer adds (ALU pipe); at most 16 terms < 2^48 land in a column, so the low 52 bits of the sum are the exact
    // column value and the exponent bits above never interfere.  The columns are carried into 32-bit words and the
    // 2N-word product is Montgomery-reduced word by word with the carry-chain rows (the m*p half stays on IMAD.WIDE).
    static constexpr int LD = (32 * N + 47) / 48;  // 48-bit limbs
    static __device__ __forceinline__ void limbs48(double* d, const Mont& a) {
#pragma unroll
        for (int i = 0; i < LD; i++) {
            const int bit = 48 * i, w = bit >> 5, sh = bit & 31;  // sh is 0 or 16
            uint32_t w0 = w < N ? a.v[w] : 0u, w1 = w + 1 < N ? a.v[w + 1] : 0u, w2 = w + 2 < N ? a.v[w + 2] : 0u;
            uint32_t lo, hi16;
            if (sh == 0) { lo = w0; hi16 = w1 & 0xffffu; }
            else { lo = (w0 >> 16) | (w1 << 16); hi16 = (w1 >> 16) | ((w2 & 0u) << 16); hi16 &= 0xffffu; }
            // 2^52 + limb as a bit pattern, minus 2^52: exact conversion without the (slow) I2F unit
            d[i] = __dsub_rn(__hiloint2double((int)(0x43300000u | hi16), (int)lo), 4503599627370496.0);
        }
    }
    static __device__ __forceinline__ Mont mul_dfma(const Mont& a, const Mont& b) {
        double ad[LD], bd[LD];
        limbs48(ad, a);
        limbs48(bd, b);
        const double C1 = 1267650600228229401496703205376.0;                       // 2^100
        const double C1P = 1267650600228229401496703205376.0 + 4503599627370496.0;  // 2^100 + 2^52 (exact: 49 bits)
        uint64_t col[2 * LD];
#pragma unroll
        for (int k = 0; k < 2 * LD; k++) col[k] = 0;
#pragma unroll
        for (int i = 0; i < LD; i++) {
#pragma unroll
            for (int j = 0; j < LD; j++) {
                double s1 = __fma_rz(ad[i], bd[j], C1);
                double s2 = __fma_rz(ad[i], bd[j], __dsub_rn(C1P, s1));
                col[i + j + 1] += (uint64_t)__double_as_longlong(s1);
                col[i + j] += (uint64_t)__double_as_longlong(s2);
            }
        }
        // columns (weight 2^(48k), value < 2^52 in the low 52 bits) -> 2N words of 32 bits
        uint32_t T[2 * N + 1];
        {
            const uint64_t M52 = (1ull << 52) - 1, M48 = (1ull << 48) - 1;
            uint64_t carry = 0;
            uint64_t limb[2 * LD];
#pragma unroll
            for (int k = 0; k < 2 * LD; k++) {
                uint64_t v = (col[k] & M52) + carry;
                limb[k] = v & M48;
                carry = v >> 48;
            }
#pragma unroll
            for (int w = 0; w < 2 * N; w++) {
                const int bit = 32 * w, k = bit / 48, off = bit % 48;  // off in {0, 32, 16}
                uint64_t x = limb[k] >> off;
                if (off > 16 && k + 1 < 2 * LD) x |= limb[k + 1] << (48 - off);
                T[w] = (uint32_t)x;
            }
            T[2 * N] = 0;
        }
        // word-serial Montgomery reduction of the 2N-word product: T += m_i * mod * 2^(32 i)
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t m = T[i] * P::INV;
            // even-indexed modulus words: 64-bit products aligned at T[i + 2k]
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(T[i]), "+r"(T[i + 1]) : "r"(m), "r"(P::mod(0)));
#pragma unroll
            for (int k = 2; k < N; k += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(T[i + k]), "+r"(T[i + k + 1]) : "r"(m), "r"(P::mod(k)));
#pragma unroll
            for (int k = i + N; k < 2 * N; k++) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(T[k]));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(T[2 * N]));
            // odd-indexed modulus words: aligned at T[i + 1 + 2k]
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(T[i + 1]), "+r"(T[i + 2]) : "r"(m), "r"(P::mod(1)));
#pragma unroll
            for (int k = 3; k < N; k += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(T[i + k]), "+r"(T[i + k + 1]) : "r"(m), "r"(P::mod(k)));
#pragma unroll
            for (int k = i + N + 1; k < 2 * N; k++) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(T[k]));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(T[2 * N]));
        }
        Mont r;
#pragma unroll
        for (int k = 0; k < N; k++) r.v[k] = T[N + k];
        r.final_sub(T[2 * N]);
        return r;
    }
    // MODE 3: the unrolled multiplier behind a real call, so a point addition is ~10 calls instead of 60 KiB of inlined
    // code (the accumulate loop body otherwise exceeds the instruction cache: ncu shows "no instruction" stalls)
    static __device__ __noinline__ Mont mul_call(Mont a, Mont b) { return mul_unrolled(a, b); }
    // ---- radix-2^28 product ------------------------------------------------------------------------------------
    static constexpr int L28 = (32 * N + 27) / 28;            // limbs
    static constexpr int LAST_BITS = 32 * N - 28 * (L28 - 1);  // width of the final reduction step
    static constexpr uint32_t M28 = (1u << 28) - 1;
    static __device__ __forceinline__ constexpr uint32_t mod28(int k) {  // limb k of the modulus in radix 2^28
        int bit = 28 * k, w = bit >> 5, sh = bit & 31;
        uint64_t two = (uint64_t)(w < N ? P::mod(w) : 0u) | ((uint64_t)(w + 1 < N ? P::mod(w + 1) : 0u) << 32);
        return (uint32_t)(two >> sh) & M28;
    }
    static __device__ __forceinline__ void unpack28(uint32_t* l, const Mont& a) {
#pragma unroll
        for (int k = 0; k < L28; k++) {
            const int bit = 28 * k, w = bit >> 5, sh = bit & 31;
            uint32_t lo = a.v[w], hi = w + 1 < N ? a.v[w + 1] : 0u;
            l[k] = (sh ? __funnelshift_r(lo, hi, sh) : lo) & M28;
        }
    }
    // c + a * b as one carry-free IMAD.WIDE.U32 (explicit PTX: the C++ form makes nvcc treat constant operands as
    // 64-bit and emit a redundant high-word add per product)
    static __device__ __forceinline__ uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
        return c + (uint64_t)a * b;
    }
    static __device__ __forceinline__ Mont mul_r28(const Mont& a, const Mont& b, bool square) {
        uint32_t al[L28], bl[L28];
        unpack28(al, a);
        if (square) {
#pragma unroll
            for (int k = 0; k < L28; k++) bl[k] = al[k] << 1;  // doubled operand for the cross terms, < 2^29
        } else {
            unpack28(bl, b);
        }
        uint64_t c[L28 + 1];
#pragma unroll
        for (int k = 0; k <= L28; k++) c[k] = 0;
#pragma unroll
        for (int i = 0; i < L28; i++) {
            // row i of the product (columns are relative to the current shift).  Squaring: only the pairs (i, j >= i),
            // cross terms doubled; every pair (i', j') reaches absolute column i' + j' in row min(i', j') <= column.
            if (square) {
                c[i] = madw(al[i], al[i], c[i]);
#pragma unroll
                for (int j = i + 1; j < L28; j++) c[j] = madw(bl[j], al[i], c[j]);
            } else {
#pragma unroll
                for (int j = 0; j < L28; j++) c[j] = madw(al[j], bl[i], c[j]);
            }
            // reduction step: make the lowest column divisible by 2^bits, then drop it
            const int bits = i == L28 - 1 ? LAST_BITS : 28;
            const uint32_t mask = (1u << bits) - 1;
            // (inline PTX keeps m a 32-bit value: otherwise nvcc widens the masked product to 64 bits and the row
            // below becomes 64x64 multiplies)
            uint32_t m, c0 = (uint32_t)c[0];
            asm("mul.lo.u32 %0, %1, %2; and.b32 %0, %0, %3;" : "=r"(m) : "r"(c0), "r"(P::INV), "r"(mask));
#pragma unroll
            for (int j = 0; j < L28; j++) c[j] = madw(m, mod28(j), c[j]);
            if (i < L28 - 1) {
                c[1] += c[0] >> 28;
#pragma unroll
                for (int k = 0; k < L28; k++) c[k] = c[k + 1];
                c[L28] = 0;
            }
        }
        // value = (sum_k c[k] 2^(28k)) >> LAST_BITS : normalise the columns to 28-bit limbs, then re-pack to 32-bit words
        uint32_t l[L28 + 2];
#pragma unroll
        for (int k = 0; k < L28; k++) {
            l[k] = (uint32_t)c[k] & M28;
            c[k + 1] += c[k] >> 28;
        }
        l[L28] = (uint32_t)c[L28] & M28;
        l[L28 + 1] = (uint32_t)(c[L28] >> 28);
        Mont r;
#pragma unroll
        for (int w = 0; w < N; w++) {
            const int bit = LAST_BITS + 32 * w, k = bit / 28, off = bit % 28;
            // 32 bits starting at limb k, offset off: off <= 24 for both fields, so two limbs always suffice... unless
            // 28 - off < 4; take a third limb when needed
            uint32_t x = l[k] >> off;
            x |= l[k + 1] << (28 - off);
            if (56 - off < 32) x |= l[k + 2] << (56 - off);
            r.v[w] = x;
        }
        // T < 2 * mod: the bit above the packed words is zero for both fields (2p < 2^384, 2r < 2^256)
        r.final_sub(0);
        return r;
    }
    static __device__ __forceinline__ Mont mul_compact(const Mont& a, const Mont& b) {
        constexpr int H = N / 2;
        uint32_t E[N + 2], O[N + 2], bw[N];
        uint32_t mod_[N];
#pragma unroll
        for (int i = 0; i < N; i++) { mod_[i] = P::mod(i); bw[i] = b.v[i]; }
        Chain<H, false>::mul(E, a.v, bw[0]);
        Chain<H, false>::mul(O, a.v + 1, bw[0]);
        E[N] = 0;
#pragma unroll 1
        for (int i = 1; i < N; i++) {
            uint32_t m = E[0] * P::INV;
            Chain<H, false>::mad(O, mod_ + 1, m);
            Chain<H, false>::mad(E, mod_, m);
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N]));
            uint32_t stray = E[1];
            // T >>= 32 by moving registers; rotate the multiplier words so that bw[0] is always the next one
#pragma unroll
            for (int k = 0; k < N; k++) { uint32_t t = O[k]; O[k] = k < N - 1 ? E[k + 2] : 0; E[k] = t; }
#pragma unroll
            for (int k = 0; k < N - 1; k++) bw[k] = bw[k + 1];
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(stray));
            Chain<H, true>::mad(O, a.v + 1, bw[0]);
            Chain<H, false>::mad(E, a.v, bw[0]);
            asm volatile("addc.u32 %0, 0, 0;" : "=r"(E[N]));
        }
        {
            uint32_t m = E[0] * P::INV;
            Chain<H, false>::mad(O, mod_ + 1, m);
            Chain<H, false>::mad(E, mod_, m);
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N]));
        }
        Mont r;
        uint32_t hi;
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(O[0]), "r"(E[1]));
#pragma unroll
        for (int k = 1; k < N; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.v[k]) : "r"(O[k]), "r"(E[k + 1]));
        asm volatile("addc.u32 %0, 0, 0;" : "=r"(hi));
        r.final_sub(hi);
        return r;
    }
    // ---- MODE 5: one level of Karatsuba on the a*b half --------------------------------------------------------------
    // Every 32x32->64 multiply costs 4 cycles of the FMA-heavy pipe and that pipe is what bounds the MSM
    // (profiles/r01_multiplier_variants.md), while the ALU pipe idles: so trade multiplies for additions.  The product
    // a*b is formed from three half-size products, (a0 + a1)(b0 + b1) - a0 b0 - a1 b1 for the middle term: 3 (N/2)^2
    // wide multiplies instead of N^2, ~6N additions on the ALU pipe; then the 2N-limb product is Montgomery-reduced row by
    // row with the same staggered carry chains as the interleaved form (N^2 wide multiplies + N low ones, the limbs of the
    // high half entering one per row).  Fp: 108 + 144 = 252 wide multiplies instead of 288; Fr: 48 + 64 = 112 instead of 128.
    static __device__ __forceinline__ void redc(Mont& r, const uint32_t* T) {
        constexpr int H = N / 2;
        uint32_t E[N + 2], O[N + 2];
        uint32_t mod_[N];
#pragma unroll
        for (int i = 0; i < N; i++) { mod_[i] = P::mod(i); E[i] = T[i]; O[i] = 0; }
        E[N] = 0;
#pragma unroll
        for (int i = 0;; i++) {
            // the last limb of the product joins before the last row (T < mod^2: it cannot overflow the carry limb)
            if (i == N - 1) E[N] += T[2 * N - 1];
            uint32_t m = E[0] * P::INV;
            if (i == 0) Chain<H, false>::mad(O, mod_ + 1, m);
            else Chain<H, true>::mad(O, mod_ + 1, m);     // carry-in: the stray word added to E[0] below
            Chain<H, false>::mad(E, mod_, m);
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N]));
            if (i == N - 1) break;
            // S >>= 32:  E' = O (+ stray E[1]),  O'[k] = E[k+2]; then limb N + i of the product enters at limb N - 1
            uint32_t stray = E[1];
            uint32_t nE[N + 2], nO[N + 2];
#pragma unroll
            for (int k = 0; k < N; k++) nE[k] = O[k];
#pragma unroll
            for (int k = 0; k < N - 1; k++) nO[k] = E[k + 2];
            nO[N - 1] = 0;
#pragma unroll
            for (int k = 0; k < N; k++) { E[k] = nE[k]; O[k] = nO[k]; }
            asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, 0, 0;" : "+r"(E[N - 1]), "=r"(E[N]) : "r"(T[N + i]));
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(stray));
        }
        uint32_t hi;
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(O[0]), "r"(E[1]));
#pragma unroll
        for (int k = 1; k < N; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.v[k]) : "r"(O[k]), "r"(E[k + 1]));
        asm volatile("addc.u32 %0, 0, 0;" : "=r"(hi));
        r.final_sub(hi);
    }
    static __device__ __forceinline__ Mont mul_kara(const Mont& a, const Mont& b) {
        constexpr int H = N / 2;
        uint32_t P0[N], P2[N], P1[N + 1], sa[H], sb[H], ca, cb;
        wide_mul<H>(P0, a.v, b.v);
        wide_mul<H>(P2, a.v + H, b.v + H);
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(sa[0]) : "r"(a.v[0]), "r"(a.v[H]));
#pragma unroll
        for (int i = 1; i < H; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(sa[i]) : "r"(a.v[i]), "r"(a.v[H + i]));
        asm volatile("addc.u32 %0, 0, 0;" : "=r"(ca));
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(sb[0]) : "r"(b.v[0]), "r"(b.v[H]));
#pragma unroll
        for (int i = 1; i < H; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(sb[i]) : "r"(b.v[i]), "r"(b.v[H + i]));
        asm volatile("addc.u32 %0, 0, 0;" : "=r"(cb));
        wide_mul<H>(P1, sa, sb);
        // (ca 2^(32H) + sa)(cb 2^(32H) + sb): the carry bits add sb, sa and 1 one half-width up
        {
            const uint32_t ma = 0u - ca, mb = 0u - cb;
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(P1[H]) : "r"(sb[0] & ma));
#pragma unroll
            for (int i = 1; i < H; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(P1[H + i]) : "r"(sb[i] & ma));
            asm volatile("addc.u32 %0, %1, 0;" : "=r"(P1[N]) : "r"(ca & cb));
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(P1[H]) : "r"(sa[0] & mb));
#pragma unroll
            for (int i = 1; i < H; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(P1[H + i]) : "r"(sa[i] & mb));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(P1[N]));
        }
        // middle term M = P1 - P0 - P2 (N + 1 limbs, non-negative)
        asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(P1[0]) : "r"(P0[0]));
#pragma unroll
        for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(P1[i]) : "r"(P0[i]));
        asm volatile("subc.u32 %0, %0, 0;" : "+r"(P1[N]));
        asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(P1[0]) : "r"(P2[0]));
#pragma unroll
        for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(P1[i]) : "r"(P2[i]));
        asm volatile("subc.u32 %0, %0, 0;" : "+r"(P1[N]));
        // T = P0 + (M << 32H) + (P2 << 32N)
        uint32_t T[2 * N];
#pragma unroll
        for (int i = 0; i < H; i++) T[i] = P0[i];
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(T[H]) : "r"(P0[H]), "r"(P1[0]));
#pragma unroll
        for (int i = 1; i < H; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(T[H + i]) : "r"(P0[H + i]), "r"(P1[i]));
#pragma unroll
        for (int i = 0; i < H; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(T[N + i]) : "r"(P2[i]), "r"(P1[H + i]));
        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(T[N + H]) : "r"(P2[H]), "r"(P1[N]));
#pragma unroll
        for (int i = 1; i < H - 1; i++) asm volatile("addc.cc.u32 %0, %1, 0;" : "=r"(T[N + H + i]) : "r"(P2[H + i]));
        asm volatile("addc.u32 %0, %1, 0;" : "=r"(T[2 * N - 1]) : "r"(P2[N - 1]));
        Mont r;
        redc(r, T);
        return r;
    }
    static __device__ __forceinline__ Mont mul_unrolled(const Mont& a, const Mont& b) {
        constexpr int H = N / 2;
        uint32_t E[N + 2], O[N + 2];  // E uses N+1 limbs, O uses N; +1 so that the renaming below stays in bounds
        uint32_t mod_[N];
#pragma unroll
        for (int i = 0; i < N; i++) mod_[i] = P::mod(i);

        // row 0: T = a * b[0]
        Chain<H, false>::mul(E, a.v, b.v[0]);
        Chain<H, false>::mul(O, a.v + 1, b.v[0]);
        E[N] = 0;
#pragma unroll
        for (int i = 0;; i++) {
            // reduce: T += m * mod, m chosen so that the low word becomes zero
            uint32_t m = E[0] * P::INV;
            Chain<H, false>::mad(O, mod_ + 1, m);  // no carry out: (O << 32) <= T < 2^(32(N+1))
            Chain<H, false>::mad(E, mod_, m);
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N]));
            if (i == N - 1) break;
            // T >>= 32:  E' = O (+ stray E[1]),  O'[k] = E[k+2]
            uint32_t stray = E[1];
            uint32_t nE[N + 2], nO[N + 2];
#pragma unroll
            for (int k = 0; k < N; k++) nE[k] = O[k];
#pragma unroll
            for (int k = 0; k < N - 1; k++) nO[k] = E[k + 2];
            nO[N - 1] = 0;
#pragma unroll
            for (int k = 0; k < N; k++) { E[k] = nE[k]; O[k] = nO[k]; }
            // next row: T += a * b[i+1]; the stray word's carry enters the O chain, one limb up
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(stray));
            Chain<H, true>::mad(O, a.v + 1, b.v[i + 1]);
            Chain<H, false>::mad(E, a.v, b.v[i + 1]);
            asm volatile("addc.u32 %0, 0, 0;" : "=r"(E[N]));
        }
        // result = (E >> 32) + O, then one conditional subtraction
        Mont r;
        uint32_t hi;
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(O[0]), "r"(E[1]));
#pragma unroll
        for (int k = 1; k < N; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.v[k]) : "r"(O[k]), "r"(E[k + 1]));
        asm volatile("addc.u32 %0, 0, 0;" : "=r"(hi));
        r.final_sub(hi);
        return r;
    }
    __device__ __forceinline__ Mont sqr() const {
        if (MODE == MONT_R28) return mul_r28(*this, *this, true);
        return *this * *this;
    }

    // a^e for an exponent array (little-endian u32 words): fixed 4-bit windows over a table of a^1 .. a^15 -- for the
    // 379-bit square-root exponent (p + 1) / 4 that is 379 squarings + ~90 products + 14 for the table instead of the ~190
    // products of plain square-and-multiply (every use is a dependent chain: point decoding, the Fermat cross-check)
    template <int W>
    __device__ __noinline__ Mont pow_words(const uint32_t (&e)[W]) const {
        Mont t[16];
        t[0] = one();
        t[1] = *this;
#pragma unroll 1
        for (int k = 2; k < 16; k++) t[k] = t[k - 1] * *this;
        Mont acc = one();
        bool started = false;
#pragma unroll 1
        for (int i = W * 8 - 1; i >= 0; i--) {
            const uint32_t d = (e[i >> 3] >> ((i & 7) * 4)) & 15u;
            if (started) {
                acc = acc.sqr();
                acc = acc.sqr();
                acc = acc.sqr();
                acc = acc.sqr();
                if (d) acc = acc * t[d];
            } else if (d) {
                acc = t[d];
                started = true;
            }
        }
        return acc;
    }
    // a^(mod-2) (Fermat); kept as the cross-check of inverse()
    __device__ __noinline__ Mont inverse_fermat() const {
        uint32_t e[N];
#pragma unroll
        for (int i = 0; i < N; i++) e[i] = P::mod(i);
        uint32_t borrow = 2;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t t = e[i];
            e[i] = t - borrow;
            borrow = t < borrow ? 1u : 0u;
        }
        return pow_words(e);
    }
    // Modular inverse by the binary extended Euclid (right-shift) algorithm on the ALU pipe: ~1.5 * bits iterations
    // of shifts / subtractions instead of ~1.5 * bits Montgomery multiplications -- several times lower latency than
    // Fermat, which matters because every use is a latency-bound tail (compression, affine conversion, batch
    // inversion).  Variable time: no secrets on this path.  0 -> 0 like blst_fr_eucl_inverse / blst_fp_inverse.
    // Input is the Montgomery residue x = a R; the integer inverse t = x^-1 = a^-1 R^-1 is lifted back with two
    // multiplications by R^2:  (t R^2 R^-1) R^2 R^-1 = a^-1 R.
    // Default inversion: the optimised binary GCD of inv_bingcd.cuh (31 GCD steps at a time on 64-bit approximations, then
    // one small-matrix update of the full-width numbers): ~5x fewer instructions and much shorter dependency chains than
    // the bit-serial loop below, which stays as inverse_euclid(), its cross-check.  Same contract: Montgomery in and out,
    // 0 -> 0, variable time (no secrets on this path).
    __device__ __noinline__ Mont inverse() const {
        if (is_zero()) return *this;
        Mont t;
        inverse_bingcd<P>(t.v, v);        // (a R)^-1 = a^-1 R^-1 as a plain residue
        return (t * rr()) * rr();         // two Montgomery products by R^2: a^-1 R
    }
    __device__ __noinline__ Mont inverse_euclid() const {
        if (is_zero()) return *this;
        uint32_t u[N], w[N], x1[N], x2[N];  // invariants: x1 * x == u, x2 * x == w (mod m); w stays odd
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = v[i]; w[i] = P::mod(i); x1[i] = 0; x2[i] = 0; }
        x1[0] = 1;
        for (;;) {
            uint32_t nz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) nz |= u[i];
            if (nz == 0) break;
            if ((u[0] & 1) == 0) {
                // u /= 2, x1 /= 2 mod m
#pragma unroll
                for (int i = 0; i < N - 1; i++) u[i] = __funnelshift_r(u[i], u[i + 1], 1);
                u[N - 1] >>= 1;
                uint32_t odd = 0u - (x1[0] & 1), carry;
                asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x1[0]) : "r"(P::mod(0) & odd));
#pragma unroll
                for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x1[i]) : "r"(P::mod(i) & odd));
                asm volatile("addc.u32 %0, 0, 0;" : "=r"(carry));
#pragma unroll
                for (int i = 0; i < N - 1; i++) x1[i] = __funnelshift_r(x1[i], x1[i + 1], 1);
                x1[N - 1] = __funnelshift_r(x1[N - 1], carry, 1);
            } else {
                // both odd: make u the larger, then u -= w (even), x1 -= x2 mod m
                bool lt = false;
#pragma unroll
                for (int i = N - 1; i >= 0; i--) {
                    if (u[i] != w[i]) { lt = u[i] < w[i]; break; }
                }
                if (lt) {
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        uint32_t t = u[i]; u[i] = w[i]; w[i] = t;
                        t = x1[i]; x1[i] = x2[i]; x2[i] = t;
                    }
                }
                asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(u[0]) : "r"(w[0]));
#pragma unroll
                for (int i = 1; i < N - 1; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(w[i]));
                asm volatile("subc.u32 %0, %0, %1;" : "+r"(u[N - 1]) : "r"(w[N - 1]));
                uint32_t borrow;
                asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(x1[0]) : "r"(x2[0]));
#pragma unroll
                for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(x1[i]) : "r"(x2[i]));
                asm volatile("subc.u32 %0, 0, 0;" : "=r"(borrow));
                asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x1[0]) : "r"(P::mod(0) & borrow));
#pragma unroll
                for (int i = 1; i < N - 1; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x1[i]) : "r"(P::mod(i) & borrow));
                asm volatile("addc.u32 %0, %0, %1;" : "+r"(x1[N - 1]) : "r"(P::mod(N - 1) & borrow));
            }
        }
        // gcd = w = 1 (the modulus is prime), inverse = x2
        Mont t;
#pragma unroll
        for (int i = 0; i < N; i++) t.v[i] = x2[i];
        return (t * rr()) * rr();
    }
    __device__ __forceinline__ Mont to_mont() const { return *this * rr(); }  // canonical -> Montgomery
    __device__ __forceinline__ Mont from_mont() const {                        // Montgomery -> canonical
        Mont o = zero();
        o.v[0] = 1;
        return *this * o;
    }
};

// All variants share one data layout and give identical results.  The default is the unrolled carry-chain
// multiplier: measured on B200 (scripts/ubench, profiles/r01_multiplier_variants.md) EVERY IMAD.WIDE form issues at
// 32/clk/SM, with or without carry, so the radix-2^28 variant -- more, carry-free products -- is slower (21.2 vs
// 29.8 G Fp-mul/s) even though both keep the FMA-heavy pipe > 90 % busy.  It stays as a tested cross-check.
typedef Mont<FpParams, MONT_UNROLLED> fp_t;
typedef Mont<FrParams, MONT_UNROLLED> fr_t;
typedef Mont<FpParams, MONT_COMPACT> fpc_t;   // compact-code carry-chain variant (cold kernels)
typedef Mont<FrParams, MONT_COMPACT> frc_t;
typedef Mont<FpParams, MONT_R28> fp28_t;      // radix-2^28 variant
typedef Mont<FrParams, MONT_R28> fr28_t;
typedef fp_t fpu_t;
typedef fr_t fru_t;
typedef Mont<FpParams, MONT_KARA> fpk_t;      // Karatsuba a*b + row-wise reduction
typedef Mont<FrParams, MONT_KARA> frk_t;
typedef Mont<FpParams, MONT_DFMA> fpd_t;      // a*b on the FP64 pipe, reduction on IMAD.WIDE
typedef Mont<FrParams, MONT_DFMA> frd_t;

// load / store through 128-bit accesses (fp_t = 48 B = 3 x uint4, fr_t = 32 B = 2 x uint4)
template <class F>
__device__ __forceinline__ F load_field(const void* p) {
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) {
        uint4 t = q[i];
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
template <class F>
__device__ __forceinline__ F load_field_ro(const void* p) {  // read-only path (LDG.NC)
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) {
        uint4 t = __ldg(q + i);
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
template <class F>
__device__ __forceinline__ void store_field(void* p, const F& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) q[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}

}  // namespace b200
