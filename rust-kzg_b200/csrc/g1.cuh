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
// Instantiated twice: b200::* uses the unrolled multiplier (throughput kernels), b200::cc::* the compact one
// (latency-bound, instruction-fetch-sensitive kernels).  Memory layouts are identical.
#include "g1_body.inc"
namespace cc {
typedef fpc_t fp_t;
#include "g1_body.inc"
}  // namespace cc
namespace ck {  // Karatsuba product + row-wise reduction (mont.cuh MODE 5): fewer wide multiplies, more ALU work
typedef fpk_t fp_t;
#include "g1_body.inc"
}  // namespace ck
namespace cl {  // multiplier behind a call: small loop bodies for the instruction cache
typedef Mont<FpParams, MONT_CALL> fp_t;
#include "g1_body.inc"
}  // namespace cl
}  // namespace b200
