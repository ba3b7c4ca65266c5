/*
 * b200_kzg.h -- C ABI of libb200kzg.so, the B200 (sm_100a) backend for the rust-kzg hot path.
 *
 * Every entry point replaces (or extends) one interface of the reference, cited as file:line under
 * grandinetech/rust-kzg @ 62f9fd85.  Plain pointers and sizes only; host pointers unless a name says _device.
 * All field/curve data uses blst's memory layouts (kzg/src/eth/c_bindings.rs:427-474): little-endian u64 limbs,
 * Montgomery form.  There is NO CPU fallback: every compute entry point needs a CUDA device and fails loudly
 * (error code / C_KZG_ERROR) when none is usable.
 */
#ifndef B200_KZG_H
#define B200_KZG_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_API __attribute__((visibility("default")))

/* ---- blst data layouts (kzg/src/eth/c_bindings.rs:427-474) -------------------------------------------------- */
typedef struct { uint64_t l[4]; } blst_fr;          /* Montgomery, R = 2^256 */
typedef struct { uint64_t l[6]; } blst_fp;          /* Montgomery, R = 2^384 */
typedef struct { blst_fp x, y, z; } blst_p1;        /* Jacobian; infinity <=> z == 0 */
typedef struct { blst_fp x, y; } blst_p1_affine;    /* infinity <=> all zero */
typedef struct { blst_fp fp[2]; } blst_fp2;
typedef struct { blst_fp2 x, y, z; } blst_p2;

/* ============================================================================================================== */
/* B1 -- GPU MSM plug-in FFI.  Same names, argument order and meaning as the sppark shim the blst backend binds   */
/* under --features sppark:  blst-sppark/cuda/pippenger.cu:23-38  <->  blst-sppark/src/lib.rs:8-62.               */
/* ============================================================================================================== */

/* sppark's RustError, returned BY VALUE; code == 0 is success; message is strdup'd and owned by the caller
 * (arkworks3-sppark-wlc/sppark/util/rusterror.h:15-27). */
typedef struct { int code; char *message; } RustError;

/* blst-sppark/cuda/pippenger.cu:23-26.  Uploads the bases and builds the device-resident fixed-base table
 * (rows 2^(c*j) * P_i).  Returns an opaque handle, NULL on failure.  Thread-safe; the handle may be shared
 * between threads (kzg/src/msm/sppark.rs:24-44 declares it Send + Sync). */
B200_API void *prepare_msm(const blst_p1_affine points[], size_t npoints);

/* blst-sppark/cuda/pippenger.cu:28-31.  out = sum scalars[i] * points[i], i < npoints <= prepared npoints.
 * scalars are Montgomery-form blst_fr (the reference passes mont = true). */
B200_API RustError mult_pippenger_prepared(void *msm, blst_p1 *out, size_t npoints, const blst_fr scalars[]);

/* blst-sppark/cuda/pippenger.cu:33-38.  Variable-base MSM; bases travel with the call. */
B200_API RustError mult_pippenger(blst_p1 *out, const blst_p1_affine points[], size_t npoints, const blst_fr scalars[]);

/* Extension: release a prepare_msm handle (the blst shim leaks it; the wlc variant exports a free,
 * arkworks3-sppark-wlc/src/lib.rs:24-42). */
B200_API void b200_free_msm(void *msm);

/* Extension for measurement and for callers that keep data on the GPU: scalars_dev / out_dev are DEVICE pointers
 * (batch * npoints blst_fr, batch blst_p1), stream is a cudaStream_t (NULL = default).  Asynchronous. */
B200_API RustError b200_msm_prepared_device(void *msm, void *out_dev, size_t npoints, const void *scalars_dev, int batch,
                                   void *stream);
/* batched host variant: batch scalar vectors of npoints each over the same prepared bases -- the device side of
 * G1LinComb::g1_lincomb_batch (kzg/src/lib.rs:160-181). */
B200_API RustError b200_msm_prepared_batch(void *msm, blst_p1 out[], size_t npoints, const blst_fr scalars[], int batch);

/* measurement hooks: bracket the accumulate kernel of every run with CUDA events on the launching stream; read
 * returns the summed duration (ms) and the number of runs since the last read */
B200_API void b200_msm_set_profiling(void *msm, int on);
B200_API RustError b200_msm_profile_read(void *msm, double *accumulate_ms_sum, int *runs);
/* out_dev = sum of n Jacobian points (device pointers): the local combine after an all-gather of per-GPU partial
 * MSM results (NCCL has no reduction over curve points) */
B200_API RustError b200_g1_sum_device(void *out_dev, const void *points_dev, size_t n, void *stream);

/* introspection (bench / tests): window bits, windows, table bytes, kernel launches of the last run */
B200_API void b200_msm_info(void *msm, int *c, int *W, size_t *table_bytes, int *launches);
/* the window plan a handle for `npoints` bases would get (host-only, no device needed): window bits c, width of window 0,
 * number of windows W (c0 + (W-1) c >= 256), bucket-index bits folded by segments ahead of the reduce (c > 16) */
B200_API void b200_msm_plan(size_t npoints, int fixed, int *c, int *c0, int *W, int *fold_bits);
/* work of the last run on this handle: entries = non-zero digits sorted into buckets, tasks = accumulate tasks; the
 * accumulate kernel did  entries - tasks  mixed additions (a task's first point is a load).  Synchronises the handle's
 * stream. */
B200_API RustError b200_msm_last_counts(void *msm, size_t *entries, size_t *tasks);
/* all work counters of the last run, for an exact addition count (bench.py): stats = [entries, tasks, non-empty buckets,
 * bucket keys, segment-fold bits, buckets per group the marginal reduce sees, digit axes of that reduce, groups] */
B200_API RustError b200_msm_last_stats(void *msm, uint64_t stats[8]);
/* 1 when the last run on this handle used the batch-affine bucket accumulation (k_accumulate_affine: 6 field
 * multiplications per addition + a shared inversion), 0 for the XYZZ task kernel (10 per addition) */
B200_API int b200_msm_last_affine(void *msm);
/* 1 when the prepared table is scalar-randomised: every scalar is multiplied by a per-base pseudo-random rho_i and the table
 * holds rho_i^-1 * P_i, so the digit distribution -- and the running time -- does not depend on the caller's scalars.  Applied
 * at prepare when every base is in the prime-order subgroup (B200_MSM_RANDOMIZE=0 turns it off); the result is the same
 * group element either way. */
B200_API int b200_msm_randomized(void *msm);
/* Fixed-base handles of 1024..8192 points (the 4096-point MSM behind g1_lincomb, kzg/src/eip_4844.rs:463-476) also hold a
 * direct-lookup table of every signed-digit multiple of every window of every base: a full-length call is then
 * the plain sum of npoints * W looked-up points in ONE launch (0.2 ms instead of 0.55 ms for 4096 points), shorter calls
 * and every other size use the bucket pipeline.  Returns the table's window width (B200_MSM_DIRECT_BITS, default 11 =
 * 7.5 GiB for 4096 points; narrower when HBM is short), 0 when the handle has none (B200_MSM_DIRECT=0). */
B200_API int b200_msm_direct_bits(void *msm);


/* ============================================================================================================== */
/* NTT -- FFTSettings / FFTFr / DASExtension (kzg/src/lib.rs:421-431, 465-481) as implemented by the blst backend  */
/* (blst/src/types/fft_settings.rs:28-58, blst/src/fft_fr.rs:112-165, blst/src/data_availability_sampling.rs:78-100) */
/* Errors: code 1 carries the reference's Err(String) text (bad length etc.); other codes are CUDA errors.        */
/* ============================================================================================================== */

/* FsFFTSettings::new(scale): roots_of_unity[0..=2^scale] resident on the device. NULL on failure (scale >= 32, no GPU). */
B200_API void *b200_fft_settings_new(int scale);
B200_API void b200_fft_settings_free(void *fs);
B200_API size_t b200_fft_settings_max_width(void *fs);
/* copy a roots table to the host: which = 0 roots_of_unity (max_width+1), 1 brp_roots_of_unity (max_width),
 * 2 reverse_roots_of_unity (max_width+1)   (FFTSettings getters, kzg/src/lib.rs:465-481) */
B200_API RustError b200_fft_settings_roots(void *fs, int which, blst_fr *out);

/* FFTFr::fft_fr(data, inverse) -> out; n must be a power of two <= max_width; natural order in and out. */
B200_API RustError b200_fft_fr(void *fs, blst_fr *out, const blst_fr *in, size_t n, bool inverse);
/* DASExtension::das_fft_extension(evens) -> odds; n = len(evens), 2n <= max_width. */
B200_API RustError b200_das_fft_extension(void *fs, blst_fr *odds, const blst_fr *evens, size_t n);
/* device-pointer, batched (batch contiguous transforms of n), asynchronous on stream; out_dev != in_dev */
B200_API RustError b200_fft_fr_device(void *fs, void *out_dev, const void *in_dev, size_t n, int inverse, int batch, void *stream);
B200_API RustError b200_das_fft_extension_device(void *fs, void *odds_dev, const void *evens_dev, size_t n, int batch, void *stream);
/* FFTG1::fft_g1(data, inverse) (kzg/src/lib.rs:425-427, blst/src/fft_g1.rs:53-83): transform over G1 points */
B200_API RustError b200_fft_g1(void *fs, blst_p1 *out, const blst_p1 *in, size_t n, bool inverse);
B200_API RustError b200_fft_g1_device(void *fs, void *out_dev, const void *in_dev, size_t n, int inverse, int batch, void *stream);
B200_API int b200_fft_launches(void *fs);

/* ============================================================================================================== */
/* B2 -- c-kzg-4844 C ABI (pinned upstream 00ae727c, .github/workflows/backend-tests.yml:4), commitment / proof     */
/* and verification path.  Replaces blst/src/eip_4844.rs:160-530 for the functions below; types from                */
/* kzg/src/eth/c_bindings.rs:16-113.                                                                                */
/* ============================================================================================================== */
typedef enum { C_KZG_OK = 0, C_KZG_BADARGS = 1, C_KZG_ERROR = 2, C_KZG_MALLOC = 3 } C_KZG_RET;  /* c_bindings.rs:16-23 */
typedef struct { uint8_t bytes[32]; } Bytes32;
typedef struct { uint8_t bytes[48]; } Bytes48;
typedef struct { uint8_t bytes[131072]; } Blob;
typedef Bytes48 KZGCommitment;
typedef Bytes48 KZGProof;
/* CKZGSettings (kzg/src/eth/c_bindings.rs:55-108).  All arrays are host memory owned by the library between
 * load_trusted_setup* and free_trusted_setup; the device context is found through g1_values_lagrange_brp. */
typedef struct {
    blst_fr *roots_of_unity;          /* 8193 */
    blst_fr *brp_roots_of_unity;      /* 8192 */
    blst_fr *reverse_roots_of_unity;  /* 8193 */
    blst_p1 *g1_values_monomial;      /* 4096 */
    blst_p1 *g1_values_lagrange_brp;  /* 4096 */
    blst_p2 *g2_values_monomial;      /* 65 */
    blst_p1 **x_ext_fft_columns;      /* [128][64] FK20 Toeplitz columns (blst/src/types/kzg_settings.rs:84-101), one heap row each */
    blst_p1_affine **tables;          /* NULL: the reference's CPU precomputation; this backend's tables live in HBM */
    size_t wbits;
    size_t scratch_size;
} KZGSettings;
typedef KZGSettings CKZGSettings;

/* blst/src/eip_4844.rs:180-222.  Decompresses the points on the GPU, builds the fixed-base MSM table and the
 * roots-of-unity tables, decodes the 65 G2 points and tabulates the Miller-loop lines of [1]G2, [s]G2, [s^64]G2, fills
 * the host arrays.  BADARGS on wrong counts / undecodable points / a monomial-form array in the Lagrange slot (the
 * pairing check of kzg/src/eip_4844.rs:1005-1020, 1064-1068, run on the device). */
B200_API C_KZG_RET load_trusted_setup(KZGSettings *out, const uint8_t *g1_monomial_bytes, uint64_t num_g1_monomial_bytes,
                             const uint8_t *g1_lagrange_bytes, uint64_t num_g1_lagrange_bytes,
                             const uint8_t *g2_monomial_bytes, uint64_t num_g2_monomial_bytes, uint64_t precompute);
/* blst/src/eip_4844.rs:227-269 (text format of kzg/src/eip_4844.rs:151-228) */
B200_API C_KZG_RET load_trusted_setup_file(KZGSettings *out, FILE *in);
/* blst/src/eip_4844.rs:296-378: frees and NULLs every array, drops the device context; NULL-safe */
B200_API void free_trusted_setup(KZGSettings *s);
/* blst/src/eip_4844.rs:163-175 */
B200_API C_KZG_RET blob_to_kzg_commitment(KZGCommitment *out, const Blob *blob, const KZGSettings *s);
/* blst/src/eip_4844.rs:476-496 */
B200_API C_KZG_RET compute_kzg_proof(KZGProof *proof_out, Bytes32 *y_out, const Blob *blob, const Bytes32 *z_bytes, const KZGSettings *s);
/* blst/src/eip_4844.rs:274-291 */
B200_API C_KZG_RET compute_blob_kzg_proof(KZGProof *out, const Blob *blob, const Bytes48 *commitment_bytes, const KZGSettings *s);

/* blst/src/eip_4844.rs:383-405: *ok = e(C - [y]G1, G2) == e(proof, [s]G2 - [z]G2) (check_proof_single,
 * blst/src/types/kzg_settings.rs:178-196).  BADARGS for non-canonical z / y, malformed, off-curve or out-of-subgroup
 * points.  The pairing runs on the device against line tables of the fixed G2 points built at load time. */
B200_API C_KZG_RET verify_kzg_proof(bool *ok, const Bytes48 *commitment_bytes, const Bytes32 *z_bytes, const Bytes32 *y_bytes,
                                    const Bytes48 *proof_bytes, const KZGSettings *s);
/* blst/src/eip_4844.rs:410-430 */
B200_API C_KZG_RET verify_blob_kzg_proof(bool *ok, const Blob *blob, const Bytes48 *commitment_bytes, const Bytes48 *proof_bytes,
                                         const KZGSettings *s);
/* blst/src/eip_4844.rs:435-471 (*ok preset to false; n == 0 is true): challenges hashed on the host, evaluations and
 * one random-linear-combination pairing check (kzg/src/eip_4844.rs:328-435) on the device */
B200_API C_KZG_RET verify_blob_kzg_proof_batch(bool *ok, const Blob *blobs, const Bytes48 *commitments_bytes, const Bytes48 *proofs_bytes,
                                               size_t n, const KZGSettings *s);
/* extension: verify_kzg_proof_batch (kzg/src/eip_4844.rs:380-435) over caller-supplied (C, z, y, proof) tuples */
B200_API C_KZG_RET b200_verify_kzg_proof_batch(bool *ok, const Bytes48 *commitments, const Bytes32 *zs, const Bytes32 *ys,
                                               const Bytes48 *proofs, size_t n, const KZGSettings *s);
/* test hook: out = sum scalars[i] * points[i] through the lane-quad GLV scalar multiplication (verifiers, fft_g1) */
B200_API RustError b200_selftest_lincomb_quads(blst_p1 *out, const blst_p1_affine *points, const blst_fr *scalars, size_t n);
/* test hook: e(a1, Q[qa]) == e(b1, Q[qb]) with Q = {[1]G2, [s]G2, [s^64]G2} (pairings_verify, blst/src/kzg_proofs.rs:74-100) */
B200_API C_KZG_RET b200_selftest_pairings_verify(bool *ok, const blst_p1 *a1, int qa, const blst_p1 *b1, int qb, const KZGSettings *s);

/* blst/src/eip_4844.rs:498-530: helper exports of the reference's C ABI (used by its binding tests).  compute_challenge
 * writes the Fiat-Shamir challenge (Montgomery blst_fr) of a blob and a Jacobian commitment; the blob must be valid (the
 * reference unwraps).  bytes_to_kzg_commitment = G1::from_bytes (no subgroup check); bytes_from_bls_field = Fr::to_bytes. */
B200_API void compute_challenge(blst_fr *eval_challenge_out, const Blob *blob, const blst_p1 *commitment);
B200_API C_KZG_RET bytes_to_kzg_commitment(blst_p1 *out, const Bytes48 *b);
B200_API void bytes_from_bls_field(Bytes32 *out, const blst_fr *inp);

/* EIP-7594 compute_cells_and_kzg_proofs (kzg/src/das.rs:244-292; C ABI kzg/src/eth/c_bindings.rs:134-199):
 * cells = BRP(NTT_8192(INTT_4096(BRP(blob)))); proofs by FK20 (64 x NTT_128, 128 fixed-base lincombs of 64 points over
 * x_ext_fft_columns, inverse + forward fft_g1 of size 128).  Either output pointer may be NULL, not both. */
typedef struct { uint8_t bytes[2048]; } Cell;
B200_API C_KZG_RET compute_cells_and_kzg_proofs(Cell *cells, KZGProof *proofs, const Blob *blob, const KZGSettings *s);
/* kzg/src/eth/c_bindings.rs:201-286 (DAS::recover_cells_and_kzg_proofs, kzg/src/das.rs:101-207): 64..128 cells with strictly
 * ascending indices -> all 128 cells and, unless recovered_proofs is NULL, their 128 proofs */
B200_API C_KZG_RET recover_cells_and_kzg_proofs(Cell *recovered_cells, KZGProof *recovered_proofs, const uint64_t *cell_indices,
                                                const Cell *cells, uint64_t num_cells, const KZGSettings *s);
/* kzg/src/eth/c_bindings.rs:288-352 (DAS::verify_cell_kzg_proof_batch, kzg/src/das.rs:294-388) */
B200_API C_KZG_RET verify_cell_kzg_proof_batch(bool *ok, const Bytes48 *commitments_bytes, const uint64_t *cell_indices, const Cell *cells,
                                               const Bytes48 *proofs_bytes, uint64_t num_cells, const KZGSettings *s);
/* blst/src/eip_7594.rs:35-97: the Fiat-Shamir challenge of the cell batch verifier as a Montgomery blst_fr.  Needs a loaded
 * trusted setup (its device context runs the argument checks); C_KZG_ERROR otherwise. */
B200_API C_KZG_RET compute_verify_cell_kzg_proof_batch_challenge(blst_fr *challenge_out, const Bytes48 *commitment_bytes,
                                                                 uint64_t num_commitments, const uint64_t *commitment_indices,
                                                                 const uint64_t *cell_indices, const Cell *cells,
                                                                 const Bytes48 *proofs_bytes, uint64_t num_cells);
B200_API C_KZG_RET b200_compute_cells_batch(Cell *cells, const Blob *blobs, size_t n, const KZGSettings *s);
B200_API C_KZG_RET b200_compute_cell_proofs_batch(KZGProof *proofs, const Blob *blobs, size_t n, const KZGSettings *s);
/* both outputs of compute_cells_and_kzg_proofs (kzg/src/das.rs:244-292) for n blobs in one pass: the blobs are uploaded and
 * brought to monomial form once, and the cells are copied out while the proofs are still being computed.  Either output
 * pointer may be NULL, not both; cells = n x 128 Cell, proofs = n x 128 KZGProof.  The single-blob symbol is this with n = 1. */
B200_API C_KZG_RET b200_compute_cells_and_kzg_proofs_batch(Cell *cells, KZGProof *proofs, const Blob *blobs, size_t n,
                                                           const KZGSettings *s);

/* Batched extensions: n independent blobs in one launch sequence (BASELINE config 3: 64 blobs).  Any invalid
 * element makes the whole call return C_KZG_BADARGS.  n may exceed the context's batch capacity (chunked). */
B200_API C_KZG_RET b200_blob_to_kzg_commitment_batch(KZGCommitment *out, const Blob *blobs, size_t n, const KZGSettings *s);
B200_API C_KZG_RET b200_compute_kzg_proof_batch(KZGProof *proofs, Bytes32 *ys, const Blob *blobs, const Bytes32 *zs, size_t n, const KZGSettings *s);
B200_API C_KZG_RET b200_compute_blob_kzg_proof_batch(KZGProof *out, const Blob *blobs, const Bytes48 *commitments, size_t n, const KZGSettings *s);
/* Device-pointer extensions (asynchronous on stream; status_dev[i] != 0 marks an invalid blob / z; n <= capacity) */
B200_API C_KZG_RET b200_blob_to_kzg_commitment_device(void *out48_dev, const void *blobs_dev, size_t n, int *status_dev, const KZGSettings *s, void *stream);
B200_API C_KZG_RET b200_compute_kzg_proof_device(void *proofs48_dev, void *y32_dev, const void *blobs_dev, const void *z32_dev, size_t n, int z_reduce, int *status_dev, const KZGSettings *s, void *stream);
B200_API int b200_kzg_launches(const KZGSettings *s);
B200_API int b200_kzg_max_batch(const KZGSettings *s);
/* Concurrent calls of the single-blob functions above are coalesced into shared launch sequences (csrc/coalesce.cuh).
 * Counters since load: out = [batches run, requests served, ns leaders waited for a device lane, ns batches spent on a
 * lane, largest batch].  B200_KZG_COALESCE=1 disables packing, B200_KZG_LANES=k limits the lanes single calls may use. */
B200_API void b200_kzg_coalesce_stats(const KZGSettings *s, uint64_t out[5]);
/* compute_cells_and_kzg_proofs (with proofs) is coalesced the same way, in batches of up to 16 blobs (B200_KZG_CELLS_COALESCE;
 * 1 disables it): a block's blobs under a parallel iterator share one FK20 pass.  out = [batches run, requests served]. */
B200_API void b200_kzg_cells_coalesce_stats(const KZGSettings *s, uint64_t out[2]);
/* verify_blob_kzg_proof / verify_kzg_proof called one at a time from several threads are checked as ONE batch (the random
 * linear combination of the reference's verify_blob_kzg_proof_batch, kzg/src/eip_4844.rs:380-435); when a batch does not pass
 * as a whole, or an input does not decode, every caller in it is re-checked alone, so each gets exactly its own answer and
 * status.  A call without concurrent company is the plain single check.  B200_KZG_VERIFY_COALESCE (default 32; 1 disables).
 * out = [batches checked, requests served, batches whose callers had to be re-checked alone]. */
B200_API void b200_kzg_verify_coalesce_stats(const KZGSettings *s, uint64_t out[3]);
/* Direct-lookup tables held by a settings object: out = [window bits of the Lagrange-point table (13 by default,
 * 11 / 8 when HBM is short, 0 = none: bucket engine), largest blob batch it serves, window bits of the FK20 column table (0
 * before the first cell-proof call)].  Knobs: B200_BLOB_DIRECT, B200_BLOB_DIRECT_BITS, B200_FK20_DIRECT, B200_FK20_DIRECT_BITS,
 * B200_DIRECT_RESERVE_GB (HBM that must stay free beside a table, default 40). */
B200_API void b200_kzg_direct_tables(const KZGSettings *s, int out[3]);
B200_API void b200_selftest_sha256(uint8_t out[32], const uint8_t *msg, size_t len, int portable);

/* ---- device self-test hooks: elementwise field / point kernels on host arrays, used by the parity tests ------- */
/* op: 0 mul, 1 add, 2 sub, 3 neg(a), 4 inverse(a), 5 to-Montgomery(a), 6 from-Montgomery(a) */
B200_API RustError b200_selftest_fp(int op, blst_fp *out, const blst_fp *a, const blst_fp *b, size_t n);
B200_API RustError b200_selftest_fr(int op, blst_fr *out, const blst_fr *a, const blst_fr *b, size_t n);
/* out[i] = a[i] + b[i] (Jacobian in/out, through the XYZZ mixed / full adders; mixed != 0 converts b to affine first) */
B200_API RustError b200_selftest_p1_add(blst_p1 *out, const blst_p1 *a, const blst_p1 *b, size_t n, int mixed);
/* out48[i] = compress(p[i]) */
B200_API RustError b200_selftest_p1_compress(uint8_t *out48, const blst_p1 *p, size_t n);
/* integer-pipe microbenchmark: returns achieved 32x32->64 multiply-adds per second (all SMs) through *imad_per_s
 * and the same for a stream of dependent Fp multiplications through *fpmul_per_s */
B200_API RustError b200_microbench_int(double *imad_per_s, double *fpmul_per_s);
/* field multiplications per second of one multiplier variant (dependent chains, all SMs): field 0 = Fp, 1 = Fr;
 * mode 0 = interleaved CIOS, 5 = Karatsuba product + row-wise reduction.  Variants are selected in the element-wise hooks
 * above by adding 16 (rolled loop), 32 (radix 2^28), 64 (FP64 product) or 128 (Karatsuba) to op. */
B200_API RustError b200_microbench_mul(int field, int mode, double *mul_per_s);

/* ---- multi-GPU MSM (BASELINE configs[4]): one process per GPU, terms sharded by rank ----------
 * The reference has no multi-GPU code (its sppark plug uses device 0 only, arkworks3-sppark-wlc/sppark/msm/pippenger.cuh:573-575);
 * this is the north star's own contract.  Rank 0 creates an NCCL unique id and ships it to the other ranks by any means
 * (MPI, a file, torch.distributed); every rank then prepares ITS slice of the bases on its current device.  A mult call
 * takes the rank's slice of the scalars and leaves the FULL sum on every rank: local MSM -> ncclAllGather of the 144-byte
 * partial results -> one-warp quad-tree add, all on one stream.  world == 1 needs neither an id nor NCCL. */
B200_API int b200_msm_sharded_unique_id(uint8_t id[128]);   /* 0 on success */
B200_API void *b200_msm_sharded_prepare(const blst_p1_affine local_points[], size_t n_local, int rank, int world, const uint8_t id[128]);
B200_API RustError b200_msm_sharded_mult(void *sharded, blst_p1 *out, size_t n_local, const blst_fr local_scalars[]);
B200_API RustError b200_msm_sharded_mult_device(void *sharded, void *out_dev, size_t n_local, const void *scalars_dev, void *stream);
B200_API void *b200_msm_sharded_local(void *sharded);        /* the rank-local prepared handle (b200_msm_info, profiling) */
B200_API void b200_msm_sharded_free(void *sharded);

/* microbenchmark behind profiles/r02_affine.md: the kernel time of npairs independent batch-affine point additions (K per
 * thread, two gathered points per pair, one field inversion per 128 K pairs) over a table of npoints points */
B200_API RustError b200_bench_affine_pairs(uint32_t npoints, size_t npairs, int K, double *ms);

/* number of CUDA devices usable; 0 means every compute entry point will fail */
B200_API int b200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_KZG_H */
