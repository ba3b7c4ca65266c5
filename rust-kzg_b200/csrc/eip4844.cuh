// eip4844.cuh -- device-resident KZGSettings and the batched EIP-4844 commitment / proof pipeline.
//
// Mirrors, for the hot path only, the reference's FsKZGSettings (blst/src/types/kzg_settings.rs:66-136) and
// kzg::eip_4844::{blob_to_kzg_commitment_rust, compute_kzg_proof_rust, compute_blob_kzg_proof_rust}
// (kzg/src/eip_4844.rs:278-295, 437-519, 541-563).  A batch of blobs is one launch sequence; every blob owns one
// bucket set of the fixed-base MSM over the bit-reversed Lagrange points.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <memory>

#include "msm.cuh"
#include "ntt.cuh"

namespace b200 {

constexpr size_t kFieldElementsPerBlob = 4096;  // kzg/src/eip_4844.rs:32
constexpr size_t kBytesPerBlob = 131072;

class KzgSettingsDev {
public:
    // g1_monomial / g1_lagrange: 4096 x 48-byte compressed points each (host), as parsed from the trusted setup
    // (kzg/src/eip_4844.rs:151-228).  Throws CudaError(code 1) on malformed / off-curve points.
    KzgSettingsDev(const uint8_t* g1_monomial, const uint8_t* g1_lagrange, int max_batch, cudaStream_t stream);
    ~KzgSettingsDev();
    KzgSettingsDev(const KzgSettingsDev&) = delete;

    int max_batch() const { return max_batch_; }
    FFTSettingsDev& fft() { return *fs_; }
    MsmEngine& msm() { return *lanes_[0].msm; }
    // Jacobian (blst_p1) copies for CKZGSettings: bit-reversed Lagrange points and monomial points, device memory
    const void* g1_lagrange_brp_jac_dev() const { return lagrange_jac_; }
    const void* g1_monomial_jac_dev() const { return monomial_jac_; }

    // All pointers below are DEVICE pointers; status[i] (int, device) is set to 1 when blob / argument i is invalid
    // (the reference's Err -> C_KZG_BADARGS); outputs of invalid items are unspecified.
    // Two independent "lanes" (MSM engine + workspace each) let a caller keep two batches in flight on two streams:
    // the latency-bound tail of one batch (bucket reduction, compression) overlaps the accumulation of the next.
    static constexpr int kLanes = 4;
    // blob_to_kzg_commitment_raw (kzg/src/eip_4844.rs:297-314), n <= max_batch
    void blob_to_commitments(const uint8_t* blobs, int n, uint8_t* out48, int* status, cudaStream_t st, int lane = 0);
    // compute_kzg_proof_raw (kzg/src/eip_4844.rs:521-539); z_bytes: n x 32 big-endian; z_reduce: 0 = reject z >= r
    // (Fr::from_bytes), 1 = reduce mod r (hash_to_bls_field, kzg/src/eip_4844.rs:916-918)
    void compute_proofs(const uint8_t* blobs, const uint8_t* z_bytes, int z_reduce, int n, uint8_t* proofs48,
                        uint8_t* y32, int* status, cudaStream_t st, int lane = 0);
    // G1::from_bytes + (is_inf || is_valid) of compute_blob_kzg_proof_rust (kzg/src/eip_4844.rs:556-558)
    void validate_commitments(const uint8_t* commitments48, int n, int* status, cudaStream_t st);
    // cells of compute_cells_and_kzg_proofs(cells, None, blob) (kzg/src/das.rs:244-275): n x 128 cells x 2048 bytes
    void compute_cells(const uint8_t* blobs, int n, uint8_t* cells_out, int* status, cudaStream_t st);
    // FK20 proofs of compute_cells_and_kzg_proofs (kzg/src/das.rs:276-289, 660-696): n x 128 proofs x 48 bytes,
    // n <= fk20_batch().  The 128 x 64 table of x_ext_fft_columns is built on first use.
    void compute_cell_proofs(const uint8_t* blobs, int n, uint8_t* proofs48, int* status, cudaStream_t st);
    // both at once: the blob -> monomial pass is shared; cells_done (may be nullptr) is recorded when cells_out is complete
    void compute_cells_and_proofs(const uint8_t* blobs, int n, uint8_t* cells_out, uint8_t* proofs48, int* status, cudaStream_t st,
                                  cudaEvent_t cells_done);
    int fk20_batch(cudaStream_t st) { ensure_fk20(st); return fk_batch_; }
    int direct_bits() const { return lag_direct_ ? direct_c_ : 0; }
    int direct_max_batch() const { return lag_direct_ ? direct_max_ : 0; }
    int fk_direct_bits() const { return fk_direct_ ? fk_direct_c_ : 0; }
    // the 128 x 64 blst_p1 of FsKZGSettings::x_ext_fft_columns (blst/src/types/kzg_settings.rs:84-101), row-major, into a
    // DEVICE buffer of 128 * 64 * 144 bytes (for the host-side KZGSettings struct; synchronises st)
    void x_ext_fft_columns(void* out_dev, cudaStream_t st);
    void fk20_from_mono(const void* mono, size_t stride, int n, uint8_t* proofs48, cudaStream_t st);

    // ---- verification (verify.cu; kzg/src/eip_4844.rs:328-435, 586-866) -----------------------------------------
    // g2_monomial: 65 x 96-byte compressed points (host).  Decodes them (on-curve check, blst/src/types/g2.rs:50-72)
    // and precomputes the Miller-loop line tables of [1]G2, [s]G2 and [s^64]G2.  Throws CudaError(1) when malformed.
    void load_g2(const uint8_t* g2_monomial, int count, cudaStream_t st);
    bool has_g2() const { return g2_lines_ != nullptr; }
    const void* g2_monomial_jac_dev() const { return g2_jac_; }  // 65 x blst_p2 (288 B)
    // y_i = p_i(z_i) for n <= max_batch blobs; z_bytes as in compute_proofs; z32 / y32: canonical big-endian out
    void evaluate_blobs(const uint8_t* blobs, const uint8_t* z_bytes, int z_reduce, int n, uint8_t* z32, uint8_t* y32,
                        int* status, cudaStream_t st, int lane = 0);
    // verify_kzg_proof_batch (kzg/src/eip_4844.rs:380-435) for any n >= 1: all inputs are device arrays of big-endian
    // wire bytes; r32 = the Fiat-Shamir hash (reduced mod r on the device; ignored for n == 1, where the check is
    // check_proof_single, blst/src/types/kzg_settings.rs:178-196).  status[0] = 1 when an input is invalid
    // (non-canonical field element, malformed / off-curve / out-of-subgroup point); *result = 1 iff the pairing
    // equation holds.
    void verify_batch(const uint8_t* commitments48, const uint8_t* proofs48, const uint8_t* z32, const uint8_t* y32,
                      int z_reduce, const uint8_t* r32, int n, int* status, int* result, cudaStream_t st, bool skip_decode = false);
    // the point-decoding stage of verify_batch on its own (then call verify_batch(..., skip_decode = true) with the same n)
    // decoded: when given, the subgroup test is a second launch and `decoded` is recorded between the two -- a caller
    // whose critical path only needs the coordinates (the lincombs and the pairing) waits on `decoded` and joins the stream
    // before it reads the status
    void verify_decode(const uint8_t* commitments48, const uint8_t* proofs48, int n, int* status, cudaStream_t st,
                       cudaEvent_t decoded = nullptr);
    // ---- EIP-7594 recovery and cell verification (das7594.cu; kzg/src/das.rs:101-207, 294-388) -------------------
    // recover_cells_and_kzg_proofs for one extended blob.  cells: n x 2048 wire bytes (device); cell_idx: their cell
    // indices (host, already validated: n in [64, 128], < 128, strictly ascending).  cells_out: 128 x 2048 bytes,
    // proofs48: 128 x 48 bytes or nullptr (device).  status[0] = 1 when a field element is not canonical.
    void recover_cells(const uint8_t* cells, const uint64_t* cell_idx, int n, uint8_t* cells_out, uint8_t* proofs48, int* status,
                       cudaStream_t st);
    // verify_cell_kzg_proof_batch: commitments48 = m unique commitments (device), comm_idx / cell_idx: n indices each
    // (device, uint32), cells n x 2048, proofs48 n x 48, r32 = Fiat-Shamir hash.  status[0..n) flags invalid inputs.
    void verify_cells(const uint8_t* commitments48, int m, const uint32_t* comm_idx, const uint32_t* cell_idx, const uint8_t* cells,
                      const uint8_t* proofs48, const uint8_t* r32, int n, int* status, int* result, cudaStream_t st);
    void check_challenge_inputs(const uint8_t* commitments48, int m, const uint8_t* cells, const uint8_t* proofs48, int n, int* status,
                                cudaStream_t st);
    // e(a1, Q[qa]) == e(b1, Q[qb]) for Jacobian G1 points on the device, Q[i] in {0: [1]G2, 1: [s]G2, 2: [s^64]G2}
    void pairings_verify(const void* a1_jac, int qa, const void* b1_jac, int qb, int* result, cudaStream_t st);
    int launches_last() const { return launches_; }

private:
    void ensure_fk20(cudaStream_t st);
    std::unique_ptr<MsmEngine> fk_msm_;   // bucket engine of the lincombs, only when there is no direct table
    bool fk_ready_ = false;
    int fk_batch_ = 0;
    void *fk_a_ = nullptr, *fk_b_ = nullptr, *fk_pts_ = nullptr;
    void* fk_direct_ = nullptr;  // every digit multiple of the 8192 column points (fk20_direct.cu), or nullptr
    void* fk_team_part_ = nullptr;     // partial sums / completion counters of the team form of the lincombs (small batches)
    unsigned* fk_team_cnt_ = nullptr;
    int fk_team_max_ = 0;              // largest blob count the team form serves
    int fk_direct_c_ = 0;        // its window width
    int max_batch_;
    int launches_ = 0;
    std::unique_ptr<FFTSettingsDev> fs_;
    struct Lane {
        std::unique_ptr<MsmEngine> msm;    // commitments (blob elements: usually < 2^248, window width 12)
        std::unique_ptr<MsmEngine> msm_q;  // proofs (quotient values: uniform in Fr, window width 13 + segment fold)
        void* scalars = nullptr;   // max_batch * 4096 canonical scalars (MSM input)
        void* poly = nullptr;      // max_batch * 4096 Montgomery field elements
        void* z = nullptr;         // max_batch Montgomery
        void* y = nullptr;         // max_batch Montgomery
        void* out_jac = nullptr;   // max_batch Jacobian results
        void* direct_part = nullptr;  // partial sums of the direct small-batch MSM
        unsigned* direct_cnt = nullptr;  // per-vector completion counters of its fused kernel (zero between launches)
    } lanes_[kLanes];
    void* lag_direct_ = nullptr;   // every digit multiple of the 4096 Lagrange points (direct small-batch MSM), or nullptr
    int direct_max_ = 0;           // largest batch the direct form serves (0: none; larger ones go to the bucket engine)
    int direct_c_ = 0;             // window width of lag_direct_
    // sum over the Lagrange points for n vectors of canonical scalars in ln.scalars -> out48 (compressed); returns launches
    int lagrange_msm(int lane, MsmEngine& eng, int n, uint8_t* out48, cudaStream_t st);
    void* lagrange_jac_ = nullptr;
    void* monomial_jac_ = nullptr;
    void* domain_ = nullptr;    // brp_roots_of_unity[0..4096) of the 8192 table (Montgomery)
    void* cells_a_ = nullptr;   // max_batch * 8192 Fr ping-pong buffers (compute_cells), allocated on first use
    void* cells_b_ = nullptr;
    void* g2_affine_ = nullptr;  // 65 x (x.re, x.im, y.re, y.im)
    void* g2_jac_ = nullptr;     // 65 x blst_p2
    void* g2_lines_ = nullptr;   // 3 x 68 x 288 B line tables
    void* vf_buf_ = nullptr;     // verification workspace, grown on demand
    size_t vf_cap_ = 0;          // terms the workspace holds
    void ensure_verify_ws(size_t n);
    void lincomb2_and_pair(const uint8_t* pts, const uint8_t* scalars, size_t L, uint8_t* partials, uint8_t* sums, uint8_t* scratch,
                           int qa, int qb, int* result, cudaStream_t st);
    void* das_buf_ = nullptr;    // recovery / cell-verification workspace (das7594.cu), grown on demand
    size_t das_bytes_ = 0;
    uint8_t* ensure_das_ws(size_t bytes);
};

// argument checks of compute_verify_cell_kzg_proof_batch_challenge without a settings object (das7594.cu)
size_t check_challenge_ws_bytes(int m, int n);
void launch_check_challenge_inputs(uint8_t* workspace_dev, const uint8_t* commitments48, int m, const uint8_t* cells, const uint8_t* proofs48,
                                   int n, int* status, cudaStream_t st);
// fixed-base lincombs by direct table lookup (fk20_direct.cu): every signed digit multiple of every c-bit window of every
// point; rows = an MSM engine's [W][npts] fixed-base rows for the same window width (W = direct_windows(c))
int direct_windows(int c);
size_t direct_table_bytes(size_t npts, int c);
void launch_direct_build(const void* rows, size_t row_stride, void* table, size_t npts, int c, cudaStream_t st);
// FK20: vector v is a lincomb of the npv points of column block (v mod period); Jacobian results
void launch_direct_lincomb(const void* scalars, const void* table, void* out_jac, int nvec, int period, int npv, int c, cudaStream_t st);
// npts-term MSMs over one point set, compressed results; partials: nvec * 128 * 192 bytes, counters: nvec zeroed words
void launch_direct_msm_compressed(const void* scalars, const void* table, void* partials, unsigned* counters, uint8_t* out48, int nvec,
                                  int npts, int c, cudaStream_t st);
// the same with a choice of result form: out48 (compressed) or, when out_jac != nullptr, blst_p1 Jacobian points
// period > 1: vector v uses the point block (v mod period) of a table over period * npts points (the FK20 columns)
void launch_direct_msm(const void* scalars, const void* table, void* partials, unsigned* counters, uint8_t* out48, uint8_t* out_jac,
                       int nvec, int npts, int c, cudaStream_t st, int period = 1);
void launch_fr_from_mont(const void* in, void* out, size_t n, cudaStream_t st);
// widest window <= want whose table leaves B200_DIRECT_RESERVE_GB free (0: none fits); table of every digit multiple of the
// n * period affine points at points_dev (nullptr when the allocation fails)
int pick_direct_bits(size_t npts, int want);
void* build_direct_table(const void* points_dev, size_t n, int period, int c, cudaStream_t st);
// test hook (verify.cu): sum k_i P_i through the quad / GLV scalar multiplication used by the verifiers and fft_g1
void selftest_lincomb_quads(const void* points_affine_dev, const void* scalars_mont_dev, int n, void* out_jac_dev, cudaStream_t st);
// uncompress n 48-byte points into affine Montgomery form; flags[i] = 1 on malformed / off-curve input
void launch_uncompress_g1(const uint8_t* in48_dev, void* affine_out_dev, int* flags_dev, int n, cudaStream_t st);
// uncompress + subgroup check (G1::from_bytes followed by is_inf() || is_valid()); status[i] = 1 on failure
// (status index = i % status_mod, status_mod = 0 -> n; affine_out_dev may be nullptr)
void launch_decode_g1_checked(const uint8_t* in48_dev, void* affine_out_dev, int* status_dev, int n, cudaStream_t st, int status_mod = 0);
// ... the same in two launches: decode only (malformed / off-curve -> status), then the subgroup test on the decoded points
void launch_decode_g1_unchecked(const uint8_t* in48_dev, void* affine_out_dev, int* status_dev, int n, cudaStream_t st, int status_mod = 0);
void launch_subgroup_g1(const void* affine_dev, int* status_dev, int n, cudaStream_t st, int status_mod = 0);
// blst_p1_from_affine for n points
void launch_affine_to_jac(const void* affine_dev, void* jac_dev, int n, cudaStream_t st);
// Fr::from_bytes (reduce = 0: status[i] = 1 when >= r) / hash_to_bls_field (reduce = 1) -> Montgomery; and back
void launch_fr_from_bytes(const uint8_t* bytes32_dev, int n, int reduce, void* fr_mont_dev, int* status_dev, cudaStream_t st);
// bit-reverse (13 bits per 8192-element extended blob) + big-endian serialisation: evaluations -> cells
void launch_cells_out(const void* ext_dev, uint8_t* cells_dev, size_t total, cudaStream_t st);
void launch_fr_to_bytes(const void* fr_mont_dev, int n, uint8_t* bytes32_dev, cudaStream_t st);

}  // namespace b200
