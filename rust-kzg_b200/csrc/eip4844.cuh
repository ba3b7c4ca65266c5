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
    static constexpr int kLanes = 2;
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
    int fk20_batch(cudaStream_t st) { ensure_fk20(st); return fk_batch_; }
    int launches_last() const { return launches_; }

private:
    void ensure_fk20(cudaStream_t st);
    std::unique_ptr<MsmEngine> fk_msm_;
    int fk_batch_ = 0;
    void *fk_a_ = nullptr, *fk_b_ = nullptr, *fk_pts_ = nullptr;
    int max_batch_;
    int launches_ = 0;
    std::unique_ptr<FFTSettingsDev> fs_;
    struct Lane {
        std::unique_ptr<MsmEngine> msm;
        void* scalars = nullptr;   // max_batch * 4096 canonical scalars (MSM input)
        void* poly = nullptr;      // max_batch * 4096 Montgomery field elements
        void* z = nullptr;         // max_batch Montgomery
        void* y = nullptr;         // max_batch Montgomery
        void* out_jac = nullptr;   // max_batch Jacobian results
    } lanes_[kLanes];
    void* lagrange_jac_ = nullptr;
    void* monomial_jac_ = nullptr;
    void* domain_ = nullptr;    // brp_roots_of_unity[0..4096) of the 8192 table (Montgomery)
    void* cells_a_ = nullptr;   // max_batch * 8192 Fr ping-pong buffers (compute_cells), allocated on first use
    void* cells_b_ = nullptr;
};

// uncompress n 48-byte points into affine Montgomery form; flags[i] = 1 on malformed / off-curve input
void launch_uncompress_g1(const uint8_t* in48_dev, void* affine_out_dev, int* flags_dev, int n, cudaStream_t st);

}  // namespace b200
