// msm.cuh -- interface of the bucket-method (Pippenger) MSM engine over BLS12-381 G1 for sm_100a.
//
// Replaces, behind the same contract, the reference's CPU MSM (kzg/src/msm/msm_impls.rs:114-148 -> tiling_pippenger /
// BgmwTable) and its sppark GPU plug (blst-sppark/cuda/pippenger.cu:23-38).  Results are the same group element,
// hence byte-identical after compression; window size, digit recoding and addition order are free
// (SURVEY.md section 7 "hard parts").
//
// One engine serves both shapes of the path:
//   FIXED    bases known in advance (prepare_msm / trusted setup): the table holds W rows  2^(c*j) * P_i  in affine
//            form, so all W signed digits of a scalar go to ONE bucket set and no per-window Horner tail exists
//            (the idea of the reference's default BGMW table, kzg/src/msm/bgmw.rs:206-304, re-laid for HBM).
//            A batch of B scalar vectors over the same bases (B blobs) is one launch sequence: bucket sets are
//            indexed (vector, bucket).
//   VARIABLE bases arrive with the call (mult_pippenger, g1_lincomb(.., None)): one bucket set per window,
//            then a device Horner pass.
//
// Pipeline (all on one stream, no host synchronisation inside):
//   1 digits+count   signed c-bit digits of every scalar, histogram of (group, bucket) keys        [HBM/atomics]
//   2 scan           exclusive prefix sums -> bucket offsets, task bases                           [tiny]
//   3 scatter        counting-sort the (point index, sign) entries by key                          [HBM/atomics]
//   4 tasks          cut every bucket into tasks of <= L entries, counting-sort tasks by length    [tiny]
//   5 accumulate     one thread per task: XYZZ += affine over its entries  (the dominant kernel)   [integer pipe]
//   6 reduce         per group: sum_b b * B_b.  Windows wider than 16 bits (or MsmConfig::fold): running sums over
//                    segments of 2^k buckets first (k_segment_fold); then digit-marginal tree sums over <= 2^15
//                    (segment) sums and a warp suffix scan per digit                               [latency]
//   7 (VARIABLE)     Horner over the W window sums
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

struct MsmConfig {
    int c;          // window width in bits
    int W;          // number of windows, c*W >= 256
    bool fixed;     // FIXED (table rows) or VARIABLE
    size_t n;       // points per scalar vector
    int max_batch;  // scalar vectors per call (FIXED only; VARIABLE uses 1)
    int L;          // max entries per accumulate task
    int c0 = 0;     // FIXED only: width of window 0 when it differs from c (0 = c); windows j >= 1 start at bit c0 + (j-1)*c.
                    // With c0 = 256 - (W-1)*c the TOP window ends at bit 256 and is full: inputs with zero top bits (blob
                    // elements < 2^248) then still spread its digits over many buckets instead of one or two.
    int fold = -1;  // bucket-index bits folded by k_segment_fold before the marginal reduce; -1: only what the reduce
                    // cannot take (c - 1 - 15 bits for windows wider than 16)
    bool randomize = false;  // FIXED only: multiply every scalar by a per-base pseudo-random rho_i and build the table over
                             // rho_i^-1 * P_i (msm.cu, "scalar randomisation"): digit distributions become independent of the
                             // input.  Only applied when every base is in the prime-order subgroup (checked at construction).
    uint64_t rho_seed = 0;   // 0: the library's fixed seed
    bool affine = false;   // FIXED only: batch-affine accumulation (k_accumulate_affine) for calls that fill the machine; the table
                           // then uses 128-byte slots per point (one aligned line per gather instead of a straddling 96 bytes)
    int affine_k = 10;     // slots (chunks) per thread of the batch-affine kernel: denominators per inversion = 128 * affine_k
    int bases_period = 1;  // FIXED only: the table holds bases_period * n points per row and scalar vector v uses the
                           // bases [(v mod bases_period) * n, +n)  (FK20: 128 rows of 64 points, kzg/src/msm/bgmw.rs:306-380)
};

class MsmEngine {
public:
    // points: n affine points (blst_p1_affine layout), device or host pointer (host_points says which).
    // share_table: another FIXED engine with the same bases and window layout whose table this one reads instead of
    // building its own (the lanes of a settings object: one table in L2 for all of them); it must outlive this engine.
    MsmEngine(const MsmConfig& cfg, const void* points, bool host_points, cudaStream_t stream,
              const MsmEngine* share_table = nullptr);
    ~MsmEngine();
    MsmEngine(const MsmEngine&) = delete;

    // scalars_dev: batch * n scalars of 32 B in device memory; mont = blst_fr Montgomery form (else canonical LE).
    // out_dev: batch Jacobian points (blst_p1 layout) in device memory.  npoints <= n uses the first npoints bases.
    // scalars_host != nullptr: the scalars are still on the host; they are copied into scalars_dev (a device staging
    // buffer of batch * npoints * 32 bytes) in chunks that overlap the digit histogram.
    void run(const void* scalars_dev, size_t npoints, int batch, bool mont, void* out_dev, cudaStream_t stream,
             const void* scalars_host = nullptr);

    // VARIABLE engines only: replace the bases (device pointer, n points).
    void set_points(const void* points_dev, size_t npoints, cudaStream_t stream);

    const MsmConfig& config() const { return cfg_; }
    size_t table_bytes() const { return table_bytes_; }
    // launches of our kernels per run() (for bench.py's gpu_launches)
    int launches_per_run() const { return launches_; }
    const void* table() const { return table_; }
    size_t table_stride() const { return stride_; }
    bool last_run_affine() const { return last_affine_; }
    bool randomized() const { return rho_seed_ != 0; }
    // optional device-side timing of the accumulate kernel (bench.py's roofline): when enabled, every run() brackets
    // k_accumulate with CUDA events on the launching stream; profile_read() sums the completed pairs and resets.
    void set_profiling(bool on) { profiling_ = on; }
    void profile_read(double* accumulate_ms_sum, int* runs);
    // entries (non-zero digits) and accumulate tasks of the last run(); synchronises `stream`
    void last_counts(size_t* entries, size_t* tasks, cudaStream_t stream);
    // work counters of the last run() for the bench's addition count: [entries, tasks, non-empty buckets, keys, fold
    // bits, buckets per group the marginal reduce sees, digit axes, groups]; synchronises `stream`
    void last_stats(uint64_t out[8], cudaStream_t stream);

private:
    MsmConfig cfg_;
    int nb_;              // buckets per group = 2^(c-1)
    size_t groups_max_;   // max_batch (FIXED) or W (VARIABLE)
    size_t keys_max_;     // groups_max * nb
    size_t entries_max_;  // max_batch * n * W
    size_t tasks_max_;
    size_t table_bytes_;
    int launches_ = 0;
    size_t last_nkeys_ = 0;
    int kf_ = 0;          // segment-fold bits actually used
    bool profiling_ = false;
    static constexpr int kProfSlots = 512;
    cudaEvent_t prof_ev_[2 * kProfSlots] = {};
    int prof_count_ = 0;
    static constexpr int kCopyChunks = 4;
    cudaStream_t copy_stream_ = nullptr;
    cudaEvent_t copy_ev_[kCopyChunks] = {};
    cudaEvent_t copy_start_ = nullptr;
    void* table_ = nullptr;      // affine rows
    bool owns_table_ = true;
    uint64_t rho_seed_ = 0;      // != 0: scalar randomisation is active on this table
    uint32_t* counts_ = nullptr;  // [keys+1]
    uint32_t* offsets_ = nullptr;
    uint32_t* cursor_ = nullptr;
    uint32_t* task_base_ = nullptr;
    uint32_t* entries_ = nullptr;
    uint32_t* sorted_tasks_ = nullptr;  // 3 x u32 per task
    uint32_t* size_hist_ = nullptr;     // [L+1] hist, [L+1] base, [L+1] cursor
    uint32_t* scan_tmp_ = nullptr;
    // batch-affine accumulation (cfg.affine): chunk grid of the sorted entry list
    size_t stride_ = 96;                // bytes per table point
    bool ba_tree_ = false;              // one inversion per CTA (shared-memory product tree) instead of one per warp
    int ba_blocks_ = 0;                 // co-resident CTAs of k_accumulate_affine (SMs x occupancy): the grid
    size_t ba_smem_ = 0;
    size_t ba_chunks_max_ = 0;          // blocks x 128 x K
    uint32_t* first_slot_ = nullptr;    // [chunks] partial slot of each chunk's first segment
    uint8_t* acc_buf_ = nullptr;        // [chunks] x 96 B running affine sums (L2 resident)
    bool last_affine_ = false;          // the last run() took the batch-affine path
    void* partials_ = nullptr;  // xyzz per task
    void* chunk_sums_ = nullptr;
    void* fin_scratch_ = nullptr;      // k_group_finish: 17 partial sums per group
    unsigned* fin_cnt_ = nullptr;      // ... and its completion counters (zero between runs)
    void* group_sums_ = nullptr;
    void* seg_t_ = nullptr;            // wide windows (c > 16): per-segment sums T and running sums R, see k_segment_fold
    void* seg_r_ = nullptr;
    uint32_t* seg_ident_ = nullptr;    // identity slot map for the segment arrays
    void* chunk_sums_r_ = nullptr;     // marginals of the R sums
};

// microbenchmark of independent batch-affine pair additions (msm.cu): kernel time in ms
float bench_affine_pairs(uint32_t npoints, size_t npairs, int K, cudaStream_t st);
// helpers shared with other translation units
void launch_g1_sum(const void* jac_dev, void* out_jac_dev, int count, cudaStream_t stream);
void launch_points_to_compressed(const void* jac_dev, uint8_t* out48_dev, int count, cudaStream_t stream, int brp_bits = 0);
void launch_jac_to_affine(const void* jac_dev, void* affine_dev, int count, cudaStream_t stream);

}  // namespace b200
