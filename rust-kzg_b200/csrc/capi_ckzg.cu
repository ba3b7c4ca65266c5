// capi_ckzg.cu -- the c-kzg-4844 C ABI for the commitment / proof path (include/b200_kzg.h, section B2), replacing
// blst/src/eip_4844.rs:160-530 for: load_trusted_setup, load_trusted_setup_file, free_trusted_setup,
// blob_to_kzg_commitment, compute_kzg_proof, compute_blob_kzg_proof, plus batched and device-pointer extensions.
//
// Like the reference, the settings struct the caller holds carries only host arrays; the device context hangs off a
// side registry keyed by the g1_values_lagrange_brp pointer (the reference keys its PrecomputationTableManager the
// same way, kzg/src/eip_4844.rs:105-145) -- but instead of rebuilding an FsKZGSettings on every call
// (blst/src/types/kzg_settings.rs:314-437) the resident context is looked up.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200_kzg.h"
#include "capi_common.cuh"
#include "coalesce.cuh"
#include "eip4844.cuh"
#include "sha256.cpp.inc"
#include "util.cuh"

using namespace b200;

namespace {

constexpr size_t kG1 = 4096, kG2 = 65;

// one staging lane: stream, device buffers for a chunk of up to max_batch blobs, pinned host buffers for the results
struct Stage {
    cudaStream_t stream = nullptr, side = nullptr;
    cudaEvent_t ev_in = nullptr, ev_side = nullptr, ev_dec = nullptr;
    // end of the last enqueue a device-pointer entry point left on a CALLER's stream: whoever uses this lane's
    // workspace next makes its stream wait on it first (the workspace is shared, the streams are not)
    cudaEvent_t ev_busy = nullptr;
    bool busy = false;
    uint8_t *d_blobs = nullptr, *d_z = nullptr, *d_comm = nullptr, *d_out48 = nullptr, *d_y32 = nullptr;
    int *d_status = nullptr, *d_status2 = nullptr;
    uint8_t* h_small = nullptr;  // pinned: [48 * mb out][32 * mb y][int * mb][int * mb][32 * mb z]
    int mb = 0;
    void init(int max_batch) {
        mb = max_batch;
        B200_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        B200_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming));
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_dec, cudaEventDisableTiming));
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_busy, cudaEventDisableTiming));
        d_blobs = dev_alloc<uint8_t>((size_t)mb * kBytesPerBlob);
        d_z = dev_alloc<uint8_t>((size_t)mb * 32);
        d_comm = dev_alloc<uint8_t>((size_t)mb * 48);
        d_out48 = dev_alloc<uint8_t>((size_t)mb * 48);
        d_y32 = dev_alloc<uint8_t>((size_t)mb * 32);
        d_status = dev_alloc<int>(mb);
        d_status2 = dev_alloc<int>(mb);
        B200_CUDA_CHECK(cudaMallocHost((void**)&h_small, (size_t)mb * 120 + 64));
    }
    ~Stage() {
        cudaFree(d_blobs); cudaFree(d_z); cudaFree(d_comm); cudaFree(d_out48); cudaFree(d_y32); cudaFree(d_status); cudaFree(d_status2);
        if (h_small) cudaFreeHost(h_small);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_side) cudaEventDestroy(ev_side);
        if (ev_dec) cudaEventDestroy(ev_dec);
        if (ev_busy) cudaEventDestroy(ev_busy);
        if (stream) cudaStreamDestroy(stream);
        if (side) cudaStreamDestroy(side);
    }
    uint8_t* h_out48() { return h_small; }
    uint8_t* h_y32() { return h_small + 48 * (size_t)mb; }
    int* h_status() { return reinterpret_cast<int*>(h_small + 80 * (size_t)mb); }
    int* h_status2() { return h_status() + mb; }
    uint8_t* h_z() { return h_small + 88 * (size_t)mb; }
};

// Ownership of the device lanes (Stage + MSM engines + workspace each).  A coalesced batch of single-blob calls takes ANY
// free lane; everything else (explicit batches, cells / FK20, verification, recovery) takes ALL of them.  A pending
// take-all blocks new single takers, so it cannot starve.
struct LanePool {
    static constexpr unsigned kFull = (1u << KzgSettingsDev::kLanes) - 1;
    std::mutex m;
    std::condition_variable cv;
    unsigned usable = kFull;      // lanes a single taker may get (B200_KZG_LANES limits them; take-all always takes all)
    unsigned free_mask = kFull;
    int all_waiters = 0;
    int acquire_any() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return (free_mask & usable) != 0 && all_waiters == 0; });
        int lane = __builtin_ctz(free_mask & usable);
        free_mask &= ~(1u << lane);
        return lane;
    }
    void acquire_all() {
        std::unique_lock<std::mutex> lk(m);
        all_waiters++;
        cv.wait(lk, [&] { return free_mask == kFull; });
        all_waiters--;
        free_mask = 0;
    }
    void release(int lane) {
        { std::lock_guard<std::mutex> lk(m); free_mask |= 1u << lane; }
        cv.notify_all();
    }
    void release_all() {
        { std::lock_guard<std::mutex> lk(m); free_mask = kFull; }
        cv.notify_all();
    }
};

// ---- request coalescer (coalesce.cuh): one open batch per single-blob entry point ------------------------------------
enum CoKind { CO_COMMIT = 0, CO_PROOF = 1, CO_BLOB_PROOF = 2, CO_KINDS = 3 };
struct CoBatch : CoBatchBase {
    uint8_t* h_in = nullptr;   // pinned: [mb blobs][mb x 48 commitments][mb x 32 z]
    uint8_t* h_out = nullptr;  // pinned: [mb x 48 out][mb x 32 y][mb int status][mb int status2]
    int mb = 0;
    uint8_t* blob(int i) { return h_in + (size_t)i * kBytesPerBlob; }
    uint8_t* comm(int i) { return h_in + (size_t)mb * kBytesPerBlob + 48 * (size_t)i; }
    uint8_t* z(int i) { return h_in + (size_t)mb * (kBytesPerBlob + 48) + 32 * (size_t)i; }
    uint8_t* out48(int i) { return h_out + 48 * (size_t)i; }
    uint8_t* y32(int i) { return h_out + 48 * (size_t)mb + 32 * (size_t)i; }
    int* status() { return reinterpret_cast<int*>(h_out + 80 * (size_t)mb); }
    int* status2() { return status() + mb; }
    ~CoBatch() {
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
    }
};
typedef CoQueue<CoBatch, CO_KINDS> Coalescer;
// compute_cells_and_kzg_proofs called per blob from several threads (a block's blobs under rayon par_iter): its own queue --
// 128 cells + 128 proofs per request, batches of up to 16 blobs (the size the fused fft_g1 stages still serve)
struct CellsBatch : CoBatchBase {
    uint8_t *h_in = nullptr, *h_cells = nullptr, *h_proofs = nullptr;   // pinned: cap blobs / cap x 128 x 2048 / cap x 128 x 48
    int* h_status = nullptr;                                              // pinned: cap
    int cap = 0;
    uint8_t* blob(int i) { return h_in + (size_t)i * kBytesPerBlob; }
    uint8_t* cells(int i) { return h_cells + (size_t)i * 128 * 2048; }
    uint8_t* proofs(int i) { return h_proofs + (size_t)i * 128 * 48; }
    ~CellsBatch() {
        if (h_in) cudaFreeHost(h_in);
        if (h_cells) cudaFreeHost(h_cells);
        if (h_proofs) cudaFreeHost(h_proofs);
        if (h_status) cudaFreeHost(h_status);
    }
};
typedef CoQueue<CellsBatch, 1> CellsCoalescer;
// verify_blob_kzg_proof (kind 0) / verify_kzg_proof (kind 1) called one at a time from several threads (blob sidecars arriving
// from gossip): concurrent requests are checked as ONE batch -- the random linear combination the reference's own
// verify_blob_kzg_proof_batch uses (kzg/src/eip_4844.rs:380-435) -- and only when that batch does not verify (or an input
// does not decode) does every caller fall back to its own single check, so each still gets exactly its own answer
enum { VK_BLOB = 0, VK_PROOF = 1, VK_KINDS = 2 };
struct VerifyBatch : CoBatchBase {
    uint8_t* h_in = nullptr;   // pinned: [cap blobs][cap x 48 C][cap x 48 proof][cap x 32 z][cap x 32 y]
    int cap = 0;
    uint8_t* blob(int i) { return h_in + (size_t)i * kBytesPerBlob; }
    uint8_t* comm(int i) { return h_in + (size_t)cap * kBytesPerBlob + 48 * (size_t)i; }
    uint8_t* proof(int i) { return comm(cap) + 48 * (size_t)i; }
    uint8_t* z(int i) { return proof(cap) + 32 * (size_t)i; }
    uint8_t* y(int i) { return z(cap) + 32 * (size_t)i; }
    bool all_ok = false;       // the batch verified: every request in it is valid
    bool single_done = false;  // the batch held one request: its exact verdict is in rc / all_ok
    ~VerifyBatch() { if (h_in) cudaFreeHost(h_in); }
};
typedef CoQueue<VerifyBatch, VK_KINDS> VerifyCoalescer;

struct KzgCtx {
    LanePool pool;
    Coalescer co;
    CellsCoalescer co_cells;
    VerifyCoalescer co_verify;
    int vco_cap = 32;          // most single verifications checked as one batch (B200_KZG_VERIFY_COALESCE; 1: none)
    std::atomic<uint64_t> st_verify_batches{0}, st_verify_requests{0}, st_verify_fallbacks{0};
    int cells_grace_us = 200;  // how long a leader lingers for the rest of a burst (B200_KZG_CELLS_GRACE_US; 0: not at all)
    int cells_cap = 16;        // most single-blob cells + proofs requests per launch sequence (B200_KZG_CELLS_COALESCE; 1: none)
    int device = 0;            // CUDA device the context lives on: every entry point switches to it (DeviceScope)
    int co_cap = 0;            // most single-blob requests packed into one launch sequence (<= max_batch)
    // coalescer counters (b200_kzg_coalesce_stats): batches run, requests served, ns spent waiting for a lane, ns on a lane
    std::atomic<uint64_t> st_batches{0}, st_requests{0}, st_wait_ns{0}, st_exec_ns{0}, st_max_batch{0};
    std::atomic<uint64_t> st_cells_batches{0}, st_cells_requests{0};   // the same for coalesced compute_cells_and_kzg_proofs
    std::unique_ptr<KzgSettingsDev> dev;
    int max_batch = 0;
    Stage stage[KzgSettingsDev::kLanes];
    // lane 0 aliases used by the single-lane paths (cells, FK20)
    cudaStream_t stream = nullptr;
    uint8_t* d_blobs = nullptr;
    int* d_status = nullptr;
    int* h_status() { return stage[0].h_status(); }
    uint8_t* d_cells = nullptr;   // max_batch x 128 x 2048, allocated on first use (compute_cells)
    uint8_t* d_proofs = nullptr;  // fk20 batch x 128 x 48, allocated on first use
    // verification staging, grown on demand: device [c48 n][p48 n][z32 n][y32 n][r32][status n ints][result int],
    // pinned host mirror of the same layout
    uint8_t* d_verify = nullptr;
    uint8_t* h_verify = nullptr;
    size_t verify_cap = 0;
    static size_t verify_bytes(size_t n) { return n * (48 + 48 + 32 + 32 + sizeof(int)) + 32 + 64; }
    // general staging for the EIP-7594 entry points: device + pinned host, grown on demand, reused across calls
    uint8_t* d_scratch = nullptr;
    uint8_t* h_scratch = nullptr;
    size_t scratch_cap = 0;
    void ensure_scratch(size_t bytes) {
        if (bytes <= scratch_cap) return;
        cudaFree(d_scratch);
        if (h_scratch) cudaFreeHost(h_scratch);
        d_scratch = nullptr; h_scratch = nullptr; scratch_cap = 0;
        d_scratch = dev_alloc<uint8_t>(bytes);
        B200_CUDA_CHECK(cudaMallocHost((void**)&h_scratch, bytes));
        scratch_cap = bytes;
    }
    void ensure_verify(size_t n) {
        if (n <= verify_cap) return;
        cudaFree(d_verify);
        if (h_verify) cudaFreeHost(h_verify);
        d_verify = nullptr; h_verify = nullptr; verify_cap = 0;
        d_verify = dev_alloc<uint8_t>(verify_bytes(n));
        B200_CUDA_CHECK(cudaMallocHost((void**)&h_verify, verify_bytes(n)));
        verify_cap = n;
    }
    ~KzgCtx() {
        cudaFree(d_cells); cudaFree(d_proofs); cudaFree(d_verify); cudaFree(d_scratch);
        if (h_verify) cudaFreeHost(h_verify);
        if (h_scratch) cudaFreeHost(h_scratch);
        dev.reset();
    }
};

// Before a lane's workspace is used on `user` (nullptr: the lane's own streams), order that stream after whatever a
// device-pointer entry point last enqueued against the same workspace on some caller's stream.
void lane_enter(Stage& g, cudaStream_t user) {
    if (!g.busy) return;
    if (user) {
        B200_CUDA_CHECK(cudaStreamWaitEvent(user, g.ev_busy, 0));
    } else {
        B200_CUDA_CHECK(cudaStreamWaitEvent(g.stream, g.ev_busy, 0));
        B200_CUDA_CHECK(cudaStreamWaitEvent(g.side, g.ev_busy, 0));
    }
}
// ... and after such an enqueue, remember where it ends
void lane_leave_async(Stage& g, cudaStream_t user) {
    B200_CUDA_CHECK(cudaEventRecord(g.ev_busy, user));
    g.busy = true;
}
struct AllLanes {   // exclusive use of the context: all lanes, on the context's device
    KzgCtx& c;
    DeviceScope ds;
    explicit AllLanes(KzgCtx& ctx) : c(ctx), ds(ctx.device) {
        c.pool.acquire_all();
        try {
            for (Stage& g : c.stage) lane_enter(g, nullptr);
        } catch (...) {
            c.pool.release_all();
            throw;
        }
    }
    ~AllLanes() { c.pool.release_all(); }
    AllLanes(const AllLanes&) = delete;
};
struct OneLane {    // one lane (any), for a coalesced batch or a device-pointer call
    KzgCtx& c;
    DeviceScope ds;
    int lane;
    explicit OneLane(KzgCtx& ctx, cudaStream_t user = nullptr) : c(ctx), ds(ctx.device), lane(ctx.pool.acquire_any()) {
        try {
            lane_enter(c.stage[lane], user);
        } catch (...) {
            c.pool.release(lane);
            throw;
        }
    }
    ~OneLane() { c.pool.release(lane); }
    OneLane(const OneLane&) = delete;
};

// Run `n` items in chunks of at most `cap`, alternating between the lanes: chunk k+1 is enqueued (copies, host
// hashing, kernels) while chunk k is still executing, and a lane is drained only when it is needed again.
// launch(lane, off, m) enqueues a chunk; finish(lane, off, m) runs after that lane's stream has been synchronised.
template <class Launch, class Finish>
C_KZG_RET run_chunks(KzgCtx& ctx, size_t n, size_t cap, Launch&& launch, Finish&& finish) {
    struct Pending { bool live = false; size_t off = 0; int m = 0; } pend[KzgSettingsDev::kLanes];
    C_KZG_RET rc = C_KZG_OK;
    auto drain = [&](int lane) {
        if (!pend[lane].live) return;
        B200_CUDA_CHECK(cudaStreamSynchronize(ctx.stage[lane].stream));
        pend[lane].live = false;
        if (rc == C_KZG_OK) rc = finish(lane, pend[lane].off, pend[lane].m);
    };
    size_t k = 0;
    for (size_t off = 0; off < n && rc == C_KZG_OK; off += cap, k++) {
        int lane = (int)(k % KzgSettingsDev::kLanes);
        drain(lane);
        if (rc != C_KZG_OK) break;
        int m = (int)std::min(cap, n - off);
        launch(lane, off, m);
        pend[lane] = {true, off, m};
    }
    // drain in submission order
    for (int d = 0; d < KzgSettingsDev::kLanes; d++) drain((int)((k + d) % KzgSettingsDev::kLanes));
    return rc;
}

std::mutex g_reg_mu;
std::map<const void*, std::shared_ptr<KzgCtx>> g_registry;

std::shared_ptr<KzgCtx> find_ctx(const KZGSettings* s) {
    if (!s || !s->g1_values_lagrange_brp) return nullptr;
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registry.find(s->g1_values_lagrange_brp);
    return it == g_registry.end() ? nullptr : it->second;
}

void zero_settings(KZGSettings* out) { memset(out, 0, sizeof(*out)); }

C_KZG_RET load_impl(KZGSettings* out, const uint8_t* g1_monomial, size_t n_mono, const uint8_t* g1_lagrange, size_t n_lag,
                    const uint8_t* g2_monomial, size_t n_g2) {
    zero_settings(out);
    // load_trusted_setup_rust's length checks (kzg/src/eip_4844.rs:1037-1049)
    // (chunks(48) + from_bytes: a trailing partial point fails to decode, so the byte counts must be exact)
    if (n_mono != kG1 * 48 || n_lag != kG1 * 48 || n_g2 != kG2 * 96) return C_KZG_BADARGS;
    if (!g1_monomial || !g1_lagrange || !g2_monomial) return C_KZG_BADARGS;
    std::shared_ptr<KzgCtx> ctx(new KzgCtx());
    try {
        require_device();
        B200_CUDA_CHECK(cudaGetDevice(&ctx->device));
        ctx->max_batch = env_int("B200_KZG_MAX_BATCH", 64);
        if (ctx->max_batch < 1) ctx->max_batch = 1;
        ctx->co_cap = std::max(1, std::min(ctx->max_batch, env_int("B200_KZG_COALESCE", ctx->max_batch)));
        ctx->cells_cap = std::max(1, std::min(std::min(ctx->max_batch, 64), env_int("B200_KZG_CELLS_COALESCE", 16)));
        ctx->co_cells.max_batches = 2;
        ctx->vco_cap = std::max(1, std::min(64, env_int("B200_KZG_VERIFY_COALESCE", 32)));
        ctx->co_verify.max_batches = 3;
        ctx->cells_grace_us = std::max(0, env_int("B200_KZG_CELLS_GRACE_US", 200));
        ctx->co.max_batches = KzgSettingsDev::kLanes + 2;   // one per lane in flight + the ones filling
        {
            const int lanes = std::max(1, std::min((int)KzgSettingsDev::kLanes, env_int("B200_KZG_LANES", KzgSettingsDev::kLanes)));
            ctx->pool.usable = (1u << lanes) - 1;
        }
        const int mb = ctx->max_batch;
        for (Stage& st : ctx->stage) st.init(mb);
        ctx->stream = ctx->stage[0].stream;
        ctx->d_blobs = ctx->stage[0].d_blobs;
        ctx->d_status = ctx->stage[0].d_status;
        ctx->dev.reset(new KzgSettingsDev(g1_monomial, g1_lagrange, mb, ctx->stream));
        ctx->dev->load_g2(g2_monomial, (int)kG2, ctx->stream);  // G2::from_bytes of all 65 points (eip_4844.rs:1050-1053)
        // is_trusted_setup_in_lagrange_form (eip_4844.rs:1005-1020, 1064-1068): a monomial-form array in the Lagrange slot
        // satisfies e(L[1], G2) == e(L[0], [s]G2) and is rejected.  L[1] sits at bit-reversed position 2048.
        {
            int* d_res = dev_alloc<int>(1);
            const uint8_t* lag = (const uint8_t*)ctx->dev->g1_lagrange_brp_jac_dev();
            ctx->dev->pairings_verify(lag + (size_t)2048 * 144, 0, lag, 1, d_res, ctx->stream);
            int res = 0;
            cudaError_t e = cudaMemcpyAsync(&res, d_res, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_res);
            B200_CUDA_CHECK(e);
            if (res) throw CudaError(1, "Trusted setup is not in Lagrange form");
        }
    } catch (const CudaError& e) {
        if (e.code != 1) fprintf(stderr, "b200kzg: load_trusted_setup failed: %s\n", e.what());
        return e.code == 1 ? C_KZG_BADARGS : C_KZG_ERROR;
    } catch (const std::exception& e) {
        fprintf(stderr, "b200kzg: load_trusted_setup failed: %s\n", e.what());
        return C_KZG_ERROR;
    }
    // host arrays of the C struct (kzg_settings_to_c, blst/src/eip_4844.rs:40-144)
    FFTSettingsDev& fs = ctx->dev->fft();
    const size_t w = fs.max_width();  // 8192
    out->roots_of_unity = (blst_fr*)malloc((w + 1) * sizeof(blst_fr));
    out->brp_roots_of_unity = (blst_fr*)malloc(w * sizeof(blst_fr));
    out->reverse_roots_of_unity = (blst_fr*)malloc((w + 1) * sizeof(blst_fr));
    out->g1_values_monomial = (blst_p1*)malloc(kG1 * sizeof(blst_p1));
    out->g1_values_lagrange_brp = (blst_p1*)malloc(kG1 * sizeof(blst_p1));
    out->g2_values_monomial = (blst_p2*)calloc(kG2, sizeof(blst_p2));
    if (!out->roots_of_unity || !out->brp_roots_of_unity || !out->reverse_roots_of_unity || !out->g1_values_monomial ||
        !out->g1_values_lagrange_brp || !out->g2_values_monomial) {
        free_trusted_setup(out);
        return C_KZG_MALLOC;
    }
    bool ok = cudaMemcpy(out->g2_values_monomial, ctx->dev->g2_monomial_jac_dev(), kG2 * sizeof(blst_p2), cudaMemcpyDeviceToHost) == cudaSuccess && cudaMemcpy(out->roots_of_unity, fs.roots_dev(), (w + 1) * 32, cudaMemcpyDeviceToHost) == cudaSuccess &&
              cudaMemcpy(out->brp_roots_of_unity, fs.brp_roots_dev(), w * 32, cudaMemcpyDeviceToHost) == cudaSuccess &&
              cudaMemcpy(out->g1_values_monomial, ctx->dev->g1_monomial_jac_dev(), kG1 * 144, cudaMemcpyDeviceToHost) == cudaSuccess &&
              cudaMemcpy(out->g1_values_lagrange_brp, ctx->dev->g1_lagrange_brp_jac_dev(), kG1 * 144, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!ok) {
        free_trusted_setup(out);
        return C_KZG_ERROR;
    }
    for (size_t i = 0; i <= w; i++) out->reverse_roots_of_unity[i] = out->roots_of_unity[w - i];
    // x_ext_fft_columns[128][64], one heap row per column like kzg_settings_to_c (blst/src/eip_4844.rs:104-131): the
    // reference's TryFrom<&CKZGSettings> (blst/src/types/kzg_settings.rs:398-417) walks all 128 row pointers
    {
        const size_t rows = 128, cols = 64;
        out->x_ext_fft_columns = (blst_p1**)calloc(rows, sizeof(blst_p1*));
        bool got = out->x_ext_fft_columns != nullptr;
        for (size_t r = 0; got && r < rows; r++) got = (out->x_ext_fft_columns[r] = (blst_p1*)malloc(cols * sizeof(blst_p1))) != nullptr;
        if (!got) {
            free_trusted_setup(out);
            return C_KZG_MALLOC;
        }
        uint8_t* d_cols = nullptr;
        bool copied = false;
        try {
            d_cols = dev_alloc<uint8_t>(rows * cols * 144);
            ctx->dev->x_ext_fft_columns(d_cols, ctx->stream);
            copied = true;
            for (size_t r = 0; copied && r < rows; r++)
                copied = cudaMemcpy(out->x_ext_fft_columns[r], d_cols + r * cols * 144, cols * 144, cudaMemcpyDeviceToHost) == cudaSuccess;
        } catch (const std::exception& e) {
            fprintf(stderr, "b200kzg: load_trusted_setup failed: %s\n", e.what());
            copied = false;
        }
        cudaFree(d_cols);
        if (!copied) {
            free_trusted_setup(out);
            return C_KZG_ERROR;
        }
    }
    out->tables = nullptr;   // the reference's CPU precomputation (bgmw / wbits): this backend's table lives in HBM
    out->wbits = 0;
    out->scratch_size = 0;
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_registry[out->g1_values_lagrange_brp] = ctx;
    return C_KZG_OK;
}

int hexval(int c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }

// load_trusted_setup_string (kzg/src/eip_4844.rs:151-228): "4096 65" then hex bytes of the G1 Lagrange points,
// the G2 monomial points and the G1 monomial points, whitespace separated.
bool parse_setup_text(const char* text, size_t len, std::vector<uint8_t>& mono, std::vector<uint8_t>& lag, std::vector<uint8_t>& g2) {
    const char *p = text, *end = text + len;
    size_t counts[2];
    for (int k = 0; k < 2; k++) {
        while (p < end && isspace((unsigned char)*p)) p++;
        if (p >= end || !isdigit((unsigned char)*p)) return false;
        size_t v = 0;
        while (p < end && isdigit((unsigned char)*p)) { v = v * 10 + (size_t)(*p - '0'); p++; }
        if (p >= end) return false;
        counts[k] = v;
    }
    if (counts[0] != kG1 || counts[1] != kG2) return false;
    lag.resize(kG1 * 48); g2.resize(kG2 * 96); mono.resize(kG1 * 48);
    std::vector<uint8_t>* parts[3] = {&lag, &g2, &mono};
    for (auto* part : parts) {
        for (size_t i = 0; i < part->size(); i++) {
            while (p < end && isspace((unsigned char)*p)) p++;
            if (p >= end) return false;
            int hi = hexval(*p), lo = p + 1 < end ? hexval(p[1]) : -1;
            if (hi < 0) return false;
            if (lo >= 0) { (*part)[i] = (uint8_t)(hi << 4 | lo); p += 2; } else { (*part)[i] = (uint8_t)hi; p += 1; }
        }
    }
    return true;
}

// Fiat-Shamir challenge bytes of compute_challenge_rust (kzg/src/eip_4844.rs:920-945): the canonical blob bytes and
// the canonical commitment bytes are exactly the caller's bytes once both have passed validation.
void challenge_hash(uint8_t out[32], const uint8_t* blob, const uint8_t* commitment) {
    sha256::Ctx c;
    uint8_t head[32] = {'F', 'S', 'B', 'L', 'O', 'B', 'V', 'E', 'R', 'I', 'F', 'Y', '_', 'V', '1', '_'};
    head[30] = (kFieldElementsPerBlob >> 8) & 0xff;
    head[31] = kFieldElementsPerBlob & 0xff;
    c.update(head, 32);
    c.update(blob, kBytesPerBlob);
    c.update(commitment, 48);
    c.finish(out);
}
// Host worker pool for the Fiat-Shamir hashes of a batch (SHA-256 over 128 KiB per blob, ~77 us each with SHA-NI): the hashes
// gate the quotient kernel and the MSM, so they are spread over the host cores; the workers are created once per process and
// woken per batch (spawning 16 threads per call cost about as much as hashing four blobs on each of them).
class HashPool {
public:
    // begin(): job(i) for i in [0, n) starts on up to width - 1 workers; finish(): the calling thread joins in and returns when
    // all are done.  Whatever the caller does in between (a blocking copy of pageable blobs) overlaps the hashing.
    // One batch at a time: concurrent callers queue in begin().  `job` must stay alive until finish() returns.
    void begin(size_t n, size_t width, const std::function<void(size_t)>& job) {
        call_mu_.lock();
        width = std::max<size_t>(1, std::min(width, n));
        {
            std::lock_guard<std::mutex> lk(mu_);
            while (workers_.size() + 1 < width) workers_.emplace_back([this] { loop(); });
            job_ = &job; n_ = n; next_.store(0); live_ = width - 1; pending_ = width - 1; gen_++;
        }
        if (width > 1) cv_.notify_all();
    }
    void finish() {
        const std::function<void(size_t)>& job = *job_;
        const size_t n = n_;
        for (size_t i; (i = next_.fetch_add(1)) < n;) job(i);
        {
            std::unique_lock<std::mutex> lk(mu_);
            done_cv_.wait(lk, [&] { return pending_ == 0; });
            job_ = nullptr;
        }
        call_mu_.unlock();
    }
    void run(size_t n, size_t width, const std::function<void(size_t)>& job) {
        begin(n, width, job);
        finish();
    }
    ~HashPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }

private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(size_t)>* job;
            size_t n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || (gen_ != seen && live_ > 0); });
                if (stop_) return;
                seen = gen_;
                live_--;                                       // this worker takes part in the current batch
                job = job_; n = n_;
            }
            for (size_t i; (i = next_.fetch_add(1)) < n;) (*job)(i);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_cv_.notify_one();
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> workers_;
    const std::function<void(size_t)>* job_ = nullptr;
    size_t n_ = 0, live_ = 0, pending_ = 0;
    std::atomic<size_t> next_{0};
    uint64_t gen_ = 0;
    bool stop_ = false;
};
HashPool& hash_pool() {
    static HashPool pool;
    return pool;
}
size_t hash_width(size_t n) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::min<size_t>(n, hw ? hw : 1);
    return std::min<size_t>(nt, (size_t)env_int("B200_SHA_THREADS", 16));
}
void challenge_hash_many(uint8_t* out32, const uint8_t* blobs, const uint8_t* commitments, size_t n) {
    hash_pool().run(n, hash_width(n), [&](size_t i) { challenge_hash(out32 + 32 * i, blobs + i * kBytesPerBlob, commitments + 48 * i); });
}

template <class F>
C_KZG_RET ckzg_guard(F&& f) {
    try {
        return f();
    } catch (const CudaError& e) {
        cudaGetLastError();
        if (e.code == 1) return C_KZG_BADARGS;
        fprintf(stderr, "b200kzg: %s\n", e.what());
        return C_KZG_ERROR;
    } catch (const std::exception& e) {
        fprintf(stderr, "b200kzg: %s\n", e.what());
        return C_KZG_ERROR;
    }
}

bool any_set(const int* st, int n) {
    for (int i = 0; i < n; i++)
        if (st[i]) return true;
    return false;
}

// ---- enqueue one chunk of m <= max_batch items on a lane (host pointers in, pinned host pointers out) ---------------
// blob_to_kzg_commitment (blst/src/eip_4844.rs:163-175)
void enqueue_commit(KzgCtx& ctx, int lane, const uint8_t* blobs, int m, uint8_t* h_out48, int* h_st) {
    Stage& g = ctx.stage[lane];
    B200_CUDA_CHECK(cudaMemcpyAsync(g.d_blobs, blobs, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, g.stream));
    B200_CUDA_CHECK(cudaMemsetAsync(g.d_status, 0, m * sizeof(int), g.stream));
    ctx.dev->blob_to_commitments(g.d_blobs, m, g.d_out48, g.d_status, g.stream, lane);
    B200_CUDA_CHECK(cudaMemcpyAsync(h_out48, g.d_out48, (size_t)m * 48, cudaMemcpyDeviceToHost, g.stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_st, g.d_status, m * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
}
// compute_kzg_proof (blst/src/eip_4844.rs:476-496)
void enqueue_proof(KzgCtx& ctx, int lane, const uint8_t* blobs, const uint8_t* zs, int m, uint8_t* h_out48, uint8_t* h_y32, int* h_st) {
    Stage& g = ctx.stage[lane];
    B200_CUDA_CHECK(cudaMemcpyAsync(g.d_blobs, blobs, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, g.stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(g.d_z, zs, (size_t)m * 32, cudaMemcpyHostToDevice, g.stream));
    B200_CUDA_CHECK(cudaMemsetAsync(g.d_status, 0, m * sizeof(int), g.stream));
    ctx.dev->compute_proofs(g.d_blobs, g.d_z, 0, m, g.d_out48, g.d_y32, g.d_status, g.stream, lane);
    B200_CUDA_CHECK(cudaMemcpyAsync(h_out48, g.d_out48, (size_t)m * 48, cudaMemcpyDeviceToHost, g.stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_y32, g.d_y32, (size_t)m * 32, cudaMemcpyDeviceToHost, g.stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_st, g.d_status, m * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
}
// compute_blob_kzg_proof (blst/src/eip_4844.rs:274-291).  z_hashed: the m Fiat-Shamir hashes when the callers have
// already computed them (pinned), else nullptr: hash here, on the host, while the blobs cross PCIe.
void enqueue_blob_proof(KzgCtx& ctx, int lane, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* z_hashed, int m,
                        uint8_t* h_out48, int* h_st, int* h_st2) {
    Stage& g = ctx.stage[lane];
    // commitments first: their validation (decode + subgroup test) runs on the side stream under everything else
    B200_CUDA_CHECK(cudaMemcpyAsync(g.d_comm, commitments, (size_t)m * 48, cudaMemcpyHostToDevice, g.stream));
    B200_CUDA_CHECK(cudaMemsetAsync(g.d_status, 0, m * sizeof(int), g.stream));
    B200_CUDA_CHECK(cudaMemsetAsync(g.d_status2, 0, m * sizeof(int), g.stream));
    B200_CUDA_CHECK(cudaEventRecord(g.ev_in, g.stream));
    B200_CUDA_CHECK(cudaStreamWaitEvent(g.side, g.ev_in, 0));
    ctx.dev->validate_commitments(g.d_comm, m, g.d_status2, g.side);
    B200_CUDA_CHECK(cudaEventRecord(g.ev_side, g.side));
    if (!z_hashed) {
        // the hash chain runs on the host workers while the blobs cross PCIe (and the other lanes compute); the copy of
        // PAGEABLE blobs blocks this thread while the driver stages them, so the workers are started first
        uint8_t* hz = g.h_z();
        const std::function<void(size_t)> job = [=](size_t i) { challenge_hash(hz + 32 * i, blobs + i * kBytesPerBlob, commitments + 48 * i); };
        hash_pool().begin(m, hash_width(m), job);
        cudaError_t e = cudaMemcpyAsync(g.d_blobs, blobs, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, g.stream);
        hash_pool().finish();
        B200_CUDA_CHECK(e);
        z_hashed = hz;
    } else {
        B200_CUDA_CHECK(cudaMemcpyAsync(g.d_blobs, blobs, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, g.stream));
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(g.d_z, z_hashed, (size_t)m * 32, cudaMemcpyHostToDevice, g.stream));
    ctx.dev->compute_proofs(g.d_blobs, g.d_z, 1, m, g.d_out48, nullptr, g.d_status, g.stream, lane);
    B200_CUDA_CHECK(cudaStreamWaitEvent(g.stream, g.ev_side, 0));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_out48, g.d_out48, (size_t)m * 48, cudaMemcpyDeviceToHost, g.stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_st, g.d_status, m * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_st2, g.d_status2, m * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
}

// ---- the coalesced single-item call --------------------------------------------------------------------------------
// One request of `kind`: blob (+ z32 for CO_PROOF, + commitment48 for CO_BLOB_PROOF) -> out48 (+ y32 for CO_PROOF).
C_KZG_RET coalesced_call(KzgCtx& ctx, int kind, const uint8_t* blob, const uint8_t* arg, uint8_t* out48, uint8_t* y32) {
    Coalescer& co = ctx.co;
    Coalescer::Claim cl = co.claim(kind, ctx.co_cap, [&] {
        std::unique_ptr<CoBatch> nb(new CoBatch());
        nb->mb = ctx.max_batch;
        DeviceScope ds(ctx.device);
        B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_in, (size_t)nb->mb * (kBytesPerBlob + 48 + 32)));
        B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_out, (size_t)nb->mb * (48 + 32 + 2 * sizeof(int))));
        return nb;
    });
    CoBatch* B = cl.b;
    const int idx = cl.idx;
    // every caller stages its own inputs, in parallel with the others (and with whatever runs on the device)
    memcpy(B->blob(idx), blob, kBytesPerBlob);
    if (kind == CO_PROOF) memcpy(B->z(idx), arg, 32);
    if (kind == CO_BLOB_PROOF) {
        memcpy(B->comm(idx), arg, 48);
        challenge_hash(B->z(idx), blob, arg);
    }
    Coalescer::staged(B);
    if (cl.leader) {
        int rc = C_KZG_OK;
        try {
            const auto t0 = std::chrono::steady_clock::now();
            OneLane ln(ctx);   // blocks while every lane is busy: meanwhile the batch keeps filling
            const auto t1 = std::chrono::steady_clock::now();
            const int n = co.close(B);
            Stage& g = ctx.stage[ln.lane];
            if (kind == CO_COMMIT) enqueue_commit(ctx, ln.lane, B->h_in, n, B->out48(0), B->status());
            else if (kind == CO_PROOF) enqueue_proof(ctx, ln.lane, B->h_in, B->z(0), n, B->out48(0), B->y32(0), B->status());
            else enqueue_blob_proof(ctx, ln.lane, B->h_in, B->comm(0), B->z(0), n, B->out48(0), B->status(), B->status2());
            B200_CUDA_CHECK(cudaStreamSynchronize(g.stream));
            const auto t2 = std::chrono::steady_clock::now();
            ctx.st_batches++;
            ctx.st_requests += (uint64_t)n;
            ctx.st_wait_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
            ctx.st_exec_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t2 - t1).count();
            uint64_t mx = ctx.st_max_batch.load();
            while ((uint64_t)n > mx && !ctx.st_max_batch.compare_exchange_weak(mx, (uint64_t)n)) {}
        } catch (const std::exception& e) {
            cudaGetLastError();
            fprintf(stderr, "b200kzg: %s\n", e.what());
            rc = C_KZG_ERROR;
        }
        co.publish(B, rc);
    } else {
        co.wait(B);
    }
    C_KZG_RET rc = (C_KZG_RET)B->rc;
    if (rc == C_KZG_OK) {
        if (B->status()[idx] || (kind == CO_BLOB_PROOF && B->status2()[idx])) {
            rc = C_KZG_BADARGS;
        } else {
            memcpy(out48, B->out48(idx), 48);
            if (kind == CO_PROOF) memcpy(y32, B->y32(idx), 32);
        }
    }
    co.consume(B);
    return rc;
}

// One compute_cells_and_kzg_proofs request (proofs wanted; cells optional): concurrent callers share one pass over up to
// cells_cap blobs.  The leader takes every lane (the FK20 workspace is the settings object's, not a lane's); while a batch
// runs the next one fills.  Each caller copies its own 256 KiB of cells and 6 KiB of proofs out of the pinned staging.
C_KZG_RET coalesced_cells_call(KzgCtx& ctx, const uint8_t* blob, uint8_t* cells_out, uint8_t* proofs_out) {
    CellsCoalescer& co = ctx.co_cells;
    const int cap = ctx.cells_cap;
    CellsCoalescer::Claim cl = co.claim(0, cap, [&] {
        std::unique_ptr<CellsBatch> nb(new CellsBatch());
        nb->cap = cap;
        DeviceScope ds(ctx.device);
        B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_in, (size_t)cap * kBytesPerBlob));
        B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_cells, (size_t)cap * 128 * 2048));
        B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_proofs, (size_t)cap * 128 * 48));
        B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_status, (size_t)cap * sizeof(int)));
        return nb;
    });
    CellsBatch* B = cl.b;
    const int idx = cl.idx;
    memcpy(B->blob(idx), blob, kBytesPerBlob);
    CellsCoalescer::staged(B);
    if (cl.leader) {
        int rc = C_KZG_OK;
        try {
            DeviceScope ds(ctx.device);
            const auto tl0 = std::chrono::steady_clock::now();
            AllLanes lk(ctx);   // blocks while the previous batch (or any other user of the lanes) runs: meanwhile this one fills
            const auto tl1 = std::chrono::steady_clock::now();
            // A burst of callers (a block's blobs from a parallel iterator) arrives within microseconds of each other; the
            // pass takes ~5 ms whatever its size, so the leader lingers while claims keep coming in (at most 200 us, and
            // no longer than 60 us after the last one) instead of leaving the rest of the burst to the next pass.
            if (ctx.cells_grace_us > 0) {
                const auto t0 = std::chrono::steady_clock::now();
                auto last_change = t0;
                int seen = co.claimed_so_far(B);
                while (seen < cap) {
                    std::this_thread::yield();                  // (a 15 us sleep oversleeps by the timer slack: ~60 us each)
                    const auto now = std::chrono::steady_clock::now();
                    const int c = co.claimed_so_far(B);
                    if (c != seen) { seen = c; last_change = now; }
                    if (now - last_change > std::chrono::microseconds(60) || now - t0 > std::chrono::microseconds(ctx.cells_grace_us)) break;
                }
            }
            const auto tl2 = std::chrono::steady_clock::now();
            const int n = co.close(B);
            const auto tl3 = std::chrono::steady_clock::now();
            Stage& g = ctx.stage[0];
            ctx.dev->fk20_batch(g.stream);
            if (!ctx.d_cells) ctx.d_cells = dev_alloc<uint8_t>((size_t)ctx.max_batch * 128 * 2048);
            if (!ctx.d_proofs) ctx.d_proofs = dev_alloc<uint8_t>((size_t)ctx.dev->fk20_batch(g.stream) * 128 * 48);
            B200_CUDA_CHECK(cudaMemcpyAsync(g.d_blobs, B->h_in, (size_t)n * kBytesPerBlob, cudaMemcpyHostToDevice, g.stream));
            B200_CUDA_CHECK(cudaMemsetAsync(g.d_status, 0, n * sizeof(int), g.stream));
            ctx.dev->compute_cells_and_proofs(g.d_blobs, n, ctx.d_cells, ctx.d_proofs, g.d_status, g.stream, g.ev_side);
            B200_CUDA_CHECK(cudaMemcpyAsync(B->h_proofs, ctx.d_proofs, (size_t)n * 128 * 48, cudaMemcpyDeviceToHost, g.stream));
            B200_CUDA_CHECK(cudaStreamWaitEvent(g.side, g.ev_side, 0));          // status + cells leave under the FK20 kernels
            B200_CUDA_CHECK(cudaMemcpyAsync(B->h_status, g.d_status, n * sizeof(int), cudaMemcpyDeviceToHost, g.side));
            B200_CUDA_CHECK(cudaMemcpyAsync(B->h_cells, ctx.d_cells, (size_t)n * 128 * 2048, cudaMemcpyDeviceToHost, g.side));
            B200_CUDA_CHECK(cudaStreamSynchronize(g.side));
            B200_CUDA_CHECK(cudaStreamSynchronize(g.stream));
            ctx.st_cells_batches++;
            ctx.st_cells_requests += (uint64_t)n;
            if (getenv("B200_KZG_CELLS_TRACE")) {
                const auto tl4 = std::chrono::steady_clock::now();
                auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
                    return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count();
                };
                fprintf(stderr, "cells batch n=%d lanes %ld us grace %ld us close %ld us exec %ld us\n", n, us(tl0, tl1), us(tl1, tl2), us(tl2, tl3),
                        us(tl3, tl4));
            }
        } catch (const std::exception& e) {
            cudaGetLastError();
            fprintf(stderr, "b200kzg: %s\n", e.what());
            rc = C_KZG_ERROR;
        }
        co.publish(B, rc);
    } else {
        co.wait(B);
    }
    C_KZG_RET rc = (C_KZG_RET)B->rc;
    if (rc == C_KZG_OK) {
        if (B->h_status[idx]) {
            rc = C_KZG_BADARGS;                                                    // this caller's blob is invalid: it fails alone
        } else {
            if (cells_out) memcpy(cells_out, B->cells(idx), (size_t)128 * 2048);
            if (proofs_out) memcpy(proofs_out, B->proofs(idx), (size_t)128 * 48);
        }
    }
    co.consume(B);
    return rc;
}

}  // namespace

extern "C" {

C_KZG_RET load_trusted_setup(KZGSettings* out, const uint8_t* g1_monomial_bytes, uint64_t num_g1_monomial_bytes,
                             const uint8_t* g1_lagrange_bytes, uint64_t num_g1_lagrange_bytes,
                             const uint8_t* g2_monomial_bytes, uint64_t num_g2_monomial_bytes, uint64_t precompute) {
    (void)precompute;
    if (!out) return C_KZG_BADARGS;
    return load_impl(out, g1_monomial_bytes, num_g1_monomial_bytes, g1_lagrange_bytes, num_g1_lagrange_bytes,
                     g2_monomial_bytes, num_g2_monomial_bytes);
}

C_KZG_RET load_trusted_setup_file(KZGSettings* out, FILE* in) {
    if (!out) return C_KZG_BADARGS;
    zero_settings(out);
    if (!in) return C_KZG_BADARGS;
    // the whole stream, whatever its size (the reference reads the file to a String first, blst/src/eip_4844.rs:232-240)
    std::vector<char> buf(1 << 20);
    size_t len = 0;
    for (;;) {
        len += fread(buf.data() + len, 1, buf.size() - len, in);
        if (len < buf.size()) break;
        if (buf.size() >= ((size_t)1 << 28)) return C_KZG_BADARGS;
        buf.resize(buf.size() * 2);
    }
    std::vector<uint8_t> mono, lag, g2;
    if (!parse_setup_text(buf.data(), len, mono, lag, g2)) return C_KZG_BADARGS;
    return load_impl(out, mono.data(), mono.size(), lag.data(), lag.size(), g2.data(), g2.size());
}

void free_trusted_setup(KZGSettings* s) {
    if (!s) return;
    if (s->g1_values_lagrange_brp) {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        g_registry.erase(s->g1_values_lagrange_brp);
    }
    free(s->roots_of_unity); free(s->brp_roots_of_unity); free(s->reverse_roots_of_unity);
    free(s->g1_values_monomial); free(s->g1_values_lagrange_brp); free(s->g2_values_monomial);
    if (s->x_ext_fft_columns) {
        for (size_t r = 0; r < 128; r++) free(s->x_ext_fft_columns[r]);
        free(s->x_ext_fft_columns);
    }
    zero_settings(s);
}

// ---- batched host-pointer entry points ---------------------------------------------------------------------------
C_KZG_RET b200_blob_to_kzg_commitment_batch(KZGCommitment* out, const Blob* blobs, size_t n, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !out || !blobs) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        return run_chunks(*ctx, n, ctx->max_batch,
            [&](int lane, size_t off, int m) {
                Stage& g = ctx->stage[lane];
                enqueue_commit(*ctx, lane, (const uint8_t*)(blobs + off), m, g.h_out48(), g.h_status());
            },
            [&](int lane, size_t off, int m) -> C_KZG_RET {
                Stage& g = ctx->stage[lane];
                if (any_set(g.h_status(), m)) return C_KZG_BADARGS;
                memcpy(out + off, g.h_out48(), (size_t)m * 48);
                return C_KZG_OK;
            });
    });
}

C_KZG_RET b200_compute_kzg_proof_batch(KZGProof* proofs, Bytes32* ys, const Blob* blobs, const Bytes32* zs, size_t n,
                                       const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !proofs || !ys || !blobs || !zs) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        return run_chunks(*ctx, n, ctx->max_batch,
            [&](int lane, size_t off, int m) {
                Stage& g = ctx->stage[lane];
                enqueue_proof(*ctx, lane, (const uint8_t*)(blobs + off), (const uint8_t*)(zs + off), m, g.h_out48(), g.h_y32(), g.h_status());
            },
            [&](int lane, size_t off, int m) -> C_KZG_RET {
                Stage& g = ctx->stage[lane];
                if (any_set(g.h_status(), m)) return C_KZG_BADARGS;
                memcpy(proofs + off, g.h_out48(), (size_t)m * 48);
                memcpy(ys + off, g.h_y32(), (size_t)m * 32);
                return C_KZG_OK;
            });
    });
}

C_KZG_RET b200_compute_blob_kzg_proof_batch(KZGProof* out, const Blob* blobs, const Bytes48* commitments, size_t n,
                                            const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !out || !blobs || !commitments) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        return run_chunks(*ctx, n, ctx->max_batch,
            [&](int lane, size_t off, int m) {
                Stage& g = ctx->stage[lane];
                enqueue_blob_proof(*ctx, lane, (const uint8_t*)(blobs + off), (const uint8_t*)(commitments + off), nullptr, m, g.h_out48(),
                                   g.h_status(), g.h_status2());
            },
            [&](int lane, size_t off, int m) -> C_KZG_RET {
                Stage& g = ctx->stage[lane];
                if (any_set(g.h_status(), m) || any_set(g.h_status2(), m)) return C_KZG_BADARGS;
                memcpy(out + off, g.h_out48(), (size_t)m * 48);
                return C_KZG_OK;
            });
    });
}

// ---- the c-kzg-4844 single-blob entry points: concurrent callers are coalesced (see Coalescer) -------------------
C_KZG_RET blob_to_kzg_commitment(KZGCommitment* out, const Blob* blob, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !out || !blob) return C_KZG_BADARGS;
        return coalesced_call(*ctx, CO_COMMIT, blob->bytes, nullptr, out->bytes, nullptr);
    });
}
C_KZG_RET compute_kzg_proof(KZGProof* proof_out, Bytes32* y_out, const Blob* blob, const Bytes32* z_bytes, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !proof_out || !y_out || !blob || !z_bytes) return C_KZG_BADARGS;
        return coalesced_call(*ctx, CO_PROOF, blob->bytes, z_bytes->bytes, proof_out->bytes, y_out->bytes);
    });
}
C_KZG_RET compute_blob_kzg_proof(KZGProof* out, const Blob* blob, const Bytes48* commitment_bytes, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !out || !blob || !commitment_bytes) return C_KZG_BADARGS;
        return coalesced_call(*ctx, CO_BLOB_PROOF, blob->bytes, commitment_bytes->bytes, out->bytes, nullptr);
    });
}

// ---- device-pointer extensions (inputs resident in HBM; asynchronous on `stream`) --------------------------------
// The lane's workspace is shared with the host-pointer paths while `stream` is the caller's: the end of the enqueue is
// recorded in the lane (lane_leave_async) and the lane's next user, on whatever stream, waits on it (lane_enter).
C_KZG_RET b200_blob_to_kzg_commitment_device(void* out48_dev, const void* blobs_dev, size_t n, int* status_dev,
                                             const KZGSettings* s, void* stream) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || n > (size_t)ctx->max_batch) return C_KZG_BADARGS;
        OneLane ln(*ctx, (cudaStream_t)stream);
        ctx->dev->blob_to_commitments((const uint8_t*)blobs_dev, (int)n, (uint8_t*)out48_dev, status_dev, (cudaStream_t)stream, ln.lane);
        lane_leave_async(ctx->stage[ln.lane], (cudaStream_t)stream);
        return C_KZG_OK;
    });
}
C_KZG_RET b200_compute_kzg_proof_device(void* proofs48_dev, void* y32_dev, const void* blobs_dev, const void* z32_dev, size_t n,
                                        int z_reduce, int* status_dev, const KZGSettings* s, void* stream) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || n > (size_t)ctx->max_batch) return C_KZG_BADARGS;
        OneLane ln(*ctx, (cudaStream_t)stream);
        ctx->dev->compute_proofs((const uint8_t*)blobs_dev, (const uint8_t*)z32_dev, z_reduce, (int)n, (uint8_t*)proofs48_dev,
                                 (uint8_t*)y32_dev, status_dev, (cudaStream_t)stream, ln.lane);
        lane_leave_async(ctx->stage[ln.lane], (cudaStream_t)stream);
        return C_KZG_OK;
    });
}
// cells of compute_cells_and_kzg_proofs for n blobs (kzg/src/das.rs:244-275): cells = n x 128 x 2048 bytes
C_KZG_RET b200_compute_cells_batch(Cell* cells, const Blob* blobs, size_t n, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !cells || !blobs) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        if (!ctx->d_cells) ctx->d_cells = dev_alloc<uint8_t>((size_t)ctx->max_batch * 128 * 2048);
        uint8_t* d_cells = ctx->d_cells;
        C_KZG_RET rc = C_KZG_OK;
        {
            for (size_t off = 0; off < n && rc == C_KZG_OK; off += ctx->max_batch) {
                int m = (int)std::min<size_t>(ctx->max_batch, n - off);
                B200_CUDA_CHECK(cudaMemcpyAsync(ctx->d_blobs, blobs + off, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, ctx->stream));
                B200_CUDA_CHECK(cudaMemsetAsync(ctx->d_status, 0, m * sizeof(int), ctx->stream));
                ctx->dev->compute_cells(ctx->d_blobs, m, d_cells, ctx->d_status, ctx->stream);
                B200_CUDA_CHECK(cudaMemcpyAsync(ctx->h_status(), ctx->d_status, m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                if (any_set(ctx->h_status(), m)) { rc = C_KZG_BADARGS; break; }
                B200_CUDA_CHECK(cudaMemcpy(cells + off * 128, d_cells, (size_t)m * 128 * 2048, cudaMemcpyDeviceToHost));
            }
        }
        return rc;
    });
}
// FK20 proofs for n blobs: proofs = n x 128 x 48 bytes
C_KZG_RET b200_compute_cell_proofs_batch(KZGProof* proofs, const Blob* blobs, size_t n, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !proofs || !blobs) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        const int cap = ctx->dev->fk20_batch(ctx->stream);
        if (!ctx->d_proofs) ctx->d_proofs = dev_alloc<uint8_t>((size_t)cap * 128 * 48);
        uint8_t* d_proofs = ctx->d_proofs;
        C_KZG_RET rc = C_KZG_OK;
        {
            for (size_t off = 0; off < n; off += cap) {
                int m = (int)std::min<size_t>(cap, n - off);
                B200_CUDA_CHECK(cudaMemcpyAsync(ctx->d_blobs, blobs + off, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, ctx->stream));
                B200_CUDA_CHECK(cudaMemsetAsync(ctx->d_status, 0, m * sizeof(int), ctx->stream));
                ctx->dev->compute_cell_proofs(ctx->d_blobs, m, d_proofs, ctx->d_status, ctx->stream);
                B200_CUDA_CHECK(cudaMemcpyAsync(ctx->h_status(), ctx->d_status, m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                if (any_set(ctx->h_status(), m)) { rc = C_KZG_BADARGS; break; }
                B200_CUDA_CHECK(cudaMemcpy(proofs + off * 128, d_proofs, (size_t)m * 128 * 48, cudaMemcpyDeviceToHost));
            }
        }
        return rc;
    });
}
// cells and FK20 proofs of n blobs in one pass: the blobs cross PCIe once, the monomial form is shared, and the cells
// (2 KiB x 128 per blob) are copied out on the side stream while the proofs are still being computed
C_KZG_RET b200_compute_cells_and_kzg_proofs_batch(Cell* cells, KZGProof* proofs, const Blob* blobs, size_t n, const KZGSettings* s) {
    if (!cells && !proofs) return C_KZG_BADARGS;
    if (!proofs) return b200_compute_cells_batch(cells, blobs, n, s);
    if (!cells) return b200_compute_cell_proofs_batch(proofs, blobs, n, s);
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !blobs) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        Stage& g = ctx->stage[0];
        const int cap = std::min(ctx->max_batch, ctx->dev->fk20_batch(ctx->stream));
        if (!ctx->d_cells) ctx->d_cells = dev_alloc<uint8_t>((size_t)ctx->max_batch * 128 * 2048);
        if (!ctx->d_proofs) ctx->d_proofs = dev_alloc<uint8_t>((size_t)ctx->dev->fk20_batch(ctx->stream) * 128 * 48);
        for (size_t off = 0; off < n; off += cap) {
            const int m = (int)std::min<size_t>(cap, n - off);
            B200_CUDA_CHECK(cudaMemcpyAsync(g.d_blobs, blobs + off, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, g.stream));
            B200_CUDA_CHECK(cudaMemsetAsync(g.d_status, 0, m * sizeof(int), g.stream));
            ctx->dev->compute_cells_and_proofs(g.d_blobs, m, ctx->d_cells, ctx->d_proofs, g.d_status, g.stream, g.ev_side);
            // the blob status (Fr::from_bytes of every element) and the cells leave on the side stream under the FK20 kernels;
            // nothing is written to the caller's arrays when a blob is invalid (the reference fails before any output)
            B200_CUDA_CHECK(cudaStreamWaitEvent(g.side, g.ev_side, 0));
            B200_CUDA_CHECK(cudaMemcpyAsync(ctx->h_status(), g.d_status, m * sizeof(int), cudaMemcpyDeviceToHost, g.side));
            B200_CUDA_CHECK(cudaStreamSynchronize(g.side));
            const bool bad = any_set(ctx->h_status(), m);
            if (!bad)
                B200_CUDA_CHECK(cudaMemcpyAsync(cells + off * 128, ctx->d_cells, (size_t)m * 128 * 2048, cudaMemcpyDeviceToHost, g.side));
            B200_CUDA_CHECK(cudaStreamSynchronize(g.side));
            B200_CUDA_CHECK(cudaStreamSynchronize(g.stream));
            if (bad) return C_KZG_BADARGS;
            B200_CUDA_CHECK(cudaMemcpy(proofs + off * 128, ctx->d_proofs, (size_t)m * 128 * 48, cudaMemcpyDeviceToHost));
        }
        return C_KZG_OK;
    });
}
// kzg/src/eth/c_bindings.rs:134-199 (eip7594 macro): either output may be NULL, not both (kzg/src/das.rs:250-252)
C_KZG_RET compute_cells_and_kzg_proofs(Cell* cells, KZGProof* proofs, const Blob* blob, const KZGSettings* s) {
    if (!blob || (!cells && !proofs)) return C_KZG_BADARGS;
    if (!proofs) return b200_compute_cells_batch(cells, blob, 1, s);            // cells alone: two NTTs, nothing to share
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx) return C_KZG_BADARGS;
        if (ctx->cells_cap < 2) return b200_compute_cells_and_kzg_proofs_batch(cells, proofs, blob, 1, s);
        return coalesced_cells_call(*ctx, blob->bytes, (uint8_t*)cells, (uint8_t*)proofs);
    });
}

// ---- verification (blst/src/eip_4844.rs:383-471) -----------------------------------------------------------------
// the batch challenge of compute_r_powers (kzg/src/eip_4844.rs:328-378); z / y are the canonical encodings
static void batch_challenge_hash(uint8_t out[32], const uint8_t* c48, const uint8_t* z32, const uint8_t* y32, const uint8_t* p48, size_t n) {
    sha256::Ctx c;
    uint8_t head[32] = {'R', 'C', 'K', 'Z', 'G', 'B', 'A', 'T', 'C', 'H', '_', '_', '_', 'V', '1', '_'};
    for (int i = 0; i < 8; i++) {
        head[16 + i] = (uint8_t)((uint64_t)kFieldElementsPerBlob >> (8 * (7 - i)));
        head[24 + i] = (uint8_t)((uint64_t)n >> (8 * (7 - i)));
    }
    c.update(head, 32);
    for (size_t i = 0; i < n; i++) {
        c.update(c48 + 48 * i, 48);
        c.update(z32 + 32 * i, 32);
        c.update(y32 + 32 * i, 32);
        c.update(p48 + 48 * i, 48);
    }
    c.finish(out);
}
// verify_kzg_proof_batch (kzg/src/eip_4844.rs:380-435) on host arrays of wire bytes; ctx->mu held by the caller.
// Stage 1 (points): copy + decode + subgroup-check the commitments and proofs on the side stream -- for the blob
// verifiers this runs under the blob transfer, the host hashing and the polynomial evaluations.
// Stage 2 (finish): z, y and the Fiat-Shamir challenge arrive, two short lincombs and the pairing run on the main stream.
struct VerifyLayout {
    size_t n, o_p, o_z, o_y, o_r, o_st, o_res;
    explicit VerifyLayout(size_t n_) : n(n_), o_p(48 * n_), o_z(96 * n_), o_y(128 * n_), o_r(160 * n_), o_st(160 * n_ + 32),
                                       o_res(160 * n_ + 32 + n_ * sizeof(int)) {}
};
static void verify_stage_points(KzgCtx& ctx, const uint8_t* c48, const uint8_t* p48, size_t n) {
    ctx.ensure_verify(n);
    VerifyLayout L(n);
    uint8_t *h = ctx.h_verify, *d = ctx.d_verify;
    Stage& g = ctx.stage[0];
    memcpy(h, c48, 48 * n);
    memcpy(h + L.o_p, p48, 48 * n);
    B200_CUDA_CHECK(cudaMemcpyAsync(d, h, 96 * n, cudaMemcpyHostToDevice, g.side));
    B200_CUDA_CHECK(cudaMemsetAsync(d + L.o_st, 0, n * sizeof(int) + sizeof(int), g.side));
    // ev_dec: coordinates decoded (all the lincombs and the pairing need); ev_side: subgroup tests done too (the status)
    ctx.dev->verify_decode(d, d + L.o_p, (int)n, reinterpret_cast<int*>(d + L.o_st), g.side, g.ev_dec);
    B200_CUDA_CHECK(cudaEventRecord(g.ev_side, g.side));
}
static C_KZG_RET verify_finish(KzgCtx& ctx, bool* ok, const uint8_t* c48, const uint8_t* z32, const uint8_t* y32, const uint8_t* p48, size_t n) {
    VerifyLayout L(n);
    uint8_t *h = ctx.h_verify, *d = ctx.d_verify;
    memcpy(h + L.o_z, z32, 32 * n);
    memcpy(h + L.o_y, y32, 32 * n);
    if (n > 1) batch_challenge_hash(h + L.o_r, c48, z32, y32, p48, n); else memset(h + L.o_r, 0, 32);
    cudaStream_t st = ctx.stream;
    // the lincombs and the pairing need the decoded coordinates only; the subgroup tests of the same points keep running on
    // the side stream beside them (they only ever set status flags) and are joined before the status is read
    B200_CUDA_CHECK(cudaStreamWaitEvent(st, ctx.stage[0].ev_dec, 0));
    B200_CUDA_CHECK(cudaMemcpyAsync(d + L.o_z, h + L.o_z, 64 * n + 32, cudaMemcpyHostToDevice, st));
    ctx.dev->verify_batch(d, d + L.o_p, d + L.o_z, d + L.o_y, 0, d + L.o_r, (int)n, reinterpret_cast<int*>(d + L.o_st),
                          reinterpret_cast<int*>(d + L.o_res), st, /*skip_decode=*/true);
    B200_CUDA_CHECK(cudaStreamWaitEvent(st, ctx.stage[0].ev_side, 0));
    B200_CUDA_CHECK(cudaMemcpyAsync(h + L.o_st, d + L.o_st, n * sizeof(int) + sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    if (any_set(reinterpret_cast<int*>(h + L.o_st), (int)n)) return C_KZG_BADARGS;
    *ok = *reinterpret_cast<int*>(h + L.o_res) != 0;
    return C_KZG_OK;
}
static C_KZG_RET verify_core(KzgCtx& ctx, bool* ok, const uint8_t* c48, const uint8_t* z32, const uint8_t* y32, const uint8_t* p48, size_t n) {
    verify_stage_points(ctx, c48, p48, n);
    return verify_finish(ctx, ok, c48, z32, y32, p48, n);
}
/* b200 extension: verify_kzg_proof_batch over caller-supplied (commitment, z, y, proof) tuples */
C_KZG_RET b200_verify_kzg_proof_batch(bool* ok, const Bytes48* commitments, const Bytes32* zs, const Bytes32* ys, const Bytes48* proofs,
                                      size_t n, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !ok) return C_KZG_BADARGS;
        *ok = false;
        if (n == 0) { *ok = true; return C_KZG_OK; }
        if (!commitments || !zs || !ys || !proofs) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        return verify_core(*ctx, ok, (const uint8_t*)commitments, (const uint8_t*)zs, (const uint8_t*)ys, (const uint8_t*)proofs, n);
    });
}
static C_KZG_RET coalesced_verify(const std::shared_ptr<KzgCtx>& ctxp, int kind, bool* ok, const uint8_t* blob, const uint8_t* c48,
                                  const uint8_t* z32, const uint8_t* y32, const uint8_t* p48, bool* fall_back);
C_KZG_RET verify_kzg_proof(bool* ok, const Bytes48* commitment_bytes, const Bytes32* z_bytes, const Bytes32* y_bytes,
                           const Bytes48* proof_bytes, const KZGSettings* s) {
    if (ok && commitment_bytes && z_bytes && y_bytes && proof_bytes) {
        bool fall_back = true;
        C_KZG_RET rc = ckzg_guard([&]() -> C_KZG_RET {
            auto ctx = find_ctx(s);
            if (!ctx) return C_KZG_BADARGS;
            if (ctx->vco_cap < 2) return C_KZG_OK;
            return coalesced_verify(ctx, VK_PROOF, ok, nullptr, commitment_bytes->bytes, z_bytes->bytes, y_bytes->bytes, proof_bytes->bytes,
                                    &fall_back);
        });
        if (rc != C_KZG_OK || !fall_back) return rc;
    }
    return b200_verify_kzg_proof_batch(ok, commitment_bytes, z_bytes, y_bytes, proof_bytes, 1, s);
}
// verify_blob_kzg_proof_batch on host arrays; the caller holds every lane
static C_KZG_RET verify_blobs_locked(const std::shared_ptr<KzgCtx>& ctx, bool* ok, const Blob* blobs, const Bytes48* commitments_bytes,
                                     const Bytes48* proofs_bytes, size_t n) {
    {
        // phase 1, chunked over the two lanes: z_i = challenge(blob_i, C_i) hashed on the host while the blobs cross
        // PCIe, y_i = p_i(z_i) on the device (compute_challenges_and_evaluate_polynomial, :700-718)
        std::vector<uint8_t> zs(32 * n), ys(32 * n);
        verify_stage_points(*ctx, (const uint8_t*)commitments_bytes, (const uint8_t*)proofs_bytes, n);
        C_KZG_RET rc = run_chunks(*ctx, n, ctx->max_batch,
            [&](int lane, size_t off, int m) {
                Stage& g = ctx->stage[lane];
                B200_CUDA_CHECK(cudaMemcpyAsync(g.d_blobs, blobs + off, (size_t)m * kBytesPerBlob, cudaMemcpyHostToDevice, g.stream));
                B200_CUDA_CHECK(cudaMemsetAsync(g.d_status, 0, m * sizeof(int), g.stream));
                challenge_hash_many(g.h_z(), (const uint8_t*)(blobs + off), (const uint8_t*)(commitments_bytes + off), m);
                B200_CUDA_CHECK(cudaMemcpyAsync(g.d_z, g.h_z(), (size_t)m * 32, cudaMemcpyHostToDevice, g.stream));
                ctx->dev->evaluate_blobs(g.d_blobs, g.d_z, 1, m, g.d_out48, g.d_y32, g.d_status, g.stream, lane);
                B200_CUDA_CHECK(cudaMemcpyAsync(g.h_out48(), g.d_out48, (size_t)m * 32, cudaMemcpyDeviceToHost, g.stream));
                B200_CUDA_CHECK(cudaMemcpyAsync(g.h_y32(), g.d_y32, (size_t)m * 32, cudaMemcpyDeviceToHost, g.stream));
                B200_CUDA_CHECK(cudaMemcpyAsync(g.h_status(), g.d_status, m * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
            },
            [&](int lane, size_t off, int m) -> C_KZG_RET {
                Stage& g = ctx->stage[lane];
                if (any_set(g.h_status(), m)) return C_KZG_BADARGS;
                memcpy(zs.data() + 32 * off, g.h_out48(), (size_t)m * 32);
                memcpy(ys.data() + 32 * off, g.h_y32(), (size_t)m * 32);
                return C_KZG_OK;
            });
        if (rc != C_KZG_OK) {
            cudaStreamSynchronize(ctx->stage[0].side);   // the early point decoding must not outlive this call's buffers
            return rc;
        }
        // phase 2: one batched pairing check over all n (verify_kzg_proof_batch, :380-435)
        return verify_finish(*ctx, ok, (const uint8_t*)commitments_bytes, zs.data(), ys.data(), (const uint8_t*)proofs_bytes, n);
    }
}
C_KZG_RET verify_blob_kzg_proof_batch(bool* ok, const Blob* blobs, const Bytes48* commitments_bytes, const Bytes48* proofs_bytes, size_t n,
                                      const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !ok) return C_KZG_BADARGS;
        *ok = false;
        if (n == 0) { *ok = true; return C_KZG_OK; }  // kzg/src/eip_4844.rs:760-763
        if (!blobs || !commitments_bytes || !proofs_bytes) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        return verify_blobs_locked(ctx, ok, blobs, commitments_bytes, proofs_bytes, n);
    });
}
// One single verification (kind VK_BLOB: blob + commitment + proof; VK_PROOF: commitment + z + y + proof) among concurrent
// ones: staged into the open batch, checked together, and re-checked alone only if the batch as a whole does not pass.
static C_KZG_RET coalesced_verify(const std::shared_ptr<KzgCtx>& ctxp, int kind, bool* ok, const uint8_t* blob, const uint8_t* c48,
                                  const uint8_t* z32, const uint8_t* y32, const uint8_t* p48, bool* fall_back) {
    KzgCtx& ctx = *ctxp;
    VerifyCoalescer& co = ctx.co_verify;
    const int cap = ctx.vco_cap;
    VerifyCoalescer::Claim cl = co.claim(kind, cap, [&] {
        std::unique_ptr<VerifyBatch> nb(new VerifyBatch());
        nb->cap = cap;
        DeviceScope ds(ctx.device);
        B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_in, (size_t)cap * (kBytesPerBlob + 48 + 48 + 32 + 32)));
        return nb;
    });
    VerifyBatch* B = cl.b;
    const int idx = cl.idx;
    if (kind == VK_BLOB) memcpy(B->blob(idx), blob, kBytesPerBlob);
    memcpy(B->comm(idx), c48, 48);
    memcpy(B->proof(idx), p48, 48);
    if (kind == VK_PROOF) { memcpy(B->z(idx), z32, 32); memcpy(B->y(idx), y32, 32); }
    VerifyCoalescer::staged(B);
    if (cl.leader) {
        int rc = C_KZG_OK;
        bool all_ok = false;
        B->all_ok = false;                                      // batch objects are pooled: no verdict of an earlier use may survive
        B->single_done = false;
        try {
            DeviceScope ds(ctx.device);
            AllLanes lk(ctx);   // while another pass runs this batch fills
            if (ctx.cells_grace_us > 0) {                       // a burst of sidecars: linger briefly, as for the cells
                const auto t0 = std::chrono::steady_clock::now();
                auto last_change = t0;
                int seen = co.claimed_so_far(B);
                while (seen < cap) {
                    std::this_thread::yield();                  // (a 15 us sleep oversleeps by the timer slack: ~60 us each)
                    const auto now = std::chrono::steady_clock::now();
                    const int c = co.claimed_so_far(B);
                    if (c != seen) { seen = c; last_change = now; }
                    if (now - last_change > std::chrono::microseconds(60) || now - t0 > std::chrono::microseconds(ctx.cells_grace_us)) break;
                }
            }
            const int n = co.close(B);
            bool bok = false;
            C_KZG_RET r = kind == VK_BLOB
                ? verify_blobs_locked(ctxp, &bok, (const Blob*)B->blob(0), (const Bytes48*)B->comm(0), (const Bytes48*)B->proof(0), (size_t)n)
                : verify_core(ctx, &bok, B->comm(0), B->z(0), B->y(0), B->proof(0), (size_t)n);
            all_ok = r == C_KZG_OK && bok;                     // anything else: every caller re-checks its own request
            ctx.st_verify_batches++;
            ctx.st_verify_requests += (uint64_t)n;
            if (!all_ok && n > 1) ctx.st_verify_fallbacks++;
            if (n == 1 && r != C_KZG_OK) rc = r;               // a batch of one IS the single check: its verdict stands
            if (n == 1) B->single_done = true;
        } catch (const std::exception& e) {
            cudaGetLastError();
            fprintf(stderr, "b200kzg: %s\n", e.what());
            rc = C_KZG_ERROR;
        }
        B->all_ok = all_ok;
        co.publish(B, rc);
    } else {
        co.wait(B);
    }
    C_KZG_RET rc = (C_KZG_RET)B->rc;
    const bool all_ok = B->all_ok, single_done = B->single_done;
    co.consume(B);
    *fall_back = false;
    if (rc != C_KZG_OK) { *ok = false; return rc; }
    if (all_ok) { *ok = true; return C_KZG_OK; }
    if (single_done) { *ok = false; return C_KZG_OK; }         // the lone request was checked exactly and is invalid
    *fall_back = true;
    return C_KZG_OK;
}
C_KZG_RET verify_blob_kzg_proof(bool* ok, const Blob* blob, const Bytes48* commitment_bytes, const Bytes48* proof_bytes, const KZGSettings* s) {
    if (ok && blob && commitment_bytes && proof_bytes) {
        bool fall_back = true;
        C_KZG_RET rc = ckzg_guard([&]() -> C_KZG_RET {
            auto ctx = find_ctx(s);
            if (!ctx) return C_KZG_BADARGS;
            if (ctx->vco_cap < 2) return C_KZG_OK;             // coalescing off: the plain single check below
            return coalesced_verify(ctx, VK_BLOB, ok, blob->bytes, commitment_bytes->bytes, nullptr, nullptr, proof_bytes->bytes, &fall_back);
        });
        if (rc != C_KZG_OK || !fall_back) return rc;
    }
    return verify_blob_kzg_proof_batch(ok, blob, commitment_bytes, proof_bytes, 1, s);
}
// ---- EIP-7594 recovery and cell verification (kzg/src/eth/c_bindings.rs:201-352, blst/src/eip_7594.rs:35-97) ------
static constexpr size_t kCellsPerExtBlob = 128, kBytesPerCell = 2048;
// Fiat-Shamir input of compute_verify_cell_kzg_proof_batch_challenge (kzg/src/das.rs:390-452)
static void cell_challenge_hash(uint8_t out[32], const uint8_t* comm48, size_t m, const uint64_t* comm_idx, const uint64_t* cell_idx,
                                const uint8_t* cells, const uint8_t* proofs48, size_t n) {
    sha256::Ctx c;
    uint8_t head[48] = {'R', 'C', 'K', 'Z', 'G', 'C', 'B', 'A', 'T', 'C', 'H', '_', '_', 'V', '1', '_'};
    const uint64_t hv[4] = {kFieldElementsPerBlob, 64, m, n};
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 8; i++) head[16 + 8 * k + i] = (uint8_t)(hv[k] >> (8 * (7 - i)));
    c.update(head, 48);
    c.update(comm48, 48 * m);
    for (size_t i = 0; i < n; i++) {
        uint8_t idx[16];
        for (int b = 0; b < 8; b++) {
            idx[b] = (uint8_t)(comm_idx[i] >> (8 * (7 - b)));
            idx[8 + b] = (uint8_t)(cell_idx[i] >> (8 * (7 - b)));
        }
        c.update(idx, 16);
        c.update(cells + i * kBytesPerCell, kBytesPerCell);
        c.update(proofs48 + 48 * i, 48);
    }
    c.finish(out);
}
struct DevScratch {  // device + pinned staging: borrowed from the context (caller holds ctx.mu) or owned for one call
    uint8_t* d = nullptr;
    uint8_t* h = nullptr;
    bool owned = true;
    explicit DevScratch(size_t bytes) {
        d = dev_alloc<uint8_t>(bytes);
        if (cudaMallocHost((void**)&h, bytes ? bytes : 16) != cudaSuccess) { cudaFree(d); throw CudaError(-1, "cudaMallocHost failed"); }
    }
    DevScratch(KzgCtx& ctx, size_t bytes) : owned(false) {
        ctx.ensure_scratch(bytes);
        d = ctx.d_scratch;
        h = ctx.h_scratch;
    }
    ~DevScratch() {
        if (!owned) return;
        cudaFree(d);
        if (h) cudaFreeHost(h);
    }
};
C_KZG_RET recover_cells_and_kzg_proofs(Cell* recovered_cells, KZGProof* recovered_proofs, const uint64_t* cell_indices, const Cell* cells,
                                       uint64_t num_cells, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !recovered_cells || (num_cells && (!cell_indices || !cells))) return C_KZG_BADARGS;
        const size_t n = num_cells;
        // kzg/src/das.rs:126-165: at most 128 cells, at least 64, indices < 128 and strictly ascending
        if (n > kCellsPerExtBlob || n < kCellsPerExtBlob / 2) return C_KZG_BADARGS;
        for (size_t i = 0; i < n; i++) {
            if (cell_indices[i] >= kCellsPerExtBlob) return C_KZG_BADARGS;
            if (i + 1 < n && cell_indices[i + 1] <= cell_indices[i]) return C_KZG_BADARGS;
        }
        AllLanes lk(*ctx);
        const size_t o_out = n * kBytesPerCell, o_pr = o_out + kCellsPerExtBlob * kBytesPerCell, o_st = o_pr + kCellsPerExtBlob * 48;
        DevScratch buf(*ctx, o_st + 64);
        cudaStream_t st = ctx->stream;
        memcpy(buf.h, cells, n * kBytesPerCell);
        memset(buf.h + o_st, 0, sizeof(int));
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.d, buf.h, n * kBytesPerCell, cudaMemcpyHostToDevice, st));
        B200_CUDA_CHECK(cudaMemsetAsync(buf.d + o_st, 0, sizeof(int), st));
        ctx->dev->recover_cells(buf.d, cell_indices, (int)n, buf.d + o_out, recovered_proofs ? buf.d + o_pr : nullptr,
                                reinterpret_cast<int*>(buf.d + o_st), st);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.h + o_out, buf.d + o_out, o_st + sizeof(int) - o_out, cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (*reinterpret_cast<int*>(buf.h + o_st)) return C_KZG_BADARGS;
        memcpy(recovered_cells, buf.h + o_out, kCellsPerExtBlob * kBytesPerCell);
        if (recovered_proofs) memcpy(recovered_proofs, buf.h + o_pr, kCellsPerExtBlob * 48);
        return C_KZG_OK;
    });
}
C_KZG_RET verify_cell_kzg_proof_batch(bool* ok, const Bytes48* commitments_bytes, const uint64_t* cell_indices, const Cell* cells,
                                      const Bytes48* proofs_bytes, uint64_t num_cells, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ok) return C_KZG_BADARGS;
        *ok = false;
        if (!ctx) return C_KZG_BADARGS;
        const size_t n = num_cells;
        if (n == 0) { *ok = true; return C_KZG_OK; }   // kzg/src/das.rs:319-321
        if (!commitments_bytes || !cell_indices || !cells || !proofs_bytes) return C_KZG_BADARGS;
        for (size_t i = 0; i < n; i++)
            if (cell_indices[i] >= kCellsPerExtBlob) return C_KZG_BADARGS;
        // deduplicate_with_indices (kzg/src/das.rs:57-76) on the canonical encodings
        std::map<std::string, uint64_t> seen;
        std::vector<uint8_t> uniq;
        std::vector<uint64_t> comm_idx(n);
        for (size_t i = 0; i < n; i++) {
            std::string key((const char*)commitments_bytes[i].bytes, 48);
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, seen.size()).first;
                uniq.insert(uniq.end(), commitments_bytes[i].bytes, commitments_bytes[i].bytes + 48);
            }
            comm_idx[i] = it->second;
        }
        const size_t m = seen.size();
        AllLanes lk(*ctx);
        // device layout: [cells n*2048][proofs n*48][uniq m*48][comm_idx n u32][cell_idx n u32][r 32][status n ints][result]
        const size_t o_p = n * kBytesPerCell, o_c = o_p + 48 * n, o_ci = (o_c + 48 * m + 15) & ~(size_t)15, o_ki = o_ci + 4 * n,
                     o_r = o_ki + 4 * n, o_st = o_r + 32, o_res = o_st + 4 * n;
        DevScratch buf(*ctx, o_res + 16);
        memcpy(buf.h, cells, n * kBytesPerCell);
        memcpy(buf.h + o_p, proofs_bytes, 48 * n);
        memcpy(buf.h + o_c, uniq.data(), 48 * m);
        uint32_t* ci = reinterpret_cast<uint32_t*>(buf.h + o_ci);
        uint32_t* ki = reinterpret_cast<uint32_t*>(buf.h + o_ki);
        for (size_t i = 0; i < n; i++) { ci[i] = (uint32_t)comm_idx[i]; ki[i] = (uint32_t)cell_indices[i]; }
        cell_challenge_hash(buf.h + o_r, uniq.data(), m, comm_idx.data(), cell_indices, (const uint8_t*)cells, (const uint8_t*)proofs_bytes, n);
        memset(buf.h + o_st, 0, 4 * n + 4);
        cudaStream_t st = ctx->stream;
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.d, buf.h, o_res + 4, cudaMemcpyHostToDevice, st));
        ctx->dev->verify_cells(buf.d + o_c, (int)m, reinterpret_cast<uint32_t*>(buf.d + o_ci), reinterpret_cast<uint32_t*>(buf.d + o_ki), buf.d,
                               buf.d + o_p, buf.d + o_r, (int)n, reinterpret_cast<int*>(buf.d + o_st), reinterpret_cast<int*>(buf.d + o_res), st);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.h + o_st, buf.d + o_st, 4 * n + 4, cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (any_set(reinterpret_cast<int*>(buf.h + o_st), (int)n)) return C_KZG_BADARGS;
        *ok = *reinterpret_cast<int*>(buf.h + o_res) != 0;
        return C_KZG_OK;
    });
}
// blst/src/eip_7594.rs:35-97: the challenge as a Montgomery blst_fr; every argument must decode
C_KZG_RET compute_verify_cell_kzg_proof_batch_challenge(blst_fr* challenge_out, const Bytes48* commitment_bytes, uint64_t num_commitments,
                                                        const uint64_t* commitment_indices, const uint64_t* cell_indices, const Cell* cells,
                                                        const Bytes48* proofs_bytes, uint64_t num_cells) {
    return ckzg_guard([&]() -> C_KZG_RET {
        if (!challenge_out) return C_KZG_BADARGS;
        memset(challenge_out, 0, sizeof(*challenge_out));
        const size_t n = num_cells, m = num_commitments;
        if (m && !commitment_bytes) return C_KZG_BADARGS;
        if (n && (!commitment_indices || !cell_indices || !cells || !proofs_bytes)) return C_KZG_BADARGS;
        require_device();
        // Standalone like the reference's function: no settings object, own scratch, the default stream of the current device.
        const size_t o_p = n * kBytesPerCell, o_c = o_p + 48 * n, o_r = (o_c + 48 * m + 15) & ~(size_t)15, o_fr = o_r + 32, o_st = o_fr + 32,
                     o_ws = (o_st + 16 + 255) & ~(size_t)255;
        DevScratch buf(o_ws + check_challenge_ws_bytes((int)m, (int)n));
        if (n) { memcpy(buf.h, cells, n * kBytesPerCell); memcpy(buf.h + o_p, proofs_bytes, 48 * n); }
        if (m) memcpy(buf.h + o_c, commitment_bytes, 48 * m);
        cell_challenge_hash(buf.h + o_r, (const uint8_t*)commitment_bytes, m, commitment_indices, cell_indices, (const uint8_t*)cells,
                            (const uint8_t*)proofs_bytes, n);
        memset(buf.h + o_st, 0, 4);
        cudaStream_t st = nullptr;
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.d, buf.h, o_st + 4, cudaMemcpyHostToDevice, st));
        int* d_st = reinterpret_cast<int*>(buf.d + o_st);
        launch_check_challenge_inputs(buf.d + o_ws, buf.d + o_c, (int)m, buf.d, buf.d + o_p, (int)n, d_st, st);
        launch_fr_from_bytes(buf.d + o_r, 1, 1, buf.d + o_fr, d_st, st);   // hash_to_bls_field -> Montgomery
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.h + o_fr, buf.d + o_fr, 32 + 4, cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (*reinterpret_cast<int*>(buf.h + o_st)) return C_KZG_BADARGS;
        memcpy(challenge_out, buf.h + o_fr, 32);
        return C_KZG_OK;
    });
}

// ---- the three helper exports of blst/src/eip_4844.rs:498-530 (no settings argument: they run on the default stream) ----
void compute_challenge(blst_fr* eval_challenge_out, const Blob* blob, const blst_p1* commitment) {
    if (!eval_challenge_out) return;
    memset(eval_challenge_out, 0, sizeof(*eval_challenge_out));
    if (!blob || !commitment) return;
    ckzg_guard([&]() -> C_KZG_RET {
        require_device();
        DevScratch buf(144 + 48 + 32 + 32);
        memcpy(buf.h, commitment, 144);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.d, buf.h, 144, cudaMemcpyHostToDevice, nullptr));
        launch_points_to_compressed(buf.d, buf.d + 144, 1, nullptr);          // commitment.to_bytes() (kzg/src/eip_4844.rs:936-938)
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.h + 144, buf.d + 144, 48, cudaMemcpyDeviceToHost, nullptr));
        B200_CUDA_CHECK(cudaStreamSynchronize(nullptr));
        challenge_hash(buf.h + 192, blob->bytes, buf.h + 144);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.d + 192, buf.h + 192, 32, cudaMemcpyHostToDevice, nullptr));
        launch_fr_from_bytes(buf.d + 192, 1, 1, buf.d + 224, nullptr, nullptr);  // hash_to_bls_field -> Montgomery
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.h + 224, buf.d + 224, 32, cudaMemcpyDeviceToHost, nullptr));
        B200_CUDA_CHECK(cudaStreamSynchronize(nullptr));
        memcpy(eval_challenge_out, buf.h + 224, 32);
        return C_KZG_OK;
    });
}
C_KZG_RET bytes_to_kzg_commitment(blst_p1* out, const Bytes48* b) {
    return ckzg_guard([&]() -> C_KZG_RET {
        if (!out || !b) return C_KZG_BADARGS;
        require_device();
        DevScratch buf(48 + 96 + 144 + 16);
        memcpy(buf.h, b->bytes, 48);
        memset(buf.h + 288, 0, 4);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.d, buf.h, 48, cudaMemcpyHostToDevice, nullptr));
        B200_CUDA_CHECK(cudaMemsetAsync(buf.d + 288, 0, 4, nullptr));
        launch_uncompress_g1(buf.d, buf.d + 48, reinterpret_cast<int*>(buf.d + 288), 1, nullptr);   // FsG1::from_bytes: no subgroup check
        launch_affine_to_jac(buf.d + 48, buf.d + 144, 1, nullptr);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.h + 144, buf.d + 144, 148, cudaMemcpyDeviceToHost, nullptr));
        B200_CUDA_CHECK(cudaStreamSynchronize(nullptr));
        if (*reinterpret_cast<int*>(buf.h + 288)) return C_KZG_BADARGS;
        memcpy(out, buf.h + 144, 144);
        return C_KZG_OK;
    });
}
void bytes_from_bls_field(Bytes32* out, const blst_fr* inp) {
    if (!out || !inp) return;
    ckzg_guard([&]() -> C_KZG_RET {
        require_device();
        DevScratch buf(64);
        memcpy(buf.h, inp, 32);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.d, buf.h, 32, cudaMemcpyHostToDevice, nullptr));
        launch_fr_to_bytes(buf.d, 1, buf.d + 32, nullptr);
        B200_CUDA_CHECK(cudaMemcpyAsync(buf.h + 32, buf.d + 32, 32, cudaMemcpyDeviceToHost, nullptr));
        B200_CUDA_CHECK(cudaStreamSynchronize(nullptr));
        memcpy(out->bytes, buf.h + 32, 32);
        return C_KZG_OK;
    });
}

/* test hook: out = sum scalars[i] * points[i] through the lane-quad GLV scalar multiplication (host arrays) */
RustError b200_selftest_lincomb_quads(blst_p1* out, const blst_p1_affine* points, const blst_fr* scalars, size_t n) {
    return guarded([&] {
        require_device();
        if (!out || !points || !scalars || n == 0) throw CudaError(-1, "bad arguments");
        uint8_t* d = dev_alloc<uint8_t>(n * (96 + 32) + 144);
        cudaError_t e = cudaMemcpy(d, points, n * 96, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d + n * 96, scalars, n * 32, cudaMemcpyHostToDevice);
        try {
            B200_CUDA_CHECK(e);
            selftest_lincomb_quads(d, d + n * 96, (int)n, d + n * 128, nullptr);
            B200_CUDA_CHECK(cudaMemcpy(out, d + n * 128, 144, cudaMemcpyDeviceToHost));
        } catch (...) {
            cudaFree(d);
            throw;
        }
        cudaFree(d);
    });
}

/* test hook for the pairing alone: e(a1, Q[qa]) == e(b1, Q[qb]), Q = {[1]G2, [s]G2, [s^64]G2}; host Jacobian points */
C_KZG_RET b200_selftest_pairings_verify(bool* ok, const blst_p1* a1, int qa, const blst_p1* b1, int qb, const KZGSettings* s) {
    return ckzg_guard([&]() -> C_KZG_RET {
        auto ctx = find_ctx(s);
        if (!ctx || !ok || !a1 || !b1) return C_KZG_BADARGS;
        AllLanes lk(*ctx);
        ctx->ensure_verify(4);
        uint8_t* d = ctx->d_verify;
        B200_CUDA_CHECK(cudaMemcpyAsync(d, a1, 144, cudaMemcpyHostToDevice, ctx->stream));
        B200_CUDA_CHECK(cudaMemcpyAsync(d + 144, b1, 144, cudaMemcpyHostToDevice, ctx->stream));
        int* d_res = reinterpret_cast<int*>(d + 288);
        ctx->dev->pairings_verify(d, qa, d + 144, qb, d_res, ctx->stream);
        int res = 0;
        B200_CUDA_CHECK(cudaMemcpyAsync(&res, d_res, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        *ok = res != 0;
        return C_KZG_OK;
    });
}

int b200_kzg_launches(const KZGSettings* s) {
    auto ctx = find_ctx(s);
    return ctx ? ctx->dev->launches_last() : 0;
}
// coalescer counters since load: [batches, requests, ns waiting for a lane, ns on a lane (enqueue to drained), largest batch]
void b200_kzg_coalesce_stats(const KZGSettings* s, uint64_t out[5]) {
    auto ctx = find_ctx(s);
    if (!out) return;
    for (int i = 0; i < 5; i++) out[i] = 0;
    if (!ctx) return;
    out[0] = ctx->st_batches; out[1] = ctx->st_requests; out[2] = ctx->st_wait_ns; out[3] = ctx->st_exec_ns; out[4] = ctx->st_max_batch;
}
// window widths of the direct-lookup tables this settings object holds: out = [Lagrange table bits, largest batch it serves,
// FK20 column table bits (0 before the first cell-proof call or when the bucket engine serves the lincombs)]
void b200_kzg_direct_tables(const KZGSettings* s, int out[3]) {
    auto ctx = find_ctx(s);
    if (!out) return;
    out[0] = out[1] = out[2] = 0;
    if (!ctx) return;
    out[0] = ctx->dev->direct_bits();
    out[1] = ctx->dev->direct_max_batch();
    out[2] = ctx->dev->fk_direct_bits();
}
// coalesced single verifications: out = [batches checked, requests served, batches that did not pass as a whole (their callers
// re-checked alone)]
void b200_kzg_verify_coalesce_stats(const KZGSettings* s, uint64_t out[3]) {
    auto ctx = find_ctx(s);
    if (!out) return;
    out[0] = out[1] = out[2] = 0;
    if (!ctx) return;
    out[0] = ctx->st_verify_batches; out[1] = ctx->st_verify_requests; out[2] = ctx->st_verify_fallbacks;
}
// the same counters for coalesced compute_cells_and_kzg_proofs calls: out = [batches run, requests served]
void b200_kzg_cells_coalesce_stats(const KZGSettings* s, uint64_t out[2]) {
    auto ctx = find_ctx(s);
    if (!out) return;
    out[0] = out[1] = 0;
    if (!ctx) return;
    out[0] = ctx->st_cells_batches; out[1] = ctx->st_cells_requests;
}
int b200_kzg_max_batch(const KZGSettings* s) {
    auto ctx = find_ctx(s);
    return ctx ? ctx->max_batch : 0;
}
// test hook: SHA-256 of a host buffer (portable != 0 forces the non-SHA-NI code path)
void b200_selftest_sha256(uint8_t out[32], const uint8_t* msg, size_t len, int portable) {
    sha256::digest(out, msg, len, portable != 0);
}

}  // extern "C"
