// capi_msm.cu -- the sppark-shaped MSM FFI (include/b200_kzg.h, section B1) on top of MsmEngine.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/b200_kzg.h"
#include "capi_common.cuh"
#include "coalesce.cuh"
#include "eip4844.cuh"
#include "g1.cuh"
#include "msm.cuh"
#include "util.cuh"

using namespace b200;

namespace b200 {

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// window choice.  Work = n * W mixed additions + O(2^(c-1)) reduce additions; the reduce is latency-bound, so
// stay a notch below the pure work optimum.  Overridable for tuning: B200_MSM_C, B200_MSM_L.
MsmConfig choose_config(size_t n, bool fixed, int max_batch) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int c;
    // Window width.  Scalars are < r < 2^255 and, on a prepared table, uniform after the scalar randomisation (msm.cu), so
    // the TOP window holds t = 255 - (W - 1) c bits and its 2^(t-1) buckets receive n entries: a width is only usable when
    // that load, 2^(c - t) / W times the average, is moderate.  c = 13 (t = 8), 15 (t = 15), 16 (15), 17 (17) and 20 (15) qualify;
    // 12, 14, 18, 19, 21, 22 leave the top window 3 - 8 bits against 2^11 .. 2^21 buckets elsewhere (measured at 2^20:
    // c = 16 / 18 / 19 / 20 / 21 -> 7.22 / 14.4 / 9.85 / 6.3 / 14.6 ms, scripts/msm_window_sweep.py).
    // fixed, >= 2^20 points: c = 20, W = 13, with the segment fold in front of the 15-bit reduce.
    if (fixed) c = lg >= 20 ? 20 : lg >= 18 ? 16 : lg >= 16 ? 15 : lg >= 11 ? 13 : lg >= 8 ? 10 : 8;
    else c = lg >= 20 ? 16 : lg >= 16 ? 15 : lg >= 12 ? 13 : lg >= 9 ? 10 : 8;
    c = env_int(fixed ? "B200_MSM_C" : "B200_MSM_VC", c);
    if (c < 4) c = 4;
    // beyond 16 the bucket set is folded by segments before the 15-bit reduce (k_segment_fold); a variable-base call has
    // one bucket set per window and the scans take at most 2^24 keys
    const int cmax = fixed ? 22 : 20;
    if (c > cmax) c = cmax;
    MsmConfig cfg;
    cfg.c = c;
    cfg.W = (256 + c - 1) / c;
    cfg.c0 = fixed ? env_int("B200_MSM_C0", 0) : 0;   // narrower window 0 (tests / tuning); must keep c0 + (W-1) c >= 256
    cfg.fixed = fixed;
    cfg.n = n;
    cfg.max_batch = fixed ? max_batch : 1;
    cfg.L = env_int("B200_MSM_L", 64);
    // 13-bit windows over a few thousand points: a 2-bit segment fold in front of the marginal sums (measured on the 64-vector
    // batches of the blob pipeline, 3.33 -> 2.99 ms: scripts/blob_window_sweep.py)
    if (fixed && c == 13) cfg.fold = env_int("B200_MSM_FOLD", 2);
    // batch-affine accumulation (6 instead of 10 field multiplications per bucket addition, k_accumulate_affine): bit-exact
    // and tested, but OFF by default -- measured on B200 it loses to the XYZZ task kernel (7.3 - 7.9 ms against 4.95 ms at
    // 2^20): the 24-36 lock-step rounds each end in one field inversion whose latency (~60 us even with the bingcd inverse)
    // the two or three resident CTAs cannot hide (profiles/r02_affine.md).  B200_MSM_AFFINE=1 turns it on for A/B runs.
    cfg.affine = fixed && lg >= 17 && env_int("B200_MSM_AFFINE", 0) != 0;
    // per-base scalar randomisation (msm.cu): digit distributions independent of the caller's scalars; the engine applies it
    // only when every base is in the prime-order subgroup
    cfg.randomize = fixed && env_int("B200_MSM_RANDOMIZE", 1) != 0;
    return cfg;
}

// one open batch of single mult_pippenger_prepared calls (coalesce.cuh); slots are n scalars wide, short calls are
// zero-padded (a zero scalar has no non-zero digit: it never reaches a bucket)
struct MsmCoBatch : CoBatchBase {
    uint8_t* h_in = nullptr;    // pinned: max_batch x n x 32
    uint8_t* h_out = nullptr;   // pinned: max_batch x 144
    std::string msg;
    ~MsmCoBatch() {
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
    }
};

struct MsmHandle {
    std::mutex mu;  // the handle is Send + Sync on the Rust side: serialise users of the engine and its workspace
    std::unique_ptr<MsmEngine> eng;
    size_t npoints = 0;
    int device = 0;                  // the device the table lives on; entry points switch to it (DeviceScope)
    cudaStream_t stream = nullptr;
    // end of the last enqueue of b200_msm_prepared_device on a CALLER's stream: the engine's workspace is shared, so the
    // next user orders its stream after it, whatever stream that is
    cudaEvent_t ev_busy = nullptr;
    bool busy = false;
    uint8_t* scalars_dev = nullptr;  // staging for host-pointer calls
    uint8_t* out_dev = nullptr;
    size_t scalars_cap = 0;
    int out_cap = 0;
    CoQueue<MsmCoBatch, 1> co;
    // small fixed-base handles (the 4096-point commitment MSM behind g1_lincomb): direct lookups in a table of every signed
    // digit multiple (fk20_direct.cu) instead of the bucket pipeline, whose reduction tail is ~0.45 ms however short the call
    void* direct_table = nullptr;
    int direct_c = 0;
    uint8_t* direct_part = nullptr;   // max_batch x 128 partial sums
    unsigned* direct_cnt = nullptr;   // max_batch completion counters (zero between launches)
    uint8_t* canon_dev = nullptr;     // canonical copy of the scalars (the engine takes blst_fr, the table digits do not)
    size_t canon_cap = 0;
    // sums of `batch` vectors of npoints blst_fr scalars (device, or host when scalars_host != nullptr: copied to
    // scalars_dev first) -> batch blst_p1 at out; caller holds mu and has called ensure_staging
    void run(const void* scalars_dev_in, size_t np, int batch, void* out, cudaStream_t st, const void* scalars_host = nullptr) {
        if (!direct_table || np != npoints) {
            eng->run(scalars_dev_in, np, batch, true, out, st, scalars_host);
            return;
        }
        const size_t total = (size_t)batch * np;
        if (scalars_host) B200_CUDA_CHECK(cudaMemcpyAsync(const_cast<void*>(scalars_dev_in), scalars_host, total * 32, cudaMemcpyHostToDevice, st));
        if (total > canon_cap) {
            cudaFree(canon_dev);
            canon_dev = nullptr; canon_cap = 0;
            canon_dev = dev_alloc<uint8_t>(total * 32);
            canon_cap = total;
        }
        launch_fr_from_mont(scalars_dev_in, canon_dev, total, st);
        launch_direct_msm(canon_dev, direct_table, direct_part, direct_cnt, nullptr, (uint8_t*)out, batch, (int)np, direct_c, st);
    }
    ~MsmHandle() {
        cudaFree(direct_table); cudaFree(direct_part); cudaFree(direct_cnt); cudaFree(canon_dev);
        cudaFree(scalars_dev);
        cudaFree(out_dev);
        if (ev_busy) cudaEventDestroy(ev_busy);
        if (stream) cudaStreamDestroy(stream);
    }
    void ensure_staging(size_t nscalars, int batch) {
        if (nscalars > scalars_cap) {
            cudaFree(scalars_dev);
            scalars_dev = nullptr; scalars_cap = 0;
            scalars_dev = dev_alloc<uint8_t>(nscalars * 32);
            scalars_cap = nscalars;
        }
        if (batch > out_cap) {
            cudaFree(out_dev);
            out_dev = nullptr; out_cap = 0;
            out_dev = dev_alloc<uint8_t>((size_t)batch * 144);
            out_cap = batch;
        }
    }
    // caller holds mu
    void enter(cudaStream_t user) {
        if (busy) B200_CUDA_CHECK(cudaStreamWaitEvent(user, ev_busy, 0));
    }
    void leave_async(cudaStream_t user) {
        B200_CUDA_CHECK(cudaEventRecord(ev_busy, user));
        busy = true;
    }
};

MsmHandle* msm_handle_create(const void* points, size_t npoints, bool host_points, bool fixed, int max_batch) {
    require_device();
    std::unique_ptr<MsmHandle> h(new MsmHandle());
    B200_CUDA_CHECK(cudaGetDevice(&h->device));
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_busy, cudaEventDisableTiming));
    MsmConfig cfg = choose_config(npoints, fixed, max_batch);
    h->eng.reset(new MsmEngine(cfg, points, host_points, h->stream));
    h->npoints = npoints;
    h->co.max_batches = 3;
    // direct-lookup table for small fixed-base handles whose items fill whole 128-thread slices (B200_MSM_DIRECT=0: never;
    // B200_MSM_DIRECT_BITS: window width, default 11 = 7.5 GiB for 4096 points; narrower or none when HBM is short)
    if (fixed && points && npoints >= 1024 && npoints <= 8192 && env_int("B200_MSM_DIRECT", 1)) {
        int c = pick_direct_bits(npoints, env_int("B200_MSM_DIRECT_BITS", 11));
        while (c && (npoints * direct_windows(c)) % 128) c = c > 8 ? 8 : 0;   // 8-bit windows: 32 per point, any npoints % 4 == 0
        if (c) {
            uint8_t* aff = nullptr;
            const void* src = points;
            if (host_points) {
                aff = dev_alloc<uint8_t>(npoints * 96);
                B200_CUDA_CHECK(cudaMemcpyAsync(aff, points, npoints * 96, cudaMemcpyHostToDevice, h->stream));
                src = aff;
            }
            h->direct_table = build_direct_table(src, npoints, 1, c, h->stream);
            cudaFree(aff);
            if (h->direct_table) {
                h->direct_c = c;
                h->direct_part = dev_alloc<uint8_t>((size_t)cfg.max_batch * 128 * 192);
                h->direct_cnt = dev_alloc<unsigned>(cfg.max_batch);
                B200_CUDA_CHECK(cudaMemsetAsync(h->direct_cnt, 0, cfg.max_batch * sizeof(unsigned), h->stream));
                B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
            }
        }
    }
    return h.release();
}

// scalar vectors one launch sequence may carry: small prepared MSMs (the 4096-point commitment MSM behind g1_lincomb) are
// latency-bound one at a time, so concurrent calls are packed; a 2^18-term MSM fills the machine on its own
static int default_max_batch(size_t npoints) {
    size_t b = npoints ? ((size_t)1 << 18) / npoints : 1;
    return (int)std::min<size_t>(64, std::max<size_t>(1, b));
}

}  // namespace b200

extern "C" {

void* prepare_msm(const blst_p1_affine points[], size_t npoints) {
    try {
        if (!points || npoints == 0) return nullptr;
        return msm_handle_create(points, npoints, true, true, env_int("B200_MSM_MAX_BATCH", default_max_batch(npoints)));
    } catch (const std::exception& e) {
        fprintf(stderr, "b200kzg: prepare_msm failed: %s\n", e.what());
        return nullptr;
    }
}

void b200_free_msm(void* msm) { delete static_cast<MsmHandle*>(msm); }

void b200_msm_plan(size_t npoints, int fixed, int* c, int* c0, int* W, int* fold_bits) {
    MsmConfig cfg = choose_config(npoints ? npoints : 1, fixed != 0, 1);
    const int w0 = fixed && cfg.c0 > 0 && cfg.c0 <= cfg.c ? cfg.c0 : cfg.c;
    if (c) *c = cfg.c;
    if (c0) *c0 = w0;
    if (W) *W = cfg.W;
    // what the 15-bit marginal reduce cannot take, or the configured fold of a narrow window (msm.cu: MsmEngine::kf_)
    if (fold_bits) *fold_bits = std::max(cfg.c > 16 ? cfg.c - 16 : 0, std::min(cfg.fold, cfg.c - 2));
}

RustError b200_msm_prepared_device(void* msm, void* out_dev, size_t npoints, const void* scalars_dev, int batch,
                                   void* stream) {
    return guarded([&] {
        MsmHandle* h = static_cast<MsmHandle*>(msm);
        if (!h) throw CudaError(-1, "null msm handle");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter((cudaStream_t)stream);
        h->run(scalars_dev, npoints, batch, out_dev, (cudaStream_t)stream);
        h->leave_async((cudaStream_t)stream);
    });
}

RustError b200_msm_prepared_batch(void* msm, blst_p1 out[], size_t npoints, const blst_fr scalars[], int batch) {
    return guarded([&] {
        MsmHandle* h = static_cast<MsmHandle*>(msm);
        if (!h) throw CudaError(-1, "null msm handle");
        if (npoints > h->npoints) throw CudaError(-1, "npoints exceeds the prepared table");
        if (batch < 1 || batch > h->eng->config().max_batch) throw CudaError(-1, "batch exceeds the prepared capacity");
        if (npoints == 0) {
            memset(out, 0, sizeof(blst_p1) * batch);
            return;
        }
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter(h->stream);
        h->ensure_staging((size_t)batch * npoints, batch);
        h->run(h->scalars_dev, npoints, batch, h->out_dev, h->stream, scalars);
        B200_CUDA_CHECK(cudaMemcpyAsync(out, h->out_dev, (size_t)batch * 144, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

// blst-sppark/cuda/pippenger.cu:28-31.  Handles prepared for more than one scalar vector per launch coalesce concurrent
// callers (the handle is Send + Sync on the Rust side, kzg/src/msm/sppark.rs:24-44); see coalesce.cuh.
RustError mult_pippenger_prepared(void* msm, blst_p1* out, size_t npoints, const blst_fr scalars[]) {
    MsmHandle* h = static_cast<MsmHandle*>(msm);
    if (!h || !out || (npoints && !scalars) || npoints > h->npoints || npoints == 0 || h->eng->config().max_batch < 2)
        return b200_msm_prepared_batch(msm, out, npoints, scalars, 1);
    return guarded([&] {
        const size_t n = h->npoints;
        const int cap = h->eng->config().max_batch;
        auto cl = h->co.claim(0, cap, [&] {
            std::unique_ptr<MsmCoBatch> nb(new MsmCoBatch());
            DeviceScope ds(h->device);
            B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_in, (size_t)cap * n * 32));
            B200_CUDA_CHECK(cudaMallocHost((void**)&nb->h_out, (size_t)cap * 144));
            return nb;
        });
        MsmCoBatch* B = cl.b;
        uint8_t* slot = B->h_in + (size_t)cl.idx * n * 32;
        memcpy(slot, scalars, npoints * 32);
        if (npoints < n) memset(slot + npoints * 32, 0, (n - npoints) * 32);
        h->co.staged(B);
        if (cl.leader) {
            int rc = 0;
            try {
                DeviceScope ds(h->device);
                std::lock_guard<std::mutex> lk(h->mu);   // waits while the previous batch runs: meanwhile this one fills
                const int m = h->co.close(B);
                h->enter(h->stream);
                h->ensure_staging((size_t)m * n, m);
                h->run(h->scalars_dev, n, m, h->out_dev, h->stream, B->h_in);
                B200_CUDA_CHECK(cudaMemcpyAsync(B->h_out, h->out_dev, (size_t)m * 144, cudaMemcpyDeviceToHost, h->stream));
                B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
            } catch (const CudaError& e) {
                cudaGetLastError();
                rc = e.code ? e.code : -1;
                B->msg = e.what();
            } catch (const std::exception& e) {
                rc = -1;
                B->msg = e.what();
            }
            h->co.publish(B, rc);
        } else {
            h->co.wait(B);
        }
        const int rc = B->rc;
        std::string msg = rc ? B->msg : std::string();
        if (!rc) memcpy(out, B->h_out + (size_t)cl.idx * 144, 144);
        h->co.consume(B);
        if (rc) throw CudaError(rc, msg);
    });
}

// variable-base engines are cached per capacity (next power of two) so repeated calls do not re-allocate; two per
// capacity, so that one caller's transfers overlap another caller's kernels
static std::mutex g_var_mu;
static constexpr int kVarEngines = 2;
struct VarSlot { MsmHandle* h[kVarEngines] = {}; unsigned next = 0; };
static std::map<std::pair<int, size_t>, VarSlot> g_var_engines;

RustError mult_pippenger(blst_p1* out, const blst_p1_affine points[], size_t npoints, const blst_fr scalars[]) {
    return guarded([&] {
        require_device();
        if (npoints == 0) {
            memset(out, 0, sizeof(blst_p1));
            return;
        }
        size_t cap = 256;
        while (cap < npoints) cap <<= 1;
        int dev = 0;
        B200_CUDA_CHECK(cudaGetDevice(&dev));   // no handle: the call runs on the calling thread's current device
        MsmHandle* h;
        {
            std::lock_guard<std::mutex> lk(g_var_mu);
            VarSlot& vs = g_var_engines[std::make_pair(dev, cap)];
            // large engines (a 2^20-point one holds ~1 GiB of workspace) are not duplicated
            const unsigned k = cap >= ((size_t)1 << 18) ? 0 : vs.next++ % kVarEngines;
            if (!vs.h[k]) vs.h[k] = msm_handle_create(nullptr, cap, true, false, 1);
            h = vs.h[k];
        }
        std::lock_guard<std::mutex> lk(h->mu);
        h->ensure_staging(npoints, 1);
        // bases go straight into the engine's table; scalars to staging
        B200_CUDA_CHECK(cudaMemcpyAsync(const_cast<void*>(h->eng->table()), points, npoints * 96, cudaMemcpyHostToDevice, h->stream));
        B200_CUDA_CHECK(cudaMemcpyAsync(h->scalars_dev, scalars, npoints * 32, cudaMemcpyHostToDevice, h->stream));
        h->eng->run(h->scalars_dev, npoints, 1, true, h->out_dev, h->stream);
        B200_CUDA_CHECK(cudaMemcpyAsync(out, h->out_dev, 144, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

/* window width of the handle's direct-lookup table (full-length calls take it instead of the bucket pipeline), 0 = none */
int b200_msm_direct_bits(void* msm) {
    MsmHandle* h = static_cast<MsmHandle*>(msm);
    return h && h->direct_table ? h->direct_c : 0;
}
/* 1 when the handle's table is scalar-randomised (all bases were in the prime-order subgroup at prepare), else 0 */
int b200_msm_randomized(void* msm) {
    MsmHandle* h = static_cast<MsmHandle*>(msm);
    return h && h->eng->randomized() ? 1 : 0;
}
/* 1 when the last run on this handle used the batch-affine accumulation (k_accumulate_affine), else 0 */
int b200_msm_last_affine(void* msm) {
    MsmHandle* h = static_cast<MsmHandle*>(msm);
    return h && h->eng->last_run_affine() ? 1 : 0;
}
void b200_msm_info(void* msm, int* c, int* W, size_t* table_bytes, int* launches) {
    MsmHandle* h = static_cast<MsmHandle*>(msm);
    if (!h) return;
    if (c) *c = h->eng->config().c;
    if (W) *W = h->eng->config().W;
    if (table_bytes) *table_bytes = h->eng->table_bytes();
    if (launches) *launches = h->eng->launches_per_run();
}

RustError b200_msm_last_counts(void* msm, size_t* entries, size_t* tasks) {
    return guarded([&] {
        MsmHandle* h = static_cast<MsmHandle*>(msm);
        if (!h) throw CudaError(-1, "null msm handle");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        B200_CUDA_CHECK(cudaDeviceSynchronize());  // the last run may have been on a caller's stream (device variant)
        h->eng->last_counts(entries, tasks, h->stream);
    });
}

RustError b200_msm_last_stats(void* msm, uint64_t stats[8]) {
    return guarded([&] {
        MsmHandle* h = static_cast<MsmHandle*>(msm);
        if (!h || !stats) throw CudaError(-1, "null argument");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        B200_CUDA_CHECK(cudaDeviceSynchronize());
        h->eng->last_stats(stats, h->stream);
    });
}

void b200_msm_set_profiling(void* msm, int on) {
    MsmHandle* h = static_cast<MsmHandle*>(msm);
    if (h) h->eng->set_profiling(on != 0);
}
RustError b200_msm_profile_read(void* msm, double* accumulate_ms_sum, int* runs) {
    return guarded([&] {
        MsmHandle* h = static_cast<MsmHandle*>(msm);
        if (!h) throw CudaError(-1, "null msm handle");
        DeviceScope ds(h->device);
        h->eng->profile_read(accumulate_ms_sum, runs);
    });
}
// out = sum of n Jacobian points (device pointers): the local add after the all-gather of per-GPU partial results
RustError b200_g1_sum_device(void* out_dev, const void* points_dev, size_t n, void* stream) {
    return guarded([&] {
        require_device();
        launch_g1_sum(points_dev, out_dev, (int)n, (cudaStream_t)stream);   // current device: the pointers are the caller's
    });
}

}  // extern "C"

// ---- multi-GPU MSM: one process per GPU, terms sharded by rank ---------------------------------
// Every rank prepares ITS slice of the bases and passes ITS slice of the scalars; the only exchange is an all-gather of the
// 144-byte partial results over NCCL (NVLink / NVSwitch) followed by a local add -- NCCL has no reduction over curve points,
// so the north star's "all-reduce of partial sums" is gather + add.  Everything is enqueued on one stream: the local MSM,
// the collective and the one-warp quad-tree sum.  NCCL is bound at run time (dlopen of the libnccl.so.2 the process
// already carries, e.g. torch's), so single-GPU users of this library do not need it.
namespace {
struct NcclId128 { char internal[128]; };   // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId128*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.GetUniqueId = (int (*)(NcclId128*))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void**, int, NcclId128, int))dlsym(api.lib, "ncclCommInitRank");
        api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllGather");
        api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
    });
    if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy)
        throw CudaError(-3, "b200kzg: NCCL (libnccl.so.2) is not available in this process -- the sharded MSM needs it");
    return api;
}
static void nccl_check(int rc, const char* what) {
    if (rc == 0) return;
    NcclApi& a = nccl();
    throw CudaError(-4, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
}
struct ShardedMsm {
    MsmHandle* h = nullptr;       // local engine over this rank's slice of the bases
    void* comm = nullptr;         // ncclComm_t
    int rank = 0, world = 1;
    uint8_t* partial = nullptr;   // 144 B
    uint8_t* gathered = nullptr;  // world x 144 B
    uint8_t* total = nullptr;     // 144 B (host-pointer variant)
    ~ShardedMsm() {
        if (comm) nccl().CommDestroy(comm);
        cudaFree(partial); cudaFree(gathered); cudaFree(total);
        delete h;
    }
    // caller holds h->mu and has switched to h->device
    void enqueue(const void* scalars_dev, const void* scalars_host, size_t n_local, void* out_dev, cudaStream_t st) {
        h->eng->run(scalars_dev, n_local, 1, true, world > 1 ? partial : (uint8_t*)out_dev, st, scalars_host);
        if (world > 1) {
            nccl_check(nccl().AllGather(partial, gathered, 144, /*ncclUint8*/ 1, comm, st), "ncclAllGather");
            launch_g1_sum(gathered, out_dev, world, st);
        }
    }
};

}  // namespace

extern "C" {

int b200_msm_sharded_unique_id(uint8_t id[128]) {
    try {
        NcclId128 u;
        nccl_check(nccl().GetUniqueId(&u), "ncclGetUniqueId");
        memcpy(id, u.internal, 128);
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "b200kzg: %s\n", e.what());
        return -1;
    }
}
void* b200_msm_sharded_prepare(const blst_p1_affine local_points[], size_t n_local, int rank, int world, const uint8_t id[128]) {
    try {
        if (!local_points || n_local == 0 || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return nullptr;
        std::unique_ptr<ShardedMsm> s(new ShardedMsm());
        s->rank = rank;
        s->world = world;
        s->h = msm_handle_create(local_points, n_local, true, true, 1);
        s->partial = dev_alloc<uint8_t>(144);
        s->gathered = dev_alloc<uint8_t>((size_t)world * 144);
        s->total = dev_alloc<uint8_t>(144);
        if (world > 1) {
            NcclId128 u;
            memcpy(u.internal, id, 128);
            nccl_check(nccl().CommInitRank(&s->comm, world, u, rank), "ncclCommInitRank");
        }
        return s.release();
    } catch (const std::exception& e) {
        fprintf(stderr, "b200kzg: b200_msm_sharded_prepare failed: %s\n", e.what());
        return nullptr;
    }
}
void b200_msm_sharded_free(void* sh) { delete static_cast<ShardedMsm*>(sh); }
void* b200_msm_sharded_local(void* sh) { return sh ? static_cast<ShardedMsm*>(sh)->h : nullptr; }

// out_dev: 144 B (device) on every rank = the sum over ALL ranks' terms; asynchronous on `stream`
RustError b200_msm_sharded_mult_device(void* sh, void* out_dev, size_t n_local, const void* scalars_dev, void* stream) {
    return guarded([&] {
        ShardedMsm* s = static_cast<ShardedMsm*>(sh);
        if (!s) throw CudaError(-1, "null sharded msm handle");
        if (n_local == 0 || n_local > s->h->npoints) throw CudaError(-1, "n_local exceeds the prepared slice");
        DeviceScope ds(s->h->device);
        std::lock_guard<std::mutex> lk(s->h->mu);
        s->h->enter((cudaStream_t)stream);
        s->enqueue(scalars_dev, nullptr, n_local, out_dev, (cudaStream_t)stream);
        s->h->leave_async((cudaStream_t)stream);
    });
}
// host pointers: this rank's n_local scalars in, the full result out on every rank
RustError b200_msm_sharded_mult(void* sh, blst_p1* out, size_t n_local, const blst_fr local_scalars[]) {
    return guarded([&] {
        ShardedMsm* s = static_cast<ShardedMsm*>(sh);
        if (!s || !out || !local_scalars) throw CudaError(-1, "null argument");
        if (n_local == 0 || n_local > s->h->npoints) throw CudaError(-1, "n_local exceeds the prepared slice");
        MsmHandle* h = s->h;
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter(h->stream);
        h->ensure_staging(n_local, 1);
        s->enqueue(h->scalars_dev, local_scalars, n_local, s->total, h->stream);
        B200_CUDA_CHECK(cudaMemcpyAsync(out, s->total, 144, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

/* microbenchmark: npairs independent batch-affine additions (K per thread, one inversion per 128 K) over a table of npoints
 * points; returns the kernel time through *ms (profiles/r02_affine.md) */
RustError b200_bench_affine_pairs(uint32_t npoints, size_t npairs, int K, double* ms) {
    return guarded([&] {
        require_device();
        if (ms) *ms = bench_affine_pairs(npoints, npairs, K, nullptr);
    });
}

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

}  // extern "C"
