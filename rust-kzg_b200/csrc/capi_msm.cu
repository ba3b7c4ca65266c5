// capi_msm.cu -- the sppark-shaped MSM FFI (include/b200_kzg.h, section B1) on top of MsmEngine.
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

#include "../../include/b200_kzg.h"
#include "capi_common.cuh"
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
    // fixed, >= 2^20 points: c = 20, W = 13 (260 bits: the top window still holds 15 scalar bits, so its digits spread
    // over 2^15 buckets -- with c = 18, 19 or 21 the top window has 3 or 8 bits and a handful of buckets get 2^20 entries):
    // measured at 2^20 c = 16 / 18 / 19 / 20 / 21 -> 7.22 / 14.4 / 9.85 / 6.53 / 14.6 ms (scripts/msm_window_sweep.py)
    if (fixed) c = lg >= 20 ? 20 : lg >= 18 ? 16 : lg >= 14 ? lg - 2 : lg >= 11 ? 12 : lg >= 8 ? 10 : 8;
    else c = lg >= 20 ? 16 : lg >= 16 ? 14 : lg >= 12 ? 12 : lg >= 9 ? 10 : 8;
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
    return cfg;
}

struct MsmHandle {
    std::mutex mu;  // the handle is Send + Sync on the Rust side: serialise users
    std::unique_ptr<MsmEngine> eng;
    size_t npoints = 0;
    cudaStream_t stream = nullptr;
    uint8_t* scalars_dev = nullptr;  // staging for host-pointer calls
    uint8_t* out_dev = nullptr;
    size_t scalars_cap = 0;
    int out_cap = 0;
    ~MsmHandle() {
        cudaFree(scalars_dev);
        cudaFree(out_dev);
        if (stream) cudaStreamDestroy(stream);
    }
    void ensure_staging(size_t nscalars, int batch) {
        if (nscalars > scalars_cap) {
            cudaFree(scalars_dev);
            scalars_dev = dev_alloc<uint8_t>(nscalars * 32);
            scalars_cap = nscalars;
        }
        if (batch > out_cap) {
            cudaFree(out_dev);
            out_dev = dev_alloc<uint8_t>((size_t)batch * 144);
            out_cap = batch;
        }
    }
};

MsmHandle* msm_handle_create(const void* points, size_t npoints, bool host_points, bool fixed, int max_batch) {
    require_device();
    std::unique_ptr<MsmHandle> h(new MsmHandle());
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    MsmConfig cfg = choose_config(npoints, fixed, max_batch);
    h->eng.reset(new MsmEngine(cfg, points, host_points, h->stream));
    h->npoints = npoints;
    return h.release();
}

}  // namespace b200

extern "C" {

void* prepare_msm(const blst_p1_affine points[], size_t npoints) {
    try {
        if (!points || npoints == 0) return nullptr;
        return msm_handle_create(points, npoints, true, true, env_int("B200_MSM_MAX_BATCH", 1));
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
    if (fold_bits) *fold_bits = cfg.c > 16 ? cfg.c - 16 : 0;   // what the 15-bit marginal reduce cannot take (msm.cu)
}

RustError b200_msm_prepared_device(void* msm, void* out_dev, size_t npoints, const void* scalars_dev, int batch,
                                   void* stream) {
    return guarded([&] {
        MsmHandle* h = static_cast<MsmHandle*>(msm);
        if (!h) throw CudaError(-1, "null msm handle");
        std::lock_guard<std::mutex> lk(h->mu);
        h->eng->run(scalars_dev, npoints, batch, true, out_dev, (cudaStream_t)stream);
    });
}

RustError b200_msm_prepared_batch(void* msm, blst_p1 out[], size_t npoints, const blst_fr scalars[], int batch) {
    return guarded([&] {
        MsmHandle* h = static_cast<MsmHandle*>(msm);
        if (!h) throw CudaError(-1, "null msm handle");
        if (npoints > h->npoints) throw CudaError(-1, "npoints exceeds the prepared table");
        if (batch < 1 || batch > h->eng->config().max_batch) throw CudaError(-1, "batch exceeds the prepared capacity");
        std::lock_guard<std::mutex> lk(h->mu);
        if (npoints == 0) {
            memset(out, 0, sizeof(blst_p1) * batch);
            return;
        }
        h->ensure_staging((size_t)batch * npoints, batch);
        h->eng->run(h->scalars_dev, npoints, batch, true, h->out_dev, h->stream, scalars);
        B200_CUDA_CHECK(cudaMemcpyAsync(out, h->out_dev, (size_t)batch * 144, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

RustError mult_pippenger_prepared(void* msm, blst_p1* out, size_t npoints, const blst_fr scalars[]) {
    return b200_msm_prepared_batch(msm, out, npoints, scalars, 1);
}

// variable-base engines are cached per capacity (next power of two) so repeated calls do not re-allocate
static std::mutex g_var_mu;
static std::map<size_t, MsmHandle*> g_var_engines;

RustError mult_pippenger(blst_p1* out, const blst_p1_affine points[], size_t npoints, const blst_fr scalars[]) {
    return guarded([&] {
        require_device();
        if (npoints == 0) {
            memset(out, 0, sizeof(blst_p1));
            return;
        }
        size_t cap = 256;
        while (cap < npoints) cap <<= 1;
        MsmHandle* h;
        {
            std::lock_guard<std::mutex> lk(g_var_mu);
            auto it = g_var_engines.find(cap);
            if (it == g_var_engines.end()) {
                h = msm_handle_create(nullptr, cap, true, false, 1);
                g_var_engines[cap] = h;
            } else {
                h = it->second;
            }
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
        std::lock_guard<std::mutex> lk(h->mu);
        B200_CUDA_CHECK(cudaDeviceSynchronize());  // the last run may have been on a caller's stream (device variant)
        h->eng->last_counts(entries, tasks, h->stream);
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
        h->eng->profile_read(accumulate_ms_sum, runs);
    });
}
// out = sum of n Jacobian points (device pointers): the local add after the all-gather of per-GPU partial results
RustError b200_g1_sum_device(void* out_dev, const void* points_dev, size_t n, void* stream) {
    return guarded([&] {
        require_device();
        launch_g1_sum(points_dev, out_dev, (int)n, (cudaStream_t)stream);
    });
}

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

}  // extern "C"
