// capi_ntt.cu -- C ABI for the FFTSettings / FFTFr / DASExtension surface (include/b200_kzg.h, NTT section).
#include <cstring>
#include <memory>
#include <mutex>

#include "../../include/b200_kzg.h"
#include "capi_common.cuh"
#include "ntt.cuh"
#include "util.cuh"

using namespace b200;

namespace b200 {
struct FftHandle {
    std::mutex mu;
    std::unique_ptr<FFTSettingsDev> fs;
    cudaStream_t stream = nullptr;
    int device = 0;   // the device the tables live on; entry points switch to it (DeviceScope)
    // end of the last *_device enqueue on a caller's stream (the transform scratch is shared): the next user's stream waits
    cudaEvent_t ev_busy = nullptr;
    bool busy = false;
    uint8_t *in_dev = nullptr, *out_dev = nullptr;
    size_t cap = 0;
    ~FftHandle() {
        cudaFree(in_dev);
        cudaFree(out_dev);
        if (ev_busy) cudaEventDestroy(ev_busy);
        if (stream) cudaStreamDestroy(stream);
    }
    void enter(cudaStream_t user) {   // caller holds mu
        if (busy) B200_CUDA_CHECK(cudaStreamWaitEvent(user, ev_busy, 0));
    }
    void leave_async(cudaStream_t user) {
        B200_CUDA_CHECK(cudaEventRecord(ev_busy, user));
        busy = true;
    }
    void ensure(size_t elems) {
        if (elems <= cap) return;
        cudaFree(in_dev);
        cudaFree(out_dev);
        in_dev = dev_alloc<uint8_t>(elems * 32);
        out_dev = dev_alloc<uint8_t>(elems * 32);
        cap = elems;
    }
};
}  // namespace b200

extern "C" {

void* b200_fft_settings_new(int scale) {
    try {
        require_device();
        std::unique_ptr<FftHandle> h(new FftHandle());
        B200_CUDA_CHECK(cudaGetDevice(&h->device));
        B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_busy, cudaEventDisableTiming));
        h->fs.reset(new FFTSettingsDev(scale, h->stream));
        return h.release();
    } catch (const std::exception& e) {
        fprintf(stderr, "b200kzg: fft_settings_new failed: %s\n", e.what());
        return nullptr;
    }
}
void b200_fft_settings_free(void* fs) { delete static_cast<FftHandle*>(fs); }

size_t b200_fft_settings_max_width(void* fs) { return fs ? static_cast<FftHandle*>(fs)->fs->max_width() : 0; }

RustError b200_fft_settings_roots(void* fs, int which, blst_fr* out) {
    return guarded([&] {
        FftHandle* h = static_cast<FftHandle*>(fs);
        if (!h) throw CudaError(-1, "null fft settings");
        DeviceScope ds(h->device);
        size_t w = h->fs->max_width();
        if (which == 1) {
            B200_CUDA_CHECK(cudaMemcpy(out, h->fs->brp_roots_dev(), w * 32, cudaMemcpyDeviceToHost));
        } else {
            B200_CUDA_CHECK(cudaMemcpy(out, h->fs->roots_dev(), (w + 1) * 32, cudaMemcpyDeviceToHost));
            if (which == 2)  // reverse_roots_of_unity = reverse(roots_of_unity)
                for (size_t i = 0, j = w; i < j; i++, j--) { blst_fr t = out[i]; out[i] = out[j]; out[j] = t; }
        }
    });
}

RustError b200_fft_fr_device(void* fs, void* out_dev, const void* in_dev, size_t n, int inverse, int batch, void* stream) {
    return guarded([&] {
        FftHandle* h = static_cast<FftHandle*>(fs);
        if (!h) throw CudaError(-1, "null fft settings");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter((cudaStream_t)stream);
        h->fs->fft_fr(in_dev, out_dev, n, inverse != 0, batch, (cudaStream_t)stream);
        h->leave_async((cudaStream_t)stream);
    });
}
RustError b200_das_fft_extension_device(void* fs, void* odds_dev, const void* evens_dev, size_t n, int batch, void* stream) {
    return guarded([&] {
        FftHandle* h = static_cast<FftHandle*>(fs);
        if (!h) throw CudaError(-1, "null fft settings");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter((cudaStream_t)stream);
        h->fs->das_fft_extension(evens_dev, odds_dev, n, batch, (cudaStream_t)stream);
        h->leave_async((cudaStream_t)stream);
    });
}

RustError b200_fft_fr(void* fs, blst_fr* out, const blst_fr* in, size_t n, bool inverse) {
    return guarded([&] {
        FftHandle* h = static_cast<FftHandle*>(fs);
        if (!h) throw CudaError(-1, "null fft settings");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter(h->stream);
        if (n > h->fs->max_width()) throw CudaError(1, "Supplied list is longer than the available max width");
        if (n == 0 || (n & (n - 1))) throw CudaError(1, "A list with power-of-two length expected");
        h->ensure(n);
        B200_CUDA_CHECK(cudaMemcpyAsync(h->in_dev, in, n * 32, cudaMemcpyHostToDevice, h->stream));
        h->fs->fft_fr(h->in_dev, h->out_dev, n, inverse, 1, h->stream);
        B200_CUDA_CHECK(cudaMemcpyAsync(out, h->out_dev, n * 32, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}
RustError b200_das_fft_extension(void* fs, blst_fr* odds, const blst_fr* evens, size_t n) {
    return guarded([&] {
        FftHandle* h = static_cast<FftHandle*>(fs);
        if (!h) throw CudaError(-1, "null fft settings");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter(h->stream);
        if (n == 0) throw CudaError(1, "A non-zero list ab expected");
        if (n & (n - 1)) throw CudaError(1, "A list with power-of-two length expected");
        if (n * 2 > h->fs->max_width()) throw CudaError(1, "Supplied list is longer than the available max width");
        h->ensure(n);
        B200_CUDA_CHECK(cudaMemcpyAsync(h->in_dev, evens, n * 32, cudaMemcpyHostToDevice, h->stream));
        h->fs->das_fft_extension(h->in_dev, h->out_dev, n, 1, h->stream);
        B200_CUDA_CHECK(cudaMemcpyAsync(odds, h->out_dev, n * 32, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}
RustError b200_fft_g1_device(void* fs, void* out_dev, const void* in_dev, size_t n, int inverse, int batch, void* stream) {
    return guarded([&] {
        FftHandle* h = static_cast<FftHandle*>(fs);
        if (!h) throw CudaError(-1, "null fft settings");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter((cudaStream_t)stream);
        h->fs->fft_g1(in_dev, out_dev, n, inverse != 0, batch, (cudaStream_t)stream);
        h->leave_async((cudaStream_t)stream);
    });
}
RustError b200_fft_g1(void* fs, blst_p1* out, const blst_p1* in, size_t n, bool inverse) {
    return guarded([&] {
        FftHandle* h = static_cast<FftHandle*>(fs);
        if (!h) throw CudaError(-1, "null fft settings");
        DeviceScope ds(h->device);
        std::lock_guard<std::mutex> lk(h->mu);
        h->enter(h->stream);
        if (n > h->fs->max_width()) throw CudaError(1, "Supplied list is longer than the available max width");
        if (n == 0 || (n & (n - 1))) throw CudaError(1, "A list with power-of-two length expected");
        h->ensure(n * 9);  // 2 x 144 B per point inside the 32-byte-element staging buffers
        B200_CUDA_CHECK(cudaMemcpyAsync(h->in_dev, in, n * 144, cudaMemcpyHostToDevice, h->stream));
        h->fs->fft_g1(h->in_dev, h->out_dev, n, inverse, 1, h->stream);
        B200_CUDA_CHECK(cudaMemcpyAsync(out, h->out_dev, n * 144, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}
int b200_fft_launches(void* fs) { return fs ? static_cast<FftHandle*>(fs)->fs->launches_last() : 0; }

}  // extern "C"
