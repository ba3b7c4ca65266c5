// capi_common.cuh -- glue shared by the extern "C" translation units.
#pragma once
#include <cstdio>
#include <cstring>
#include <exception>

#include "../../include/b200_kzg.h"
#include "util.cuh"

namespace b200 {

// Fail loudly when there is no usable CUDA device: this library has no CPU path.
inline void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw CudaError(e == cudaSuccess ? -2 : kCudaErrorBase + (int)e,
                        "b200kzg: no CUDA device available -- this backend has no CPU fallback");
}

// Run `f`, mapping C++ exceptions to sppark's by-value RustError (message strdup'd, caller frees).
template <class F>
inline RustError guarded(F&& f) {
    try {
        f();
        return RustError{0, nullptr};
    } catch (const CudaError& e) {
        cudaGetLastError();  // clear sticky-free errors
        return RustError{e.code ? e.code : -1, strdup(e.what())};
    } catch (const std::exception& e) {
        return RustError{-1, strdup(e.what())};
    } catch (...) {
        return RustError{-1, strdup("unknown error")};
    }
}

int env_int(const char* name, int dflt);

// Every entry point runs on the device its context / handle was created on, whatever device is current in the calling
// thread (settings objects and handles are used from arbitrary host threads; the reference's are Send + Sync).
struct DeviceScope {
    int prev = -1, dev;
    explicit DeviceScope(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) B200_CUDA_CHECK(cudaSetDevice(dev));
    }
    ~DeviceScope() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};


}  // namespace b200
