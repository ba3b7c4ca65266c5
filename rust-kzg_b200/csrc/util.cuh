// util.cuh -- error handling and small launch helpers shared by the host drivers.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace b200 {

// CUDA failures become C++ exceptions and are turned into error codes at the extern "C" edge -- the same
// shape as the reference's sppark plug (arkworks3-sppark-wlc/sppark/util/exception.cuh:9-35, rusterror.h:15-27).
struct CudaError : public std::runtime_error {
    int code;
    CudaError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// Error domains: code 1 is the reference's Err(String) -- invalid input, C_KZG_BADARGS at the c-kzg edge -- and is only
// ever thrown explicitly by validation code.  CUDA runtime failures carry kCudaErrorBase + cudaError_t, so that
// cudaErrorInvalidValue (= 1) can never be mistaken for a validation error.  Negative codes: internal misuse.
constexpr int kCudaErrorBase = 0x10000;
#define B200_CUDA_CHECK(expr)                                                                               \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            throw ::b200::CudaError(::b200::kCudaErrorBase + (int)_e,                                       \
                                    std::string(cudaGetErrorString(_e)) + " at " + __FILE__ + ":" +          \
                                        std::to_string(__LINE__) + " in " #expr);                           \
    } while (0)

#define B200_LAUNCH_CHECK() B200_CUDA_CHECK(cudaGetLastError())

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

template <class T>
static inline T* dev_alloc(size_t count) {
    void* p = nullptr;
    B200_CUDA_CHECK(cudaMalloc(&p, count * sizeof(T) ? count * sizeof(T) : 16));
    return reinterpret_cast<T*>(p);
}

}  // namespace b200
