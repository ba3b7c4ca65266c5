// ntt.cuh -- radix-2 Fr NTT (FFTFr::fft_fr) and DAS extension on sm_100a.
//
// Contract (blst/src/fft_fr.rs:112-165): natural-order input, natural-order output,
//   out[i] = sum_j data[j] * w^(i*j),  w = roots_of_unity[max_width / n]  (inverse: w^-1 and a final * n^-1),
// for any power-of-two n <= max_width.  Field elements are canonical Montgomery residues, so any algorithm that
// computes these sums is bit-exact; the device uses a two-pass decomposition n = n1 * n2 (three passes above 2^22 points)
// with every sub-transform (<= 2^11 points) done in shared memory by one CTA:
//   pass 1  n1 column transforms of size n2 over the stride-n1 sub-sequences, times the twiddle w^(i2*j1)
//   pass 2  n2 row transforms of size n1 (contiguous rows), written transposed to natural order
// Algorithmic HBM traffic: 64 B/element/pass (32 B read + 32 B write); two passes above 2^11 points.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

class FFTSettingsDev {
public:
    // FsFFTSettings::new(scale) (blst/src/types/fft_settings.rs:28-58): roots_of_unity[0..=max_width] on the device
    FFTSettingsDev(int scale, cudaStream_t stream);
    ~FFTSettingsDev();
    FFTSettingsDev(const FFTSettingsDev&) = delete;

    size_t max_width() const { return max_width_; }
    int scale() const { return scale_; }
    const void* roots_dev() const { return roots_; }          // (max_width + 1) Fr, w^0 .. w^max_width
    const void* brp_roots_dev() const { return brp_roots_; }  // max_width Fr, bit-reversed order

    // batch independent transforms of n points each (contiguous).  in/out device pointers; out != in.
    // tmp: device scratch of batch*n Fr (only used when n > 2^11); nullptr -> internal scratch is (re)allocated.
    void fft_fr(const void* in_dev, void* out_dev, size_t n, bool inverse, int batch, cudaStream_t stream);
    // DASExtension::das_fft_extension (blst/src/data_availability_sampling.rs:78-100): odds from evens, n = len(evens)
    void das_fft_extension(const void* evens_dev, void* odds_dev, size_t n, int batch, cudaStream_t stream);
    // FFTG1::fft_g1 (blst/src/fft_g1.rs:53-83): batch transforms of n Jacobian points (blst_p1), device pointers
    // apply_scale = false leaves out the [n^-1] of the inverse transform (a caller that can scale the inputs' scalars
    // instead saves one scalar multiplication per point: the transform is linear)
    void fft_g1(const void* in_jac_dev, void* out_jac_dev, size_t n, bool inverse, int batch, cudaStream_t stream,
                bool apply_scale = true);
    // buffers and kernels of fft_g1 made ready for launches of up to max_total points (see fft_g1.cu)
    void prepare_g1(size_t max_total);
    // (2^log_n)^-1 as a Montgomery Fr, device memory
    const void* inv_pow2_dev(int log_n) const { return (const uint8_t*)roots_ + (max_width_ + 1) * 32 + 33 * 32 + (size_t)log_n * 32; }
    int launches_last() const { return launches_; }

private:
    void ensure_scratch(size_t elems);
    void run_passes(const void* in, void* out, size_t n, bool inverse, int batch, bool scale, size_t twist_unit,
                    cudaStream_t st);
    void transform(const void* in, void* out, int k, bool inverse, int batch, const uint8_t* scale_ptr, size_t twist_unit,
                   size_t in_bstride, size_t out_bstride, size_t out_mul, void* tmp, cudaStream_t st);
    int scale_;
    size_t max_width_;
    void* roots_ = nullptr;
    void* brp_roots_ = nullptr;
    void* scratch_ = nullptr;
    void* scratch2_ = nullptr;
    size_t scratch_elems_ = 0;
    void* scratch3_ = nullptr;       // transforms above 2^22 points: output of the extra column pass
    size_t scratch3_elems_ = 0;
    void* g1_work_ = nullptr;
    size_t g1_work_elems_ = 0;
    void* g1_tmp_ = nullptr;         // products of the fused double stages of fft_g1
    size_t g1_tmp_elems_ = 0;
    int launches_ = 0;
};

}  // namespace b200
