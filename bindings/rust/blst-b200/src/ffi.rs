//! Raw declarations of `include/b200_kzg.h` (sections B1 and the NTT interface; the c-kzg-4844 section B2 is bound by
//! the reference's own `kzg::eth::c_bindings` types and needs no Rust declarations -- language bindings link it).
use super::RustError;
use blst::{blst_fr, blst_p1, blst_p1_affine};
use std::ffi::{c_int, c_void};

extern "C" {
    // B1 -- blst-sppark/cuda/pippenger.cu:23-38
    pub fn prepare_msm(points: *const blst_p1_affine, npoints: usize) -> *mut c_void;
    pub fn mult_pippenger_prepared(msm: *mut c_void, out: *mut blst_p1, npoints: usize, scalars: *const blst_fr) -> RustError;
    pub fn mult_pippenger(out: *mut blst_p1, points: *const blst_p1_affine, npoints: usize, scalars: *const blst_fr) -> RustError;
    pub fn b200_free_msm(msm: *mut c_void);
    pub fn b200_msm_prepared_batch(msm: *mut c_void, out: *mut blst_p1, npoints: usize, scalars: *const blst_fr, batch: c_int) -> RustError;
    // NTT interface -- FFTFr / FFTG1 / DASExtension
    pub fn b200_fft_settings_new(scale: c_int) -> *mut c_void;
    pub fn b200_fft_settings_free(fs: *mut c_void);
    pub fn b200_fft_settings_max_width(fs: *mut c_void) -> usize;
    pub fn b200_fft_fr(fs: *mut c_void, out: *mut blst_fr, inp: *const blst_fr, n: usize, inverse: bool) -> RustError;
    pub fn b200_das_fft_extension(fs: *mut c_void, odds: *mut blst_fr, evens: *const blst_fr, n: usize) -> RustError;
    pub fn b200_fft_g1(fs: *mut c_void, out: *mut blst_p1, inp: *const blst_p1, n: usize, inverse: bool) -> RustError;
    pub fn b200_device_count() -> c_int;
}
