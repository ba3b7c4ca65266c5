//! Rust host side of the B200 backend: the three functions the blst backend calls for its GPU MSM
//! (`blst-sppark/src/lib.rs:8-62` -- same names, same signatures, so `blst/src/kzg_proofs.rs:32-61` and
//! `blst/src/types/kzg_settings.rs:108-132` compile unchanged against this crate), plus safe wrappers for the NTT,
//! DAS and `fft_g1` entry points and the batched / verification extensions of `include/b200_kzg.h`.
//!
//! SOURCE ONLY -- never compiled in the build image (no Rust toolchain there).  Memory layouts: `blst_fr`,
//! `blst_p1`, `blst_p1_affine` are the `#[repr(C)]` types of the `blst` crate, which is what `FsFr` / `FsG1` /
//! `FsG1Affine` wrap transparently (`blst/src/types/fr.rs:18-20`, `g1.rs:28-30, 290-292`).
use blst::{blst_fr, blst_p1, blst_p1_affine};
use std::ffi::{c_char, c_int, c_void, CStr};

pub mod ffi;

/// `RustError` of the C ABI (sppark's `sppark::Error`, `arkworks3-sppark-wlc/sppark/util/rusterror.h:15-27`):
/// returned by value, `message` is `strdup`'d by the library and owned by the caller.
#[repr(C)]
pub struct RustError {
    pub code: c_int,
    pub message: *mut c_char,
}

impl From<RustError> for String {
    fn from(e: RustError) -> String {
        if e.message.is_null() {
            return format!("b200kzg error {}", e.code);
        }
        let s = unsafe { CStr::from_ptr(e.message) }.to_string_lossy().into_owned();
        unsafe { libc_free(e.message as *mut c_void) };
        s
    }
}

extern "C" {
    #[link_name = "free"]
    fn libc_free(p: *mut c_void);
}

fn check(e: RustError) -> Result<(), String> {
    if e.code == 0 {
        Ok(())
    } else {
        Err(String::from(e))
    }
}

// ---- the sppark-shaped trio (blst-sppark/src/lib.rs:8-62) ----------------------------------------------------------

pub fn prepare_multi_scalar_mult(points: &[blst_p1_affine]) -> *mut c_void {
    unsafe { ffi::prepare_msm(points.as_ptr(), points.len()) }
}

pub fn multi_scalar_mult_prepared(msm: *mut c_void, scalars: &[blst_fr]) -> blst_p1 {
    let mut ret = blst_p1::default();
    let err = unsafe { ffi::mult_pippenger_prepared(msm, &mut ret, scalars.len(), scalars.as_ptr()) };
    if err.code != 0 {
        panic!("{}", String::from(err)); // the reference panics here too (blst-sppark/src/lib.rs:33-36)
    }
    ret
}

pub fn multi_scalar_mult(points: &[blst_p1_affine], scalars: &[blst_fr]) -> blst_p1 {
    if points.len() != scalars.len() {
        panic!("length mismatch")
    }
    let mut ret = blst_p1::default();
    let err = unsafe { ffi::mult_pippenger(&mut ret, points.as_ptr(), points.len(), scalars.as_ptr()) };
    if err.code != 0 {
        panic!("{}", String::from(err));
    }
    ret
}

/// `g1_lincomb_batch` (`kzg/src/lib.rs:160-181`) over one prepared table: `batch` scalar vectors of `npoints` each,
/// one launch sequence.
pub fn multi_scalar_mult_prepared_batch(msm: *mut c_void, scalars: &[blst_fr], npoints: usize) -> Vec<blst_p1> {
    assert!(npoints > 0 && scalars.len() % npoints == 0);
    let batch = scalars.len() / npoints;
    let mut out = vec![blst_p1::default(); batch];
    let err = unsafe { ffi::b200_msm_prepared_batch(msm, out.as_mut_ptr(), npoints, scalars.as_ptr(), batch as c_int) };
    if err.code != 0 {
        panic!("{}", String::from(err));
    }
    out
}

/// Releases a handle returned by `prepare_multi_scalar_mult` (the sppark plug has no such call and leaks).
pub fn free_multi_scalar_mult(msm: *mut c_void) {
    unsafe { ffi::b200_free_msm(msm) }
}

// ---- NTT / DAS / fft_g1 (FFTFr, FFTG1, DASExtension: kzg/src/lib.rs:421-431) -------------------------------------

/// Device twin of `FsFFTSettings` (`blst/src/types/fft_settings.rs:13-58`): owns the roots-of-unity table in HBM.
pub struct DeviceFFTSettings {
    handle: *mut c_void,
    pub max_width: usize,
}

unsafe impl Send for DeviceFFTSettings {}
unsafe impl Sync for DeviceFFTSettings {} // the library serialises users of one handle

impl DeviceFFTSettings {
    pub fn new(scale: usize) -> Result<Self, String> {
        if scale > 31 {
            return Err(String::from("Max scale is too large")); // blst/src/types/fft_settings.rs:30-32
        }
        let handle = unsafe { ffi::b200_fft_settings_new(scale as c_int) };
        if handle.is_null() {
            return Err(String::from("b200kzg: could not create device FFT settings (no CUDA device?)"));
        }
        Ok(Self { handle, max_width: 1usize << scale })
    }

    /// `FFTFr::fft_fr` (`blst/src/fft_fr.rs:156-165`); the error strings are the reference's own (code 1).
    pub fn fft_fr(&self, data: &[blst_fr], inverse: bool) -> Result<Vec<blst_fr>, String> {
        let mut ret = vec![blst_fr::default(); data.len()];
        check(unsafe { ffi::b200_fft_fr(self.handle, ret.as_mut_ptr(), data.as_ptr(), data.len(), inverse) })?;
        Ok(ret)
    }

    /// `DASExtension::das_fft_extension` (`blst/src/data_availability_sampling.rs:78-100`)
    pub fn das_fft_extension(&self, evens: &[blst_fr]) -> Result<Vec<blst_fr>, String> {
        let mut odds = vec![blst_fr::default(); evens.len()];
        check(unsafe { ffi::b200_das_fft_extension(self.handle, odds.as_mut_ptr(), evens.as_ptr(), evens.len()) })?;
        Ok(odds)
    }

    /// `FFTG1::fft_g1` (`blst/src/fft_g1.rs:53-83`)
    pub fn fft_g1(&self, data: &[blst_p1], inverse: bool) -> Result<Vec<blst_p1>, String> {
        let mut ret = vec![blst_p1::default(); data.len()];
        check(unsafe { ffi::b200_fft_g1(self.handle, ret.as_mut_ptr(), data.as_ptr(), data.len(), inverse) })?;
        Ok(ret)
    }
}

impl Drop for DeviceFFTSettings {
    fn drop(&mut self) {
        unsafe { ffi::b200_fft_settings_free(self.handle) }
    }
}
