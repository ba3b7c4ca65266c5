"""ctypes binding of libb200kzg.so (include/b200_kzg.h).  The library is the product; this is only the FFI stub a
Python caller needs.  Loading fails loudly if the CUDA library has not been built -- there is no fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200kzg.so")


class B200Error(RuntimeError):
    pass


class RustError(C.Structure):
    _fields_ = [("code", C.c_int), ("message", C.c_void_p)]


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def check(err: RustError):
    if err.code != 0:
        msg = C.cast(err.message, C.c_char_p).value.decode() if err.message else "unknown"
        if err.message:
            _libc.free(err.message)
        raise B200Error("b200kzg error %d: %s" % (err.code, msg))


def load():
    if not os.path.exists(LIB_PATH):
        raise B200Error("libb200kzg.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` or "
                        "`make -C rust-kzg_b200/csrc`); this backend has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
    sigs = {
        "prepare_msm": (vp, [vp, sz]),
        "mult_pippenger_prepared": (RustError, [vp, vp, sz, vp]),
        "mult_pippenger": (RustError, [vp, vp, sz, vp]),
        "b200_free_msm": (None, [vp]),
        "b200_msm_prepared_device": (RustError, [vp, vp, sz, vp, ci, vp]),
        "b200_msm_prepared_batch": (RustError, [vp, vp, sz, vp, ci]),
        "b200_msm_info": (None, [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(sz), C.POINTER(ci)]),
        "b200_msm_last_counts": (RustError, [vp, C.POINTER(sz), C.POINTER(sz)]),
        "b200_msm_last_stats": (RustError, [vp, vp]),
        "b200_msm_last_affine": (ci, [vp]),
        "b200_msm_randomized": (ci, [vp]),
        "b200_msm_direct_bits": (ci, [vp]),
        "b200_msm_plan": (None, [sz, ci, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]),
        "b200_selftest_fp": (RustError, [ci, vp, vp, vp, sz]),
        "b200_selftest_fr": (RustError, [ci, vp, vp, vp, sz]),
        "b200_selftest_p1_add": (RustError, [vp, vp, vp, sz, ci]),
        "b200_selftest_p1_compress": (RustError, [vp, vp, sz]),
        "b200_microbench_int": (RustError, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "b200_microbench_mul": (RustError, [ci, ci, C.POINTER(C.c_double)]),
        "b200_bench_affine_pairs": (RustError, [C.c_uint32, sz, ci, C.POINTER(C.c_double)]),
        "b200_device_count": (ci, []),
        "b200_msm_set_profiling": (None, [vp, ci]),
        "b200_msm_profile_read": (RustError, [vp, C.POINTER(C.c_double), C.POINTER(ci)]),
        "b200_g1_sum_device": (RustError, [vp, vp, sz, vp]),
        "b200_msm_sharded_unique_id": (ci, [vp]),
        "b200_msm_sharded_prepare": (vp, [vp, sz, ci, ci, vp]),
        "b200_msm_sharded_mult": (RustError, [vp, vp, sz, vp]),
        "b200_msm_sharded_mult_device": (RustError, [vp, vp, sz, vp, vp]),
        "b200_msm_sharded_local": (vp, [vp]),
        "b200_msm_sharded_free": (None, [vp]),
        "b200_fft_settings_new": (vp, [ci]),
        "b200_fft_settings_free": (None, [vp]),
        "b200_fft_settings_max_width": (sz, [vp]),
        "b200_fft_settings_roots": (RustError, [vp, ci, vp]),
        "b200_fft_fr": (RustError, [vp, vp, vp, sz, C.c_bool]),
        "b200_das_fft_extension": (RustError, [vp, vp, vp, sz]),
        "b200_fft_fr_device": (RustError, [vp, vp, vp, sz, ci, ci, vp]),
        "b200_das_fft_extension_device": (RustError, [vp, vp, vp, sz, ci, vp]),
        "b200_fft_launches": (ci, [vp]),
        "b200_fft_g1": (RustError, [vp, vp, vp, sz, C.c_bool]),
        "b200_fft_g1_device": (RustError, [vp, vp, vp, sz, ci, ci, vp]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(lib, name)          # AttributeError here = header / library mismatch: fail loudly
        f.restype, f.argtypes = res, args
    return lib
