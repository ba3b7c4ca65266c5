"""Host-side mirror of FFTSettings / FFTFr / DASExtension (kzg/src/lib.rs:421-431, 465-481) over the C ABI.
Arrays are numpy uint64 (n,4): blst_fr Montgomery limbs."""
import ctypes as C
import numpy as np

from . import _lib

__all__ = ["FFTSettings"]


def _L():
    from . import lib
    return lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class FFTSettings:
    """FsFFTSettings (blst/src/types/fft_settings.rs:13-58) with the roots tables resident on the GPU."""

    def __init__(self, scale: int):
        self.h = _L().b200_fft_settings_new(scale)
        if not self.h:
            raise _lib.B200Error("FFTSettings::new(%d) failed (scale out of range or no CUDA device)" % scale)
        self.max_width = 1 << scale

    def get_max_width(self):
        return self.max_width

    def _roots(self, which, n):
        out = np.zeros((n, 4), np.uint64)
        _lib.check(_L().b200_fft_settings_roots(self.h, which, _p(out)))
        return out

    def get_roots_of_unity(self):
        return self._roots(0, self.max_width + 1)

    def get_brp_roots_of_unity(self):
        return self._roots(1, self.max_width)

    def get_reversed_roots_of_unity(self):
        return self._roots(2, self.max_width + 1)

    def fft_fr(self, data, inverse=False):
        data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(data)
        _lib.check(_L().b200_fft_fr(self.h, _p(out), _p(data), data.shape[0], bool(inverse)))
        return out

    def das_fft_extension(self, evens):
        evens = np.ascontiguousarray(evens, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(evens)
        _lib.check(_L().b200_das_fft_extension(self.h, _p(out), _p(evens), evens.shape[0]))
        return out

    def fft_g1(self, points, inverse=False):
        """FFTG1::fft_g1: (n,18) Jacobian points in and out"""
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 18)
        out = np.zeros_like(pts)
        _lib.check(_L().b200_fft_g1(self.h, _p(out), _p(pts), pts.shape[0], bool(inverse)))
        return out

    def fft_g1_device(self, out_ptr, in_ptr, n, inverse=False, batch=1, stream=0):
        _lib.check(_L().b200_fft_g1_device(self.h, C.c_void_p(out_ptr), C.c_void_p(in_ptr), n, int(inverse), batch,
                                           C.c_void_p(stream)))

    def fft_fr_device(self, out_ptr, in_ptr, n, inverse=False, batch=1, stream=0):
        _lib.check(_L().b200_fft_fr_device(self.h, C.c_void_p(out_ptr), C.c_void_p(in_ptr), n, int(inverse), batch,
                                           C.c_void_p(stream)))

    def das_fft_extension_device(self, out_ptr, in_ptr, n, batch=1, stream=0):
        _lib.check(_L().b200_das_fft_extension_device(self.h, C.c_void_p(out_ptr), C.c_void_p(in_ptr), n, batch,
                                                      C.c_void_p(stream)))

    def launches(self):
        return _L().b200_fft_launches(self.h)

    def close(self):
        if self.h:
            _L().b200_fft_settings_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
