"""rust-kzg_b200 -- B200 (sm_100a) backend for the rust-kzg MSM / NTT hot path.

The product is the CUDA shared library `libb200kzg.so` (sources in csrc/, C ABI in include/b200_kzg.h).  This
package is the thin Python-side FFI used by the tests and the benchmark; it mirrors the reference's operator
names (g1_lincomb, fft_fr, das_fft_extension, blob_to_kzg_commitment, ...).  No CPU fallback exists: importing
works without a GPU (so the ABI can be inspected), every compute call raises if the library or a device is missing.
"""
from . import _lib
from ._lib import B200Error, LIB_PATH

_handle = None


def lib():
    global _handle
    if _handle is None:
        _handle = _lib.load()
    return _handle


def device_count() -> int:
    return lib().b200_device_count()


from .g1_lincomb import *  # noqa: E402,F401,F403
from .fft import *  # noqa: E402,F401,F403
from .eip4844 import *  # noqa: E402,F401,F403
from . import eip4844  # noqa: E402
from .sharded import *  # noqa: E402,F401,F403
