"""ctypes binding of oracle/libkzg_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY.

Importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).  Never the product.
Data crosses as numpy uint64 arrays in blst layouts: Fr = 4 limbs, Fp = 6, affine = 12, Jacobian = 18 (Montgomery).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkzg_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("kzg_oracle.c", "kzg_oracle_pairing.inc")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def _load():
    build()
    try:
        lib = C.CDLL(_SO)
    except OSError:
        build(force=True)
        lib = C.CDLL(_SO)
    lib.ko_init()
    vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
    sigs = {
        "ko_fr_from_bendian": (ci, [vp, vp]), "ko_fr_from_bendian_unchecked": (None, [vp, vp]),
        "ko_fr_to_bendian": (None, [vp, vp]), "ko_fr_to_scalar": (None, [vp, vp]), "ko_fr_from_u64_arr": (None, [vp, vp]),
        "ko_fr_mul_batch": (None, [vp, vp, vp, sz]), "ko_fr_add_batch": (None, [vp, vp, vp, sz]),
        "ko_fr_sub_batch": (None, [vp, vp, vp, sz]), "ko_fr_inv_batch": (None, [vp, vp, sz]),
        "ko_fp_mul_batch": (None, [vp, vp, vp, sz]), "ko_fp_add_batch": (None, [vp, vp, vp, sz]),
        "ko_fp_sub_batch": (None, [vp, vp, vp, sz]), "ko_fp_inv_batch": (None, [vp, vp, sz]),
        "ko_fp_from_canon": (None, [vp, vp]), "ko_fp_to_canon": (None, [vp, vp]),
        "ko_p1_mult": (None, [vp, vp, vp]), "ko_p1_add_or_double": (None, [vp, vp, vp]), "ko_p1_double": (None, [vp, vp]),
        "ko_p1_from_affine": (None, [vp, vp]), "ko_p1_to_affine": (None, [vp, vp]), "ko_p1_is_inf": (ci, [vp]),
        "ko_p1_is_equal": (ci, [vp, vp]), "ko_p1_in_g1": (ci, [vp]), "ko_p1s_to_affine": (None, [vp, vp, sz]),
        "ko_p1_compress": (None, [vp, vp]), "ko_p1_uncompress": (ci, [vp, vp]),
        "ko_pippenger_window_size": (sz, [sz]),
        "ko_g1_lincomb": (None, [vp, vp, vp, sz, ci]), "ko_msm_affine": (None, [vp, vp, vp, sz, ci]),
        "ko_msm_naive": (None, [vp, vp, vp, sz]),
        "ko_scale2_root_of_unity": (None, [vp, ci]),
        "ko_fft_settings_new": (vp, [ci]), "ko_fft_settings_free": (None, [vp]), "ko_fft_settings_roots": (vp, [vp, ci]),
        "ko_fft_fr": (ci, [vp, vp, vp, sz, ci, ci]), "ko_fft_fr_slow": (None, [vp, vp, vp, sz, ci]),
        "ko_das_fft_extension": (ci, [vp, vp, vp, sz]),
        "ko_fft_g1": (ci, [vp, vp, vp, sz, ci]), "ko_fft_g1_slow": (None, [vp, vp, vp, sz, ci]),
        "ko_sha256": (None, [vp, vp, sz]),
        "ko_load_trusted_setup_text": (vp, [C.c_char_p, sz]), "ko_settings_set_threads": (None, [vp, ci]),
        "ko_settings_g1_lagrange_brp": (vp, [vp]), "ko_settings_g1_monomial": (vp, [vp]), "ko_settings_fft": (vp, [vp]),
        "ko_free_trusted_setup": (None, [vp]),
        "ko_blob_to_kzg_commitment": (ci, [vp, vp, vp]), "ko_compute_challenge": (ci, [vp, vp, vp]),
        "ko_compute_quotient": (ci, [vp, vp, vp, vp, vp]),
        "ko_compute_kzg_proof": (ci, [vp, vp, vp, vp, vp]), "ko_compute_blob_kzg_proof": (ci, [vp, vp, vp, vp]),
        "ko_compute_cells": (ci, [vp, vp, vp]),
        "ko_compute_cells_and_kzg_proofs": (ci, [vp, vp, vp, vp]), "ko_settings_x_ext_fft_columns": (vp, [vp]),
        "ko_p2_uncompress": (ci, [vp, vp]), "ko_p2_generator": (None, [vp]), "ko_p2_affine_on_curve": (ci, [vp]),
        "ko_pairings_verify": (ci, [vp, vp, vp, vp]), "ko_settings_g2_monomial": (vp, [vp]),
        "ko_verify_kzg_proof": (ci, [vp, vp, vp, vp, vp, vp]), "ko_verify_blob_kzg_proof": (ci, [vp, vp, vp, vp, vp]),
        "ko_verify_blob_kzg_proof_batch": (ci, [vp, vp, vp, vp, sz, vp]),
        "ko_recover_cells_and_kzg_proofs": (ci, [vp, vp, vp, vp, sz, vp]),
        "ko_compute_verify_cell_kzg_proof_batch_challenge": (ci, [vp, vp, sz, vp, vp, vp, vp, sz]),
        "ko_verify_cell_kzg_proof_batch": (ci, [vp, vp, vp, vp, vp, sz, vp]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return lib


lib = _load()


class OracleError(ValueError):
    pass


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a, width=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if width is not None:
        a = a.reshape(-1, width)
    return a


def _bytes_arr(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


# ---- Fr ----
def fr_from_bytes(b: bytes):
    out = np.zeros(4, np.uint64)
    if lib.ko_fr_from_bendian(_p(out), _p(_bytes_arr(b))):
        raise OracleError("Invalid scalar")
    return out


def fr_from_bytes_unchecked(b: bytes):
    out = np.zeros(4, np.uint64)
    lib.ko_fr_from_bendian_unchecked(_p(out), _p(_bytes_arr(b)))
    return out


def fr_to_bytes(fr) -> bytes:
    out = np.zeros(32, np.uint8)
    lib.ko_fr_to_bendian(_p(out), _p(_u64(fr)))
    return out.tobytes()


def fr_from_ints(vals):
    """canonical Python ints -> (n,4) Montgomery limbs"""
    out = np.zeros((len(vals), 4), np.uint64)
    for i, v in enumerate(vals):
        c = np.array([(int(v) >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)], np.uint64)
        lib.ko_fr_from_u64_arr(_p(out[i]), _p(c))
    return out


def fr_to_ints(frs):
    frs = _u64(frs, 4)
    out = []
    c = np.zeros(4, np.uint64)
    for i in range(frs.shape[0]):
        lib.ko_fr_to_scalar(_p(c), _p(frs[i]))
        out.append(sum(int(c[k]) << (64 * k) for k in range(4)))
    return out


def _binop(fn, a, b, w):
    a, b = _u64(a, w), _u64(b, w)
    out = np.zeros_like(a)
    fn(_p(out), _p(a), _p(b), a.shape[0])
    return out


def fr_mul(a, b): return _binop(lib.ko_fr_mul_batch, a, b, 4)
def fr_add(a, b): return _binop(lib.ko_fr_add_batch, a, b, 4)
def fr_sub(a, b): return _binop(lib.ko_fr_sub_batch, a, b, 4)
def fp_mul(a, b): return _binop(lib.ko_fp_mul_batch, a, b, 6)
def fp_add(a, b): return _binop(lib.ko_fp_add_batch, a, b, 6)
def fp_sub(a, b): return _binop(lib.ko_fp_sub_batch, a, b, 6)


def fr_inv(a):
    a = _u64(a, 4)
    out = np.zeros_like(a)
    lib.ko_fr_inv_batch(_p(out), _p(a), a.shape[0])
    return out


def fp_inv(a):
    a = _u64(a, 6)
    out = np.zeros_like(a)
    lib.ko_fp_inv_batch(_p(out), _p(a), a.shape[0])
    return out


def fp_from_ints(vals):
    out = np.zeros((len(vals), 6), np.uint64)
    for i, v in enumerate(vals):
        c = np.array([(int(v) >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(6)], np.uint64)
        lib.ko_fp_from_canon(_p(out[i]), _p(c))
    return out


def fp_to_ints(fps):
    fps = _u64(fps, 6)
    out = []
    c = np.zeros(6, np.uint64)
    for i in range(fps.shape[0]):
        lib.ko_fp_to_canon(_p(c), _p(fps[i]))
        out.append(sum(int(c[k]) << (64 * k) for k in range(6)))
    return out


# ---- G1 ----
def p1_compress(p) -> bytes:
    out = np.zeros(48, np.uint8)
    lib.ko_p1_compress(_p(out), _p(_u64(p)))
    return out.tobytes()


def p1_uncompress(b: bytes):
    """-> Jacobian (18 limbs).  Raises OracleError like FsG1::from_bytes."""
    if len(b) != 48:
        raise OracleError("Invalid byte length")
    aff = np.zeros(12, np.uint64)
    if lib.ko_p1_uncompress(_p(aff), _p(_bytes_arr(b))):
        raise OracleError("Failed to uncompress")
    out = np.zeros(18, np.uint64)
    lib.ko_p1_from_affine(_p(out), _p(aff))
    return out


def p1_uncompress_affine(b: bytes):
    aff = np.zeros(12, np.uint64)
    if lib.ko_p1_uncompress(_p(aff), _p(_bytes_arr(b))):
        raise OracleError("Failed to uncompress")
    return aff


def p1_add(a, b):
    out = np.zeros(18, np.uint64)
    lib.ko_p1_add_or_double(_p(out), _p(_u64(a)), _p(_u64(b)))
    return out


def p1_double(a):
    out = np.zeros(18, np.uint64)
    lib.ko_p1_double(_p(out), _p(_u64(a)))
    return out


def p1_mult(p, fr):
    out = np.zeros(18, np.uint64)
    lib.ko_p1_mult(_p(out), _p(_u64(p)), _p(_u64(fr)))
    return out


def p1_in_g1(p) -> bool:
    return bool(lib.ko_p1_in_g1(_p(_u64(p))))


def p1_is_inf(p) -> bool:
    return bool(lib.ko_p1_is_inf(_p(_u64(p))))


def p1_is_equal(a, b) -> bool:
    return bool(lib.ko_p1_is_equal(_p(_u64(a)), _p(_u64(b))))


def p1s_to_affine(ps):
    ps = _u64(ps, 18)
    out = np.zeros((ps.shape[0], 12), np.uint64)
    lib.ko_p1s_to_affine(_p(out), _p(ps), ps.shape[0])
    return out


def p1_from_affine(a):
    out = np.zeros(18, np.uint64)
    lib.ko_p1_from_affine(_p(out), _p(_u64(a)))
    return out


def g1_lincomb(points, scalars, length=None, nthreads=1):
    """msm() with precomputation=None on Jacobian points + Montgomery scalars (kzg/src/msm/msm_impls.rs:114-148)."""
    points, scalars = _u64(points, 18), _u64(scalars, 4)
    if length is None:
        length = min(points.shape[0], scalars.shape[0])
    out = np.zeros(18, np.uint64)
    lib.ko_g1_lincomb(_p(out), _p(points), _p(scalars), length, nthreads)
    return out


def msm_affine(points, scalars, length=None, nthreads=1):
    points, scalars = _u64(points, 12), _u64(scalars, 4)
    if length is None:
        length = min(points.shape[0], scalars.shape[0])
    out = np.zeros(18, np.uint64)
    lib.ko_msm_affine(_p(out), _p(points), _p(scalars), length, nthreads)
    return out


def msm_naive(points, scalars):
    points, scalars = _u64(points, 18), _u64(scalars, 4)
    out = np.zeros(18, np.uint64)
    lib.ko_msm_naive(_p(out), _p(points), _p(scalars), min(points.shape[0], scalars.shape[0]))
    return out


def sha256(msg: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    m = _bytes_arr(msg) if len(msg) else np.zeros(1, np.uint8)
    lib.ko_sha256(_p(out), _p(m), len(msg))
    return out.tobytes()


# ---- FFT ----
class FFTSettings:
    """blst/src/types/fft_settings.rs:28-58"""

    def __init__(self, scale, _handle=None):
        self._own = _handle is None
        self.h = lib.ko_fft_settings_new(scale) if _handle is None else _handle
        if not self.h:
            raise OracleError("Scale is expected to be within root of unity matrix row size")
        self.max_width = 1 << scale

    def _roots(self, which, n):
        ptr = lib.ko_fft_settings_roots(self.h, which)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n, 4)).copy()

    @property
    def roots_of_unity(self): return self._roots(0, self.max_width + 1)
    @property
    def brp_roots_of_unity(self): return self._roots(1, self.max_width)
    @property
    def reverse_roots_of_unity(self): return self._roots(2, self.max_width + 1)

    def fft_fr(self, data, inverse=False, nthreads=1):
        data = _u64(data, 4)
        out = np.zeros_like(data)
        if lib.ko_fft_fr(self.h, _p(out), _p(data), data.shape[0], int(inverse), nthreads):
            raise OracleError("fft_fr: bad length")
        return out

    def fft_fr_slow(self, data, inverse=False):
        data = _u64(data, 4)
        out = np.zeros_like(data)
        lib.ko_fft_fr_slow(self.h, _p(out), _p(data), data.shape[0], int(inverse))
        return out

    def fft_g1(self, points, inverse=False):
        pts = _u64(points, 18)
        out = np.zeros_like(pts)
        if lib.ko_fft_g1(self.h, _p(out), _p(pts), pts.shape[0], int(inverse)):
            raise OracleError("fft_g1: bad length")
        return out

    def fft_g1_slow(self, points, inverse=False):
        pts = _u64(points, 18)
        out = np.zeros_like(pts)
        lib.ko_fft_g1_slow(self.h, _p(out), _p(pts), pts.shape[0], int(inverse))
        return out

    def das_fft_extension(self, evens):
        evens = _u64(evens, 4)
        out = np.zeros_like(evens)
        if lib.ko_das_fft_extension(self.h, _p(out), _p(evens), evens.shape[0]):
            raise OracleError("das_fft_extension: bad length")
        return out

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib.ko_fft_settings_free(self.h)
            self.h = None


def scale2_root_of_unity(scale):
    out = np.zeros(4, np.uint64)
    lib.ko_scale2_root_of_unity(_p(out), scale)
    return out


# ---- EIP-4844 ----
class KZGSettings:
    def __init__(self, text: str, nthreads=1):
        b = text.encode()
        self.h = lib.ko_load_trusted_setup_text(b, len(b))
        if not self.h:
            raise OracleError("Incorrect trusted setup format")
        lib.ko_settings_set_threads(self.h, nthreads)
        self.fs = FFTSettings(13, _handle=lib.ko_settings_fft(self.h))

    def set_threads(self, n):
        lib.ko_settings_set_threads(self.h, n)

    @property
    def g1_lagrange_brp(self):
        ptr = lib.ko_settings_g1_lagrange_brp(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(4096, 18)).copy()

    @property
    def g1_monomial(self):
        ptr = lib.ko_settings_g1_monomial(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(4096, 18)).copy()


def _blob(blob: bytes):
    if len(blob) != 131072:
        raise OracleError("Invalid blob: Invalid byte length")
    return _bytes_arr(blob)


def blob_to_kzg_commitment(blob: bytes, s: KZGSettings) -> bytes:
    out = np.zeros(48, np.uint8)
    if lib.ko_blob_to_kzg_commitment(_p(out), _p(_blob(blob)), s.h):
        raise OracleError("Invalid scalar")
    return out.tobytes()


def compute_challenge(blob: bytes, commitment: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    if lib.ko_compute_challenge(_p(out), _p(_blob(blob)), _p(_bytes_arr(commitment))):
        raise OracleError("bad input")
    return out.tobytes()


def compute_kzg_proof(blob: bytes, z: bytes, s: KZGSettings):
    if len(z) != 32:
        raise OracleError("Invalid byte length")
    proof, y = np.zeros(48, np.uint8), np.zeros(32, np.uint8)
    if lib.ko_compute_kzg_proof(_p(proof), _p(y), _p(_blob(blob)), _p(_bytes_arr(z)), s.h):
        raise OracleError("bad input")
    return proof.tobytes(), y.tobytes()


def compute_quotient(poly, z, s: KZGSettings):
    poly = _u64(poly, 4)
    q, y = np.zeros_like(poly), np.zeros(4, np.uint64)
    if lib.ko_compute_quotient(_p(q), _p(y), _p(poly), _p(_u64(z)), s.h):
        raise OracleError("bad input")
    return q, y


def compute_blob_kzg_proof(blob: bytes, commitment: bytes, s: KZGSettings) -> bytes:
    if len(commitment) != 48:
        raise OracleError("Invalid byte length")
    proof = np.zeros(48, np.uint8)
    if lib.ko_compute_blob_kzg_proof(_p(proof), _p(_blob(blob)), _p(_bytes_arr(commitment)), s.h):
        raise OracleError("bad input")
    return proof.tobytes()


def compute_cells(blob: bytes, s: KZGSettings):
    out = np.zeros(128 * 2048, np.uint8)
    if lib.ko_compute_cells(_p(out), _p(_blob(blob)), s.h):
        raise OracleError("Invalid scalar")
    b = out.tobytes()
    return [b[i * 2048:(i + 1) * 2048] for i in range(128)]


def compute_cells_and_kzg_proofs(blob: bytes, s: KZGSettings, want_cells=True):
    """-> (cells list or None, proofs list of 128 x 48 bytes)   (kzg/src/das.rs:244-292)"""
    cells = np.zeros(128 * 2048, np.uint8) if want_cells else None
    proofs = np.zeros(128 * 48, np.uint8)
    if lib.ko_compute_cells_and_kzg_proofs(_p(cells) if want_cells else None, _p(proofs), _p(_blob(blob)), s.h):
        raise OracleError("Invalid scalar")
    pb = proofs.tobytes()
    cb = cells.tobytes() if want_cells else None
    return ([cb[i * 2048:(i + 1) * 2048] for i in range(128)] if want_cells else None,
            [pb[i * 48:(i + 1) * 48] for i in range(128)])


def x_ext_fft_columns(s: KZGSettings):
    ptr = lib.ko_settings_x_ext_fft_columns(s.h)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(128, 64, 18)).copy()


# ---- pairings / verification (kzg/src/eip_4844.rs:328-435, 586-866; blst/src/kzg_proofs.rs:74-100) ----
def p2_uncompress(b: bytes):
    """96-byte compressed G2 -> affine (x.c0, x.c1, y.c0, y.c1) as 24 u64 limbs; infinity = all zero"""
    out = np.zeros(24, np.uint64)
    if len(b) != 96 or lib.ko_p2_uncompress(_p(out), _p(_bytes_arr(b))):
        raise OracleError("Failed to uncompress")
    return out


def p2_generator():
    out = np.zeros(24, np.uint64)
    lib.ko_p2_generator(_p(out))
    return out


def pairings_verify(a1, a2, b1, b2) -> bool:
    """e(a1, a2) == e(b1, b2); a1, b1 Jacobian G1 (18 limbs), a2, b2 affine G2 (24 limbs)"""
    return bool(lib.ko_pairings_verify(_p(_u64(a1)), _p(_u64(a2)), _p(_u64(b1)), _p(_u64(b2))))


def g2_monomial(s: KZGSettings):
    ptr = lib.ko_settings_g2_monomial(s.h)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(65, 24)).copy()


def verify_kzg_proof(commitment: bytes, z: bytes, y: bytes, proof: bytes, s: KZGSettings) -> bool:
    if len(commitment) != 48 or len(z) != 32 or len(y) != 32 or len(proof) != 48:
        raise OracleError("Invalid byte length")
    ok = C.c_int(0)
    if lib.ko_verify_kzg_proof(C.byref(ok), _p(_bytes_arr(commitment)), _p(_bytes_arr(z)), _p(_bytes_arr(y)),
                               _p(_bytes_arr(proof)), s.h):
        raise OracleError("bad input")
    return bool(ok.value)


def verify_blob_kzg_proof(blob: bytes, commitment: bytes, proof: bytes, s: KZGSettings) -> bool:
    if len(commitment) != 48 or len(proof) != 48:
        raise OracleError("Invalid byte length")
    ok = C.c_int(0)
    if lib.ko_verify_blob_kzg_proof(C.byref(ok), _p(_blob(blob)), _p(_bytes_arr(commitment)), _p(_bytes_arr(proof)), s.h):
        raise OracleError("bad input")
    return bool(ok.value)


def verify_blob_kzg_proof_batch(blobs, commitments, proofs, s: KZGSettings) -> bool:
    n = len(blobs)
    if len(commitments) != n or len(proofs) != n:
        raise OracleError("Invalid amount of arguments")
    if any(len(c) != 48 for c in commitments) or any(len(p) != 48 for p in proofs):
        raise OracleError("Invalid byte length")
    bl = np.concatenate([_blob(b) for b in blobs]) if n else np.zeros(1, np.uint8)
    cs = _bytes_arr(b"".join(commitments)) if n else np.zeros(1, np.uint8)
    ps = _bytes_arr(b"".join(proofs)) if n else np.zeros(1, np.uint8)
    ok = C.c_int(0)
    if lib.ko_verify_blob_kzg_proof_batch(C.byref(ok), _p(bl), _p(cs), _p(ps), n, s.h):
        raise OracleError("bad input")
    return bool(ok.value)


# ---- EIP-7594 recovery / cell verification (kzg/src/das.rs:101-207, 294-452) ----
def _cells_arr(cells):
    if any(len(c) != 2048 for c in cells):
        raise OracleError("Invalid byte length")
    return _bytes_arr(b"".join(cells)) if cells else np.zeros(1, np.uint8)


def _g1s_arr(items):
    if any(len(c) != 48 for c in items):
        raise OracleError("Invalid byte length")
    return _bytes_arr(b"".join(items)) if items else np.zeros(1, np.uint8)


def recover_cells_and_kzg_proofs(cell_indices, cells, s: KZGSettings, want_proofs=True):
    """-> (128 cells, 128 proofs or None)"""
    if len(cell_indices) != len(cells):
        raise OracleError("Cell indicies mismatch")
    n = len(cells)
    idx = np.asarray(list(cell_indices) or [0], dtype=np.uint64)
    out_c = np.zeros(128 * 2048, np.uint8)
    out_p = np.zeros(128 * 48, np.uint8)
    if lib.ko_recover_cells_and_kzg_proofs(_p(out_c), _p(out_p) if want_proofs else None, _p(idx), _p(_cells_arr(cells)), n, s.h):
        raise OracleError("bad input")
    cb, pb = out_c.tobytes(), out_p.tobytes()
    return ([cb[i * 2048:(i + 1) * 2048] for i in range(128)],
            [pb[i * 48:(i + 1) * 48] for i in range(128)] if want_proofs else None)


def compute_verify_cell_kzg_proof_batch_challenge(commitments, commitment_indices, cell_indices, cells, proofs) -> bytes:
    n = len(cells)
    if len(commitment_indices) != n or len(cell_indices) != n or len(proofs) != n:
        raise OracleError("Cell count mismatch")
    out = np.zeros(32, np.uint8)
    ci = np.asarray(list(commitment_indices) or [0], dtype=np.uint64)
    ki = np.asarray(list(cell_indices) or [0], dtype=np.uint64)
    if lib.ko_compute_verify_cell_kzg_proof_batch_challenge(_p(out), _p(_g1s_arr(commitments)), len(commitments), _p(ci), _p(ki),
                                                            _p(_cells_arr(cells)), _p(_g1s_arr(proofs)), n):
        raise OracleError("bad input")
    return out.tobytes()


def verify_cell_kzg_proof_batch(commitments, cell_indices, cells, proofs, s: KZGSettings) -> bool:
    n = len(cells)
    if len(commitments) != n or len(cell_indices) != n or len(proofs) != n:
        raise OracleError("count mismatch")
    ki = np.asarray(list(cell_indices) or [0], dtype=np.uint64)
    ok = C.c_int(0)
    if lib.ko_verify_cell_kzg_proof_batch(C.byref(ok), _p(_g1s_arr(commitments)), _p(ki), _p(_cells_arr(cells)),
                                          _p(_g1s_arr(proofs)), n, s.h):
        raise OracleError("bad input")
    return bool(ok.value)
