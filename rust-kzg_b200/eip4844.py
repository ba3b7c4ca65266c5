"""Host-side mirror of the c-kzg-4844 commitment / proof functions (blst/src/eip_4844.rs:160-530) over the C ABI.
Same names and argument meaning as the reference's bindings; errors (C_KZG_BADARGS) raise KzgError like the
upstream Python binding raises on a non-OK return."""
import ctypes as C
import os

import numpy as np

from . import _lib

__all__ = ["KZGSettings", "KzgError", "BYTES_PER_BLOB", "load_trusted_setup_file", "default_trusted_setup_path",
           "compute_challenge", "bytes_to_kzg_commitment", "bytes_from_bls_field", "selftest_lincomb_quads"]

BYTES_PER_BLOB = 131072
C_KZG_OK, C_KZG_BADARGS, C_KZG_ERROR, C_KZG_MALLOC = 0, 1, 2, 3


class KzgError(ValueError):
    def __init__(self, code, what):
        super().__init__("%s failed: C_KZG_RET %d" % (what, code))
        self.code = code


class CKZGSettings(C.Structure):
    _fields_ = [("roots_of_unity", C.c_void_p), ("brp_roots_of_unity", C.c_void_p), ("reverse_roots_of_unity", C.c_void_p),
                ("g1_values_monomial", C.c_void_p), ("g1_values_lagrange_brp", C.c_void_p), ("g2_values_monomial", C.c_void_p),
                ("x_ext_fft_columns", C.c_void_p), ("tables", C.c_void_p), ("wbits", C.c_size_t), ("scratch_size", C.c_size_t)]


def _L():
    from . import lib
    L = lib()
    if not getattr(L, "_ckzg_sigs", False):
        vp, sz, ci, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
        S = C.POINTER(CKZGSettings)
        sigs = {
            "load_trusted_setup": (ci, [S, vp, u64, vp, u64, vp, u64, u64]),
            "load_trusted_setup_file": (ci, [S, vp]),
            "free_trusted_setup": (None, [S]),
            "blob_to_kzg_commitment": (ci, [vp, vp, S]),
            "compute_kzg_proof": (ci, [vp, vp, vp, vp, S]),
            "compute_blob_kzg_proof": (ci, [vp, vp, vp, S]),
            "b200_blob_to_kzg_commitment_batch": (ci, [vp, vp, sz, S]),
            "b200_compute_kzg_proof_batch": (ci, [vp, vp, vp, vp, sz, S]),
            "b200_compute_blob_kzg_proof_batch": (ci, [vp, vp, vp, sz, S]),
            "b200_blob_to_kzg_commitment_device": (ci, [vp, vp, sz, vp, S, vp]),
            "b200_compute_kzg_proof_device": (ci, [vp, vp, vp, vp, sz, ci, vp, S, vp]),
            "compute_cells_and_kzg_proofs": (ci, [vp, vp, vp, S]),
            "b200_compute_cells_batch": (ci, [vp, vp, sz, S]),
            "b200_compute_cell_proofs_batch": (ci, [vp, vp, sz, S]),
            "b200_compute_cells_and_kzg_proofs_batch": (ci, [vp, vp, vp, sz, S]),
            "b200_kzg_direct_tables": (None, [S, vp]),
            "b200_kzg_cells_coalesce_stats": (None, [S, vp]),
            "b200_kzg_verify_coalesce_stats": (None, [S, vp]),
            "verify_kzg_proof": (ci, [vp, vp, vp, vp, vp, S]),
            "verify_blob_kzg_proof": (ci, [vp, vp, vp, vp, S]),
            "verify_blob_kzg_proof_batch": (ci, [vp, vp, vp, vp, sz, S]),
            "b200_verify_kzg_proof_batch": (ci, [vp, vp, vp, vp, vp, sz, S]),
            "b200_selftest_pairings_verify": (ci, [vp, vp, ci, vp, ci, S]),
            "recover_cells_and_kzg_proofs": (ci, [vp, vp, vp, vp, u64, S]),
            "verify_cell_kzg_proof_batch": (ci, [vp, vp, vp, vp, vp, u64, S]),
            "compute_verify_cell_kzg_proof_batch_challenge": (ci, [vp, vp, u64, vp, vp, vp, vp, u64]),
            "compute_challenge": (None, [vp, vp, vp]),
            "bytes_to_kzg_commitment": (ci, [vp, vp]),
            "bytes_from_bls_field": (None, [vp, vp]),
            "b200_kzg_launches": (ci, [S]),
            "b200_kzg_max_batch": (ci, [S]),
            "b200_selftest_sha256": (None, [vp, vp, sz, ci]),
        }
        for name, (res, args) in sigs.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L._ckzg_sigs = True
    return L


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


def default_trusted_setup_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "trusted_setup.txt")


def _buf(b, size=None, what="argument"):
    b = bytes(b) if not isinstance(b, (bytes, bytearray, np.ndarray)) else b
    if isinstance(b, np.ndarray):
        arr = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    else:
        arr = np.frombuffer(b, dtype=np.uint8)
    if size is not None and arr.size != size:
        # the C ABI takes fixed-size arrays; a wrong length is the reference's "Invalid byte length" Err
        raise KzgError(C_KZG_BADARGS, "%s length %d != %d" % (what, arr.size, size))
    return arr


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def selftest_lincomb_quads(points_affine, scalars_mont):
    """test hook: sum k_i P_i through the lane-quad GLV scalar multiplication; (n,12) u64 affine, (n,4) u64 Montgomery"""
    pts = np.ascontiguousarray(points_affine, dtype=np.uint64).reshape(-1, 12)
    sc = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(18, np.uint64)
    L = _L()
    f = L.b200_selftest_lincomb_quads
    f.restype, f.argtypes = _lib.RustError, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    _lib.check(f(_p(out), _p(pts), _p(sc), pts.shape[0]))
    return out


# ---- helper exports of blst/src/eip_4844.rs:498-530 (no settings argument)
def compute_challenge(blob, commitment_p1):
    """blob bytes + commitment as blst_p1 (18 u64 limbs) -> challenge as a Montgomery blst_fr (4 u64 limbs)"""
    b = _buf(blob, BYTES_PER_BLOB, "blob")
    c = np.ascontiguousarray(commitment_p1, dtype=np.uint64).reshape(18)
    out = np.zeros(4, np.uint64)
    _L().compute_challenge(_p(out), _p(b), _p(c))
    return out


def bytes_to_kzg_commitment(b48):
    b = _buf(b48, 48, "commitment")
    out = np.zeros(18, np.uint64)
    rc = _L().bytes_to_kzg_commitment(_p(out), _p(b))
    if rc != C_KZG_OK:
        raise KzgError(rc, "bytes_to_kzg_commitment")
    return out


def bytes_from_bls_field(fr) -> bytes:
    f = np.ascontiguousarray(fr, dtype=np.uint64).reshape(4)
    out = np.zeros(32, np.uint8)
    _L().bytes_from_bls_field(_p(out), _p(f))
    return out.tobytes()


class KZGSettings:
    """CKZGSettings + the device context behind it."""

    def __init__(self):
        self.c = CKZGSettings()
        self.loaded = False

    @classmethod
    def load_trusted_setup_file(cls, path=None):
        self = cls()
        path = path or default_trusted_setup_path()
        fp = _libc.fopen(path.encode(), b"r")
        if not fp:
            raise FileNotFoundError(path)
        try:
            rc = _L().load_trusted_setup_file(C.byref(self.c), fp)
        finally:
            _libc.fclose(fp)
        if rc != C_KZG_OK:
            raise KzgError(rc, "load_trusted_setup_file")
        self.loaded = True
        return self

    @classmethod
    def load_trusted_setup(cls, g1_monomial_bytes, g1_lagrange_bytes, g2_monomial_bytes, precompute=0):
        self = cls()
        a, b, c = _buf(g1_monomial_bytes), _buf(g1_lagrange_bytes), _buf(g2_monomial_bytes)
        rc = _L().load_trusted_setup(C.byref(self.c), _p(a), a.size, _p(b), b.size, _p(c), c.size, precompute)
        if rc != C_KZG_OK:
            raise KzgError(rc, "load_trusted_setup")
        self.loaded = True
        return self

    def free(self):
        if self.loaded:
            _L().free_trusted_setup(C.byref(self.c))
            self.loaded = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # host arrays as numpy copies (tests)
    def array(self, name, count, width):
        ptr = getattr(self.c, name)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(count, width)).copy()

    def x_ext_fft_columns(self):
        """the [128][64] blst_p1 of CKZGSettings.x_ext_fft_columns, read through the 128 row pointers exactly as the
        reference's TryFrom<&CKZGSettings> does (blst/src/types/kzg_settings.rs:398-417)"""
        rows = C.cast(self.c.x_ext_fft_columns, C.POINTER(C.c_void_p))
        out = np.zeros((128, 64, 18), np.uint64)
        for r in range(128):
            out[r] = np.ctypeslib.as_array(C.cast(rows[r], C.POINTER(C.c_uint64)), shape=(64, 18))
        return out

    @property
    def max_batch(self):
        return _L().b200_kzg_max_batch(C.byref(self.c))

    def launches(self):
        return _L().b200_kzg_launches(C.byref(self.c))

    # ---- c-kzg-4844 functions
    def blob_to_kzg_commitment(self, blob) -> bytes:
        b = _buf(blob, BYTES_PER_BLOB, "blob")
        out = np.zeros(48, np.uint8)
        rc = _L().blob_to_kzg_commitment(_p(out), _p(b), C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "blob_to_kzg_commitment")
        return out.tobytes()

    def compute_kzg_proof(self, blob, z):
        b, zb = _buf(blob, BYTES_PER_BLOB, "blob"), _buf(z, 32, "z")
        proof, y = np.zeros(48, np.uint8), np.zeros(32, np.uint8)
        rc = _L().compute_kzg_proof(_p(proof), _p(y), _p(b), _p(zb), C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_kzg_proof")
        return proof.tobytes(), y.tobytes()

    def compute_blob_kzg_proof(self, blob, commitment) -> bytes:
        b, cb = _buf(blob, BYTES_PER_BLOB, "blob"), _buf(commitment, 48, "commitment")
        proof = np.zeros(48, np.uint8)
        rc = _L().compute_blob_kzg_proof(_p(proof), _p(b), _p(cb), C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_blob_kzg_proof")
        return proof.tobytes()

    def compute_cells(self, blob):
        """compute_cells_and_kzg_proofs(cells, NULL, blob): -> list of 128 cells (2048 bytes each)"""
        b = _buf(blob, BYTES_PER_BLOB, "blob")
        out = np.zeros(128 * 2048, np.uint8)
        rc = _L().compute_cells_and_kzg_proofs(_p(out), None, _p(b), C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_cells_and_kzg_proofs")
        raw = out.tobytes()
        return [raw[i * 2048:(i + 1) * 2048] for i in range(128)]

    def compute_cells_and_kzg_proofs(self, blob):
        """-> (128 cells of 2048 bytes, 128 proofs of 48 bytes)"""
        b = _buf(blob, BYTES_PER_BLOB, "blob")
        cells, proofs = np.zeros(128 * 2048, np.uint8), np.zeros(128 * 48, np.uint8)
        rc = _L().compute_cells_and_kzg_proofs(_p(cells), _p(proofs), _p(b), C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_cells_and_kzg_proofs")
        cb, pb = cells.tobytes(), proofs.tobytes()
        return [cb[i * 2048:(i + 1) * 2048] for i in range(128)], [pb[i * 48:(i + 1) * 48] for i in range(128)]

    def compute_cell_proofs_batch(self, blobs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, BYTES_PER_BLOB)
        n = blobs.shape[0]
        out = np.zeros((n, 128, 48), np.uint8)
        rc = _L().b200_compute_cell_proofs_batch(_p(out), _p(blobs), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_cell_proofs_batch")
        return out

    def verify_coalesce_stats(self):
        """-> (batches checked, requests served, batches re-checked one by one) of the coalesced single verifications"""
        out = (C.c_uint64 * 3)()
        _L().b200_kzg_verify_coalesce_stats(C.byref(self.c), out)
        return int(out[0]), int(out[1]), int(out[2])

    def cells_coalesce_stats(self):
        """-> (batches run, requests served) of the coalesced compute_cells_and_kzg_proofs calls"""
        out = (C.c_uint64 * 2)()
        _L().b200_kzg_cells_coalesce_stats(C.byref(self.c), out)
        return int(out[0]), int(out[1])

    def direct_tables(self):
        """-> {"blob_bits", "blob_max_batch", "fk20_bits"}: window widths of the direct-lookup tables (0 = bucket engine)"""
        out = (C.c_int * 3)()
        _L().b200_kzg_direct_tables(C.byref(self.c), out)
        return {"blob_bits": out[0], "blob_max_batch": out[1], "fk20_bits": out[2]}

    def compute_cells_and_kzg_proofs_batch(self, blobs, cells_out=None, proofs_out=None):
        """both outputs for n blobs in one pass -> (n,128,2048) cells, (n,128,48) proofs; caller-owned arrays are filled in place"""
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, BYTES_PER_BLOB)
        n = blobs.shape[0]
        cells = np.zeros((n, 128, 2048), np.uint8) if cells_out is None else cells_out
        proofs = np.zeros((n, 128, 48), np.uint8) if proofs_out is None else proofs_out
        assert cells.flags.c_contiguous and proofs.flags.c_contiguous and cells.nbytes == n * 128 * 2048 and proofs.nbytes == n * 128 * 48
        rc = _L().b200_compute_cells_and_kzg_proofs_batch(_p(cells), _p(proofs), _p(blobs), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_cells_and_kzg_proofs_batch")
        return cells, proofs

    def compute_cells_batch(self, blobs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, BYTES_PER_BLOB)
        n = blobs.shape[0]
        out = np.zeros((n, 128, 2048), np.uint8)
        rc = _L().b200_compute_cells_batch(_p(out), _p(blobs), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_cells_batch")
        return out

    # ---- verification (blst/src/eip_4844.rs:383-471)
    def verify_kzg_proof(self, commitment, z, y, proof) -> bool:
        cb, zb, yb, pb = _buf(commitment, 48, "commitment"), _buf(z, 32, "z"), _buf(y, 32, "y"), _buf(proof, 48, "proof")
        ok = C.c_bool(False)
        rc = _L().verify_kzg_proof(C.byref(ok), _p(cb), _p(zb), _p(yb), _p(pb), C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "verify_kzg_proof")
        return bool(ok.value)

    def verify_blob_kzg_proof(self, blob, commitment, proof) -> bool:
        b, cb, pb = _buf(blob, BYTES_PER_BLOB, "blob"), _buf(commitment, 48, "commitment"), _buf(proof, 48, "proof")
        ok = C.c_bool(False)
        rc = _L().verify_blob_kzg_proof(C.byref(ok), _p(b), _p(cb), _p(pb), C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "verify_blob_kzg_proof")
        return bool(ok.value)

    def verify_blob_kzg_proof_batch(self, blobs, commitments, proofs) -> bool:
        """blobs: list of bytes / (n,131072) array; commitments, proofs: lists of 48-byte strings / (n,48) arrays"""
        if not isinstance(blobs, np.ndarray):
            if len(blobs) != len(commitments) or len(blobs) != len(proofs):
                raise KzgError(C_KZG_BADARGS, "Invalid amount of arguments")   # kzg/src/eip_4844.rs:770-772
            blobs = [_buf(b, BYTES_PER_BLOB, "blob") for b in blobs]
            commitments = [_buf(c, 48, "commitment") for c in commitments]
            proofs = [_buf(p_, 48, "proof") for p_ in proofs]
            blobs = np.concatenate(blobs) if blobs else np.zeros(0, np.uint8)
            commitments = np.concatenate(commitments) if commitments else np.zeros(0, np.uint8)
            proofs = np.concatenate(proofs) if proofs else np.zeros(0, np.uint8)
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, BYTES_PER_BLOB)
        commitments = np.ascontiguousarray(commitments, dtype=np.uint8).reshape(-1, 48)
        proofs = np.ascontiguousarray(proofs, dtype=np.uint8).reshape(-1, 48)
        n = blobs.shape[0]
        if commitments.shape[0] != n or proofs.shape[0] != n:
            raise KzgError(C_KZG_BADARGS, "Invalid amount of arguments")
        ok = C.c_bool(False)
        rc = _L().verify_blob_kzg_proof_batch(C.byref(ok), _p(blobs), _p(commitments), _p(proofs), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "verify_blob_kzg_proof_batch")
        return bool(ok.value)

    def verify_kzg_proof_batch(self, commitments, zs, ys, proofs) -> bool:
        commitments = np.ascontiguousarray(commitments, dtype=np.uint8).reshape(-1, 48)
        zs = np.ascontiguousarray(zs, dtype=np.uint8).reshape(-1, 32)
        ys = np.ascontiguousarray(ys, dtype=np.uint8).reshape(-1, 32)
        proofs = np.ascontiguousarray(proofs, dtype=np.uint8).reshape(-1, 48)
        n = commitments.shape[0]
        ok = C.c_bool(False)
        rc = _L().b200_verify_kzg_proof_batch(C.byref(ok), _p(commitments), _p(zs), _p(ys), _p(proofs), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "verify_kzg_proof_batch")
        return bool(ok.value)

    def pairings_verify(self, a1, qa, b1, qb) -> bool:
        """test hook: e(a1, Q[qa]) == e(b1, Q[qb]); a1, b1 = blst_p1 as 18 u64 limbs; Q = ([1]G2, [s]G2, [s^64]G2)"""
        a = np.ascontiguousarray(a1, dtype=np.uint64).reshape(18)
        b = np.ascontiguousarray(b1, dtype=np.uint64).reshape(18)
        ok = C.c_bool(False)
        rc = _L().b200_selftest_pairings_verify(C.byref(ok), _p(a), qa, _p(b), qb, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "pairings_verify")
        return bool(ok.value)

    # ---- EIP-7594 recovery / cell verification (kzg/src/eth/c_bindings.rs:201-352)
    @staticmethod
    def _cells(cells):
        arrs = [_buf(c, 2048, "cell") for c in cells]
        return np.concatenate(arrs) if arrs else np.zeros(1, np.uint8)

    @staticmethod
    def _g1s(items, what):
        arrs = [_buf(c, 48, what) for c in items]
        return np.concatenate(arrs) if arrs else np.zeros(1, np.uint8)

    @staticmethod
    def _u64s(vals):
        for v in vals:
            if not 0 <= int(v) < 1 << 64:
                raise KzgError(C_KZG_BADARGS, "index out of range")
        return np.asarray([int(v) for v in vals] or [0], dtype=np.uint64)

    def recover_cells_and_kzg_proofs(self, cell_indices, cells, want_proofs=True):
        """-> (128 cells of 2048 bytes, 128 proofs of 48 bytes or None)"""
        if len(cell_indices) != len(cells):
            raise KzgError(C_KZG_BADARGS, "Cell indicies mismatch")      # kzg/src/das.rs:119-124
        n = len(cells)
        idx, cb = self._u64s(cell_indices), self._cells(cells)
        out_c, out_p = np.zeros(128 * 2048, np.uint8), np.zeros(128 * 48, np.uint8)
        rc = _L().recover_cells_and_kzg_proofs(_p(out_c), _p(out_p) if want_proofs else None, _p(idx), _p(cb), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "recover_cells_and_kzg_proofs")
        craw, praw = out_c.tobytes(), out_p.tobytes()
        return ([craw[i * 2048:(i + 1) * 2048] for i in range(128)],
                [praw[i * 48:(i + 1) * 48] for i in range(128)] if want_proofs else None)

    def verify_cell_kzg_proof_batch(self, commitments, cell_indices, cells, proofs) -> bool:
        n = len(cells)
        if len(commitments) != n or len(cell_indices) != n or len(proofs) != n:
            raise KzgError(C_KZG_BADARGS, "count mismatch")               # kzg/src/das.rs:306-317
        cb, idx = self._g1s(commitments, "commitment"), self._u64s(cell_indices)
        cl, pb = self._cells(cells), self._g1s(proofs, "proof")
        ok = C.c_bool(False)
        rc = _L().verify_cell_kzg_proof_batch(C.byref(ok), _p(cb), _p(idx), _p(cl), _p(pb), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "verify_cell_kzg_proof_batch")
        return bool(ok.value)

    def compute_verify_cell_kzg_proof_batch_challenge(self, commitments, commitment_indices, cell_indices, cells, proofs):
        """-> the challenge as a Montgomery blst_fr (4 u64 limbs)"""
        n = len(cells)
        if len(commitment_indices) != n or len(cell_indices) != n or len(proofs) != n:
            raise KzgError(C_KZG_BADARGS, "Cell count mismatch")          # kzg/src/das.rs:401-406
        cb, ci, ki = self._g1s(commitments, "commitment"), self._u64s(commitment_indices), self._u64s(cell_indices)
        cl, pb = self._cells(cells), self._g1s(proofs, "proof")
        out = np.zeros(4, np.uint64)
        rc = _L().compute_verify_cell_kzg_proof_batch_challenge(_p(out), _p(cb), len(commitments), _p(ci), _p(ki), _p(cl), _p(pb), n)
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_verify_cell_kzg_proof_batch_challenge")
        return out

    # ---- batched extensions: blobs (n,131072) u8, returns (n,48) u8 ...
    def blob_to_kzg_commitment_batch(self, blobs, out=None):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, BYTES_PER_BLOB)
        n = blobs.shape[0]
        out = np.zeros((n, 48), np.uint8) if out is None else out
        rc = _L().b200_blob_to_kzg_commitment_batch(_p(out), _p(blobs), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "blob_to_kzg_commitment_batch")
        return out

    def compute_kzg_proof_batch(self, blobs, zs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, BYTES_PER_BLOB)
        zs = np.ascontiguousarray(zs, dtype=np.uint8).reshape(-1, 32)
        n = blobs.shape[0]
        proofs, ys = np.zeros((n, 48), np.uint8), np.zeros((n, 32), np.uint8)
        rc = _L().b200_compute_kzg_proof_batch(_p(proofs), _p(ys), _p(blobs), _p(zs), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_kzg_proof_batch")
        return proofs, ys

    def compute_blob_kzg_proof_batch(self, blobs, commitments, out=None):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, BYTES_PER_BLOB)
        commitments = np.ascontiguousarray(commitments, dtype=np.uint8).reshape(-1, 48)
        n = blobs.shape[0]
        out = np.zeros((n, 48), np.uint8) if out is None else out
        rc = _L().b200_compute_blob_kzg_proof_batch(_p(out), _p(blobs), _p(commitments), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_blob_kzg_proof_batch")
        return out

    # raw-pointer variants (host pinned or device memory owned by the caller)
    def blob_to_kzg_commitment_batch_ptr(self, out_ptr, blobs_ptr, n):
        rc = _L().b200_blob_to_kzg_commitment_batch(C.c_void_p(out_ptr), C.c_void_p(blobs_ptr), n, C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "blob_to_kzg_commitment_batch")

    def compute_blob_kzg_proof_batch_ptr(self, out_ptr, blobs_ptr, commitments_ptr, n):
        rc = _L().b200_compute_blob_kzg_proof_batch(C.c_void_p(out_ptr), C.c_void_p(blobs_ptr), C.c_void_p(commitments_ptr), n,
                                                    C.byref(self.c))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_blob_kzg_proof_batch")

    def blob_to_kzg_commitment_device(self, out_ptr, blobs_ptr, n, status_ptr, stream=0):
        rc = _L().b200_blob_to_kzg_commitment_device(C.c_void_p(out_ptr), C.c_void_p(blobs_ptr), n, C.c_void_p(status_ptr),
                                                     C.byref(self.c), C.c_void_p(stream))
        if rc != C_KZG_OK:
            raise KzgError(rc, "blob_to_kzg_commitment_device")

    def compute_kzg_proof_device(self, proofs_ptr, y_ptr, blobs_ptr, z_ptr, n, status_ptr, z_reduce=0, stream=0):
        rc = _L().b200_compute_kzg_proof_device(C.c_void_p(proofs_ptr), C.c_void_p(y_ptr), C.c_void_p(blobs_ptr),
                                                C.c_void_p(z_ptr), n, z_reduce, C.c_void_p(status_ptr), C.byref(self.c),
                                                C.c_void_p(stream))
        if rc != C_KZG_OK:
            raise KzgError(rc, "compute_kzg_proof_device")


def load_trusted_setup_file(path=None):
    return KZGSettings.load_trusted_setup_file(path)


def sha256(msg: bytes, portable=False) -> bytes:
    out = np.zeros(32, np.uint8)
    m = np.frombuffer(msg, dtype=np.uint8) if len(msg) else np.zeros(1, np.uint8)
    _L().b200_selftest_sha256(_p(out), _p(m), len(msg), int(portable))
    return out.tobytes()
