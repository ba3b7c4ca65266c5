"""Handles are Send + Sync on the Rust side (kzg/src/msm/sppark.rs:24-44) and the c-kzg settings are shared by rayon
workers (kzg/src/eip_4844.rs:781-815): concurrent calls on one settings object / one prepared MSM must serialise
internally and stay bit-exact; several settings objects can coexist and be freed independently."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blobs(rng, n):
    b = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    b[:, :, 0] = 0
    return b.reshape(n, -1)


def test_concurrent_calls_on_one_settings_object(B):
    ts = B.KZGSettings.load_trusted_setup_file()
    rng = np.random.default_rng(31)
    blobs = _blobs(rng, 6)
    want_c = [ts.blob_to_kzg_commitment(b) for b in blobs]
    want_p = [ts.compute_blob_kzg_proof(b, c) for b, c in zip(blobs, want_c)]
    errors = []

    def worker(k):
        try:
            for rep in range(3):
                i = (k + rep) % len(blobs)
                assert ts.blob_to_kzg_commitment(blobs[i]) == want_c[i]
                assert ts.compute_blob_kzg_proof(blobs[i], want_c[i]) == want_p[i]
                assert ts.verify_blob_kzg_proof(blobs[i], want_c[i], want_p[i]) is True
                assert ts.verify_blob_kzg_proof(blobs[i], want_c[i], want_p[(i + 1) % len(blobs)]) is False
                cells = ts.compute_cells(blobs[i])
                assert ts.recover_cells_and_kzg_proofs(list(range(0, 128, 2)), cells[0::2], want_proofs=False)[0] == cells
        except Exception as e:  # surfaced in the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    ts.free()
    assert not errors, errors


def test_two_settings_objects_and_reload(B):
    rng = np.random.default_rng(32)
    blob = _blobs(rng, 1)[0]
    a = B.KZGSettings.load_trusted_setup_file()
    b = B.KZGSettings.load_trusted_setup_file()
    ca, cb = a.blob_to_kzg_commitment(blob), b.blob_to_kzg_commitment(blob)
    assert ca == cb
    a.free()
    with pytest.raises(B.KzgError):
        a.blob_to_kzg_commitment(blob)                 # freed settings: BADARGS, not a crash
    assert b.verify_blob_kzg_proof(blob, cb, b.compute_blob_kzg_proof(blob, cb)) is True
    for _ in range(3):                                 # load / free cycles release the device context
        c = B.KZGSettings.load_trusted_setup_file()
        assert c.blob_to_kzg_commitment(blob) == cb
        c.free()
    b.free()


def test_concurrent_prepared_msm(B, K, lagrange_affine):
    rng = np.random.default_rng(33)
    msm = B.PreparedMsm(lagrange_affine)
    scs = [rng.integers(0, 1 << 62, size=(4096, 4), dtype=np.uint64) for _ in range(4)]
    want = [K.p1_compress(msm.mult(s)) for s in scs]
    assert want[0] == K.p1_compress(K.msm_affine(lagrange_affine, scs[0], nthreads=4))
    errors = []

    def worker(k):
        try:
            for rep in range(4):
                i = (k + rep) % 4
                assert K.p1_compress(msm.mult(scs[i])) == want[i]
        except Exception as e:
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    msm.close()
    assert not errors, errors


def test_sixteen_threads_mixed_entry_points_coalesce_bit_exact(B, K, oracle_settings):
    """16 host threads on the UNMODIFIED single-blob c-kzg symbols (the way rayon's par_chunks drives them,
    kzg/src/eip_4844.rs:770-816): concurrent requests are coalesced into shared launch sequences (csrc/coalesce.cuh) and
    every caller must still get exactly its own result -- checked against the oracle, with invalid blobs mixed in, which
    must fail alone."""
    ts = B.KZGSettings.load_trusted_setup_file()
    rng = np.random.default_rng(34)
    blobs = _blobs(rng, 8)
    zs = [bytes(b[64:96]) for b in blobs]
    want_c = [K.blob_to_kzg_commitment(bytes(b), oracle_settings) for b in blobs]
    want_p = [K.compute_blob_kzg_proof(bytes(b), c, oracle_settings) for b, c in zip(blobs, want_c)]
    want_z = [K.compute_kzg_proof(bytes(b), z, oracle_settings) for b, z in zip(blobs, zs)]
    bad = blobs[0].copy()
    bad[32 * 7] = 0xFF                                  # element 7 >= r
    errors = []
    start = threading.Barrier(16)

    def worker(k):
        try:
            start.wait()
            for rep in range(6):
                i = (k + rep) % len(blobs)
                op = (k + rep) % 4
                if op == 0:
                    assert ts.blob_to_kzg_commitment(blobs[i]) == want_c[i]
                elif op == 1:
                    assert ts.compute_blob_kzg_proof(blobs[i], want_c[i]) == want_p[i]
                elif op == 2:
                    assert tuple(ts.compute_kzg_proof(blobs[i], zs[i])) == tuple(want_z[i])
                else:
                    with pytest.raises(B.KzgError) as e:
                        ts.blob_to_kzg_commitment(bad)
                    assert e.value.code == 1
                    assert ts.verify_blob_kzg_proof(blobs[i], want_c[i], want_p[i]) is True
        except Exception as e:
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(16)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    ts.free()
    assert not errors, errors


def test_concurrent_cells_and_proofs_coalesce_bit_exact(B, K, oracle_settings):
    """compute_cells_and_kzg_proofs called per blob from nine threads (a block's blobs under a parallel iterator) shares FK20
    passes (csrc/capi_ckzg.cu, coalesced_cells_call): every caller gets exactly the cells and proofs of the sequential call,
    blob 0's equal the oracle's, an invalid blob fails alone, and calls mixed in from the other entry points still work"""
    ts = B.KZGSettings.load_trusted_setup_file()
    rng = np.random.default_rng(35)
    blobs = _blobs(rng, 6)
    want = [ts.compute_cells_and_kzg_proofs(b) for b in blobs]            # sequential: batches of one
    oracle_settings.set_threads(8)
    oc, op = K.compute_cells_and_kzg_proofs(bytes(blobs[0]), oracle_settings)
    oracle_settings.set_threads(1)
    assert want[0][0] == oc and want[0][1] == op
    want_c = [ts.blob_to_kzg_commitment(b) for b in blobs]
    bad = blobs[1].copy()
    bad[32 * 100] = 0xFF
    b0, r0 = ts.cells_coalesce_stats()
    errors = []
    start = threading.Barrier(9)

    def worker(k):
        try:
            start.wait()
            for rep in range(4):
                i = (k + rep) % len(blobs)
                if k == 8 and rep % 2 == 0:
                    with pytest.raises(B.KzgError) as e:
                        ts.compute_cells_and_kzg_proofs(bad)
                    assert e.value.code == 1
                    assert ts.blob_to_kzg_commitment(blobs[i]) == want_c[i]
                else:
                    cells, proofs = ts.compute_cells_and_kzg_proofs(blobs[i])
                    assert cells == want[i][0] and proofs == want[i][1], (k, rep)
        except Exception as e:
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(9)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    b1, r1 = ts.cells_coalesce_stats()
    ts.free()
    assert not errors, errors
    assert r1 - r0 == 36 and b1 - b0 < 36, (b0, r0, b1, r1)              # some batches carried more than one request


def test_concurrent_single_verifications_coalesce_exactly(B, K, oracle_settings):
    """verify_blob_kzg_proof / verify_kzg_proof called one at a time from twelve threads are checked as one batch
    (csrc/capi_ckzg.cu, coalesced_verify) and re-checked alone when the batch does not pass: valid requests must come back
    True, a swapped proof False, an undecodable commitment as an error -- each caller its own answer, whatever else shares
    the batch"""
    ts = B.KZGSettings.load_trusted_setup_file()
    rng = np.random.default_rng(36)
    blobs = _blobs(rng, 4)
    comm = [ts.blob_to_kzg_commitment(b) for b in blobs]
    proofs = [ts.compute_blob_kzg_proof(b, c) for b, c in zip(blobs, comm)]
    zs = [bytes(b[96:128]) for b in blobs]
    zp = [ts.compute_kzg_proof(b, z) for b, z in zip(blobs, zs)]
    bad_comm = bytes(48)                                               # compression flag missing: does not decode
    errors = []

    def run(threads, mode):
        start = threading.Barrier(threads)

        def worker(k):
            try:
                start.wait()
                for rep in range(5):
                    i = (k + rep) % 4
                    j = (i + 1) % 4
                    if mode == "valid" or k % 4 == 0:
                        assert ts.verify_blob_kzg_proof(blobs[i], comm[i], proofs[i]) is True
                    elif k % 4 == 1:
                        assert ts.verify_blob_kzg_proof(blobs[i], comm[i], proofs[j]) is False
                    elif k % 4 == 2:
                        assert ts.verify_kzg_proof(comm[i], zs[i], zp[i][1], zp[i][0]) is True
                        assert ts.verify_kzg_proof(comm[i], zs[i], zp[j][1], zp[i][0]) is False
                    else:
                        with pytest.raises(B.KzgError) as e:
                            ts.verify_blob_kzg_proof(blobs[i], bad_comm, proofs[i])
                        assert e.value.code == 1
            except Exception as e:
                errors.append(repr(e))

        th = [threading.Thread(target=worker, args=(k,)) for k in range(threads)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    b0, r0, f0 = ts.verify_coalesce_stats()
    run(12, "valid")
    b1, r1, f1 = ts.verify_coalesce_stats()
    assert not errors, errors
    assert r1 - r0 == 60 and b1 - b0 < 60 and f1 == f0, (b0, r0, f0, b1, r1, f1)   # shared batches, none re-checked
    run(12, "mixed")
    b2, r2, f2 = ts.verify_coalesce_stats()
    ts.free()
    assert not errors, errors
    assert f2 > f1, (f1, f2)                                           # batches with a bad request were re-checked one by one


def test_concurrent_prepared_msm_coalesces(B, K, lagrange_affine):
    """a prepared 4096-point handle packs concurrent mult_pippenger_prepared calls (different lengths included: short
    calls are zero-padded) into one launch sequence; each caller gets its own sum"""
    rng = np.random.default_rng(35)
    msm = B.PreparedMsm(lagrange_affine)
    lens = [4096, 4096, 1000, 8, 4096, 2500, 4096, 1]
    scs = [rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64) for n in lens]
    want = [K.p1_compress(K.msm_affine(lagrange_affine[:n], s, nthreads=4)) for n, s in zip(lens, scs)]
    errors = []
    start = threading.Barrier(8)

    def worker(k):
        try:
            start.wait()
            for rep in range(4):
                i = (k + rep) % len(lens)
                assert K.p1_compress(msm.mult(scs[i])) == want[i], (k, rep)
        except Exception as e:
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    msm.close()
    assert not errors, errors
