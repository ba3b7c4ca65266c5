"""Wall-clock latencies of the verification / recovery entry points through the C ABI (host buffers in, bool out).
Run under gpurun; prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e3


def main():
    rng = np.random.default_rng(5)
    ts = B.KZGSettings.load_trusted_setup_file()
    n = 64
    blobs = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    blobs = blobs.reshape(n, -1)
    comm = ts.blob_to_kzg_commitment_batch(blobs)
    proofs = ts.compute_blob_kzg_proof_batch(blobs, comm)
    zs = blobs[:, :32].copy()
    pz, ys = ts.compute_kzg_proof_batch(blobs, zs)
    out = {}
    g1 = ts.array("g1_values_monomial", 4096, 18)
    assert ts.pairings_verify(g1[1], 0, g1[0], 1)
    out["pairing_check_2pairs_ms"] = timeit(lambda: ts.pairings_verify(g1[1], 0, g1[0], 1), 20)
    assert ts.verify_kzg_proof(comm[0], zs[0], ys[0], pz[0])
    out["verify_kzg_proof_ms"] = timeit(lambda: ts.verify_kzg_proof(comm[0], zs[0], ys[0], pz[0]), 20)
    assert ts.verify_blob_kzg_proof(blobs[0], comm[0], proofs[0])
    out["verify_blob_kzg_proof_ms"] = timeit(lambda: ts.verify_blob_kzg_proof(blobs[0], comm[0], proofs[0]), 20)
    assert ts.verify_blob_kzg_proof_batch(blobs, comm, proofs)
    ms = timeit(lambda: ts.verify_blob_kzg_proof_batch(blobs, comm, proofs), 10)
    out["verify_blob_kzg_proof_batch_64_ms"] = ms
    out["verify_blob_kzg_proof_batch_64_blobs_per_s"] = n / ms * 1e3
    cells, cproofs = ts.compute_cells_and_kzg_proofs(blobs[0].tobytes())
    c0 = comm[0].tobytes()
    idx = list(range(128))
    assert ts.verify_cell_kzg_proof_batch([c0] * 128, idx, cells, cproofs)
    out["verify_cell_kzg_proof_batch_128_ms"] = timeit(lambda: ts.verify_cell_kzg_proof_batch([c0] * 128, idx, cells, cproofs), 10)
    half = list(range(0, 128, 2))
    hc = [cells[i] for i in half]
    rc, rp = ts.recover_cells_and_kzg_proofs(half, hc)
    assert rc == cells and rp == cproofs
    out["recover_cells_and_kzg_proofs_ms"] = timeit(lambda: ts.recover_cells_and_kzg_proofs(half, hc), 10)
    out["recover_cells_only_ms"] = timeit(lambda: ts.recover_cells_and_kzg_proofs(half, hc, want_proofs=False), 10)
    out["compute_cells_and_kzg_proofs_1_ms"] = timeit(lambda: ts.compute_cells_and_kzg_proofs(blobs[0].tobytes()), 10)
    print(json.dumps(out, indent=1))
    ts.free()


if __name__ == "__main__":
    main()
