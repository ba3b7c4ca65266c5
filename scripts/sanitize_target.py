"""Small end-to-end pass over the verification / EIP-7594 kernels for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B

rng = np.random.default_rng(9)
ts = B.KZGSettings.load_trusted_setup_file()
blobs = rng.integers(0, 256, size=(3, 4096, 32), dtype=np.uint8)
blobs[:, :, 0] = 0
blobs = blobs.reshape(3, -1)
comm = ts.blob_to_kzg_commitment_batch(blobs)
proofs = ts.compute_blob_kzg_proof_batch(blobs, comm)
assert ts.verify_blob_kzg_proof_batch(blobs, comm, proofs)
assert ts.verify_blob_kzg_proof(blobs[0], comm[0], proofs[0])
assert not ts.verify_blob_kzg_proof(blobs[0], comm[0], proofs[1])
if len(sys.argv) > 1 and sys.argv[1] == "cells":
    cells, cproofs = ts.compute_cells_and_kzg_proofs(blobs[0].tobytes())
    idx = list(range(0, 128, 2))
    rc, rp = ts.recover_cells_and_kzg_proofs(idx, [cells[i] for i in idx])
    assert rc == cells and rp == cproofs
    assert ts.verify_cell_kzg_proof_batch([comm[0].tobytes()] * 5, [0, 7, 7, 100, 127], [cells[i] for i in (0, 7, 7, 100, 127)],
                                          [cproofs[i] for i in (0, 7, 7, 100, 127)])
fs = B.FFTSettings(13)
a = rng.integers(0, 1 << 62, size=(8192, 4), dtype=np.uint64)
assert np.array_equal(fs.fft_fr(fs.fft_fr(a, False), True), a)
ts.free()
print("sanitize target ok")
