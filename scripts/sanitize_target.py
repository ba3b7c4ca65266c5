"""Small end-to-end pass over the verification / EIP-7594 kernels for compute-sanitizer (memcheck / racecheck).
Modes: (none) blob pipeline + NTT; `cells` adds EIP-7594; `widemsm` adds the wide-window MSM kernels."""
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
if len(sys.argv) > 1 and sys.argv[1] == "widemsm":
    # the wide-window kernels (k_segment_fold, the R row of k_marginals, mixed window widths) on a small problem
    from oracle import c_oracle as K
    text = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
    L = K.p1s_to_affine(K.KZGSettings(text).g1_lagrange_brp)
    sc = rng.integers(0, 1 << 62, size=(4096, 4), dtype=np.uint64)
    exp = K.p1_compress(K.msm_affine(L, sc))
    for c, c0 in ((20, 0), (17, 0), (13, 9)):
        os.environ["B200_MSM_C"], os.environ["B200_MSM_C0"] = str(c), str(c0)
        h = B.PreparedMsm(L)
        assert K.p1_compress(h.mult(sc)) == exp, (c, c0)
        h.close()
    del os.environ["B200_MSM_C"], os.environ["B200_MSM_C0"]
fs = B.FFTSettings(13)
a = rng.integers(0, 1 << 62, size=(8192, 4), dtype=np.uint64)
assert np.array_equal(fs.fft_fr(fs.fft_fr(a, False), True), a)
ts.free()
print("sanitize target ok")
