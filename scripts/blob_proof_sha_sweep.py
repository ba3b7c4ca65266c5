"""compute_blob_kzg_proof_batch (64 blobs, host to host) against the number of host threads hashing the Fiat-Shamir challenges
(B200_SHA_THREADS); run under gpurun."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B

rng = np.random.default_rng(1)
ts = B.KZGSettings.load_trusted_setup_file()
blobs = rng.integers(0, 256, size=(64, 4096, 32), dtype=np.uint8)
blobs[:, :, 0] = 0
blobs = blobs.reshape(64, -1)
comm = ts.blob_to_kzg_commitment_batch(blobs)
ref = None
for nt in (1, 2, 4, 8, 16):
    os.environ["B200_SHA_THREADS"] = str(nt)
    p = ts.compute_blob_kzg_proof_batch(blobs, comm)
    for _ in range(3):
        ts.compute_blob_kzg_proof_batch(blobs, comm)
    t = time.perf_counter()
    for _ in range(10):
        ts.compute_blob_kzg_proof_batch(blobs, comm)
    ms = (time.perf_counter() - t) / 10 * 1e3
    if ref is None:
        ref = p.tobytes()
    print("sha threads", nt, "%.3f ms per 64 blobs" % ms, "same" if p.tobytes() == ref else "DIFF", flush=True)
