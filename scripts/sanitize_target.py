"""Small end-to-end pass over the verification / EIP-7594 kernels for compute-sanitizer (memcheck / racecheck).
Modes: (none) blob pipeline + NTT; `cells` adds EIP-7594; `widemsm` adds the wide-window MSM kernels; `round2` adds what
round 2 built: the proof engine on a 30-blob batch (bucket engine) next to the direct-lookup path (3 blobs), the coalesced
single-blob calls from 8 threads, scalar randomisation with and without the subgroup guard, the batch-affine accumulation
(both inversion-sharing modes) and the Karatsuba accumulate variant on a 2^17-point table, the sharded entry point (world 1),
the three-pass NTT."""
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
if len(sys.argv) > 1 and sys.argv[1] == "round2":
    import threading
    from oracle import c_oracle as K
    text = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
    osettings = K.KZGSettings(text, nthreads=8)
    L = K.p1s_to_affine(osettings.g1_lagrange_brp)
    # proof engine + bucket path (batch above the direct limit) and the direct path (3 blobs, above)
    big = rng.integers(0, 256, size=(30, 4096, 32), dtype=np.uint8)
    big[:, :, 0] = 0
    big = big.reshape(30, -1)
    zs = big[0].reshape(4096, 32)[:30].copy()
    c30 = ts.blob_to_kzg_commitment_batch(big)
    p30, y30 = ts.compute_kzg_proof_batch(big, zs)
    assert bytes(c30[29]) == K.blob_to_kzg_commitment(big[29].tobytes(), osettings)
    assert (bytes(p30[7]), bytes(y30[7])) == tuple(K.compute_kzg_proof(big[7].tobytes(), zs[7].tobytes(), osettings))
    # coalesced single-blob calls
    errs = []

    def worker(k):
        try:
            for rep in range(3):
                i = (k + rep) % 3
                assert ts.blob_to_kzg_commitment(blobs[i]) == bytes(comm[i])
                assert ts.compute_blob_kzg_proof(blobs[i], comm[i]) == bytes(proofs[i])
        except Exception as e:
            errs.append(repr(e))
    th = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    # 2^17-point table: randomised XYZZ path, batch-affine (warp / tree), Karatsuba accumulate runs in its own process
    n = 1 << 17
    pts = np.tile(L, (n // 4096, 1))
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    cur = sc
    while cur.shape[0] > 4096:
        cur = K.fr_add(cur[: cur.shape[0] // 2], cur[cur.shape[0] // 2:])
    exp = K.p1_compress(K.msm_affine(L, np.ascontiguousarray(cur), nthreads=8))
    for env in ({}, {"B200_MSM_RANDOMIZE": "0"}, {"B200_MSM_AFFINE": "1"}, {"B200_MSM_AFFINE": "1", "B200_MSM_AFFINE_TREE": "1"}):
        os.environ.update(env)
        h = B.PreparedMsm(pts)
        assert K.p1_compress(h.mult(sc)) == exp, env
        h.close()
        for k in env:
            del os.environ[k]
    sm = B.ShardedMsm(pts[:8192], 0, 1)
    assert K.p1_compress(sm.mult(sc[:8192])) == K.p1_compress(K.msm_affine(pts[:8192], sc[:8192], nthreads=8))
    sm.close()
    fs23 = B.FFTSettings(23)
    big_a = rng.integers(0, 1 << 62, size=(1 << 23, 4), dtype=np.uint64)
    assert np.array_equal(fs23.fft_fr(fs23.fft_fr(big_a, False), True), big_a)
    fs23.close()
fs = B.FFTSettings(13)
a = rng.integers(0, 1 << 62, size=(8192, 4), dtype=np.uint64)
assert np.array_equal(fs.fft_fr(fs.fft_fr(a, False), True), a)
# the cluster / distributed-shared-memory kernel (2^9 .. 2^12: clusters of 8 and 16 CTAs) against the two-launch form
for k in (9, 10, 11, 12):
    v = a[:1 << k]
    one = (fs.fft_fr(v, False), fs.fft_fr(v, True), fs.das_fft_extension(v))
    os.environ["B200_NTT_CLUSTER"] = "0"
    two = (fs.fft_fr(v, False), fs.fft_fr(v, True), fs.das_fft_extension(v))
    del os.environ["B200_NTT_CLUSTER"]
    assert all(np.array_equal(x, y) for x, y in zip(one, two)), k
# prepare_msm on 4096 points: the direct-lookup table with a Jacobian result, one vector and a batch of three
if True:
    from oracle import c_oracle as K2
    text2 = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
    L2 = K2.p1s_to_affine(K2.KZGSettings(text2).g1_lagrange_brp)
    os.environ["B200_MSM_DIRECT_BITS"] = "8"
    h2 = B.PreparedMsm(L2)
    del os.environ["B200_MSM_DIRECT_BITS"]
    assert h2.info()["direct_bits"] == 8
    s3 = rng.integers(0, 1 << 62, size=(3 * 4096, 4), dtype=np.uint64)
    got3 = h2.mult_batch(s3, 3)
    assert K2.p1_compress(got3[2]) == K2.p1_compress(K2.msm_affine(L2, s3[2 * 4096:], nthreads=8))
    assert K2.p1_compress(h2.mult(s3[:4096])) == K2.p1_compress(got3[0])
    h2.close()
if len(sys.argv) > 1 and sys.argv[1] == "cells":
    # coalesced compute_cells_and_kzg_proofs: three callers at once share a pass
    import threading
    errs2 = []

    def cworker():
        try:
            c2, p2 = ts.compute_cells_and_kzg_proofs(blobs[0].tobytes())
            assert c2 == cells and p2 == cproofs
        except Exception as e:
            errs2.append(repr(e))
    th2 = [threading.Thread(target=cworker) for _ in range(3)]
    [t.start() for t in th2]
    [t.join() for t in th2]
    assert not errs2, errs2
ts.free()
print("sanitize target ok")
