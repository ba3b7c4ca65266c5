"""Small target for ncu launch lists: a few invocations of one hot-path call (not a bench)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B
from oracle import c_oracle as K

what, logn, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(1)
n = 1 << logn if what in ('msm', 'ntt') else logn


def rand_fr(n):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    return a


if what == "msm":
    text = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
    L = K.p1s_to_affine(K.KZGSettings(text).g1_lagrange_brp)
    pts = np.tile(L, (max(1, n // 4096), 1))[:n]
    h = B.PreparedMsm(pts)
    d_sc = torch.from_numpy(rand_fr(n).view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    for _ in range(reps):
        h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
    torch.cuda.synchronize()
elif what == "ntt":
    fs = B.FFTSettings(max(logn, 1))
    d_in = torch.from_numpy(rand_fr(n).view(np.int64)).cuda()
    d_out = torch.zeros_like(d_in)
    for _ in range(reps):
        fs.fft_fr_device(d_out.data_ptr(), d_in.data_ptr(), n, False, 1, 0)
    torch.cuda.synchronize()
elif what == "blob":
    ts = B.KZGSettings.load_trusted_setup_file()
    nb = n if n <= 64 else 64
    blobs = rng.integers(0, 256, size=(nb, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    d_blobs = torch.from_numpy(blobs.reshape(nb, 131072)).cuda()
    d_out = torch.zeros((nb, 48), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(nb, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        ts.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), nb, d_st.data_ptr(), 0)
    torch.cuda.synchronize()
elif what == "proof":
    ts = B.KZGSettings.load_trusted_setup_file()
    nb = 64
    blobs = rng.integers(0, 256, size=(nb, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    d_blobs = torch.from_numpy(blobs.reshape(nb, 131072)).cuda()
    d_z = torch.from_numpy(blobs.reshape(nb, 131072)[0, :32 * nb].copy()).cuda()
    d_out = torch.zeros((nb, 48), dtype=torch.uint8, device="cuda")
    d_y = torch.zeros((nb, 32), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(nb, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        ts.compute_kzg_proof_device(d_out.data_ptr(), d_y.data_ptr(), d_blobs.data_ptr(), d_z.data_ptr(), nb, d_st.data_ptr(), 0, 0)
    torch.cuda.synchronize()
elif what in ("cells", "fk20"):
    ts = B.KZGSettings.load_trusted_setup_file()
    nb = n
    blobs = rng.integers(0, 256, size=(nb, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    blobs = blobs.reshape(nb, 131072)
    if what == "fk20":
        ts.compute_cell_proofs_batch(blobs[:1])      # builds the FK20 table outside the launches of interest
    for _ in range(reps):
        (ts.compute_cells_batch if what == "cells" else ts.compute_cell_proofs_batch)(blobs)
    torch.cuda.synchronize()
elif what == "verify":
    ts = B.KZGSettings.load_trusted_setup_file()
    nb = n
    blobs = rng.integers(0, 256, size=(nb, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    blobs = blobs.reshape(nb, 131072)
    comm = ts.blob_to_kzg_commitment_batch(blobs)
    proofs = ts.compute_blob_kzg_proof_batch(blobs, comm)
    for _ in range(reps):
        assert ts.verify_blob_kzg_proof_batch(blobs, comm, proofs)
    torch.cuda.synchronize()
elif what in ("verifycells", "recover"):
    ts = B.KZGSettings.load_trusted_setup_file()
    blob = rng.integers(0, 256, size=(4096, 32), dtype=np.uint8)
    blob[:, 0] = 0
    blob = blob.tobytes()
    cells, proofs = ts.compute_cells_and_kzg_proofs(blob)
    c0 = ts.blob_to_kzg_commitment(blob)
    for _ in range(reps):
        if what == "verifycells":
            assert ts.verify_cell_kzg_proof_batch([c0] * 128, list(range(128)), cells, proofs)
        else:
            half = list(range(0, 128, 2))
            ts.recover_cells_and_kzg_proofs(half, [cells[i] for i in half], want_proofs=False)
    torch.cuda.synchronize()
