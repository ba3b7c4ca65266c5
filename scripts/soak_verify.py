"""Randomised soak of the verification / recovery entry points against the oracle (not part of the test suite; run under
gpurun).  Every verdict and every recovered byte must agree; prints one JSON summary."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B
from oracle import c_oracle as K

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(seed)
text = open(B.default_trusted_setup_path()).read()
os_ = K.KZGSettings(text, nthreads=os.cpu_count() or 1)
ts = B.KZGSettings.load_trusted_setup_file()
t0 = time.time()
stats = {"verify_kzg_proof": 0, "verify_blob_batch": 0, "verify_cells": 0, "recover": 0, "true": 0, "false": 0, "err": 0}


def rand_blobs(n):
    b = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    b[:, :, 0] = 0
    return b.reshape(n, -1)


def both(fn_b, fn_k):
    try:
        g = fn_b()
    except (B.KzgError, ValueError):
        g = None
    try:
        w = fn_k()
    except (K.OracleError, ValueError):
        w = None
    assert g == w, (g, w)
    stats["true" if g is True else "false" if g is False else "err"] += 1
    return g


nb = 12
blobs = rand_blobs(nb)
comm = ts.blob_to_kzg_commitment_batch(blobs)
proofs = ts.compute_blob_kzg_proof_batch(blobs, comm)
# single-point proofs, honest and perturbed in every argument
for it in range(60):
    i = int(rng.integers(nb))
    z = int.from_bytes(rng.bytes(32), "big") % R
    zb = z.to_bytes(32, "big")
    p, y = ts.compute_kzg_proof(blobs[i], zb)
    c = comm[i].tobytes()
    mode = it % 6
    if mode == 1:
        y = ((int.from_bytes(y, "big") + 1) % R).to_bytes(32, "big")
    elif mode == 2:
        zb = ((z + 1) % R).to_bytes(32, "big")
    elif mode == 3:
        c = comm[(i + 1) % nb].tobytes()
    elif mode == 4:
        p = proofs[i].tobytes()
    elif mode == 5:
        y = (R + int(rng.integers(5))).to_bytes(32, "big")          # non-canonical: error on both sides
    both(lambda: ts.verify_kzg_proof(c, zb, y, p), lambda: K.verify_kzg_proof(c, zb, y, p, os_))
    stats["verify_kzg_proof"] += 1
# blob batches of random size with random corruption
for it in range(12):
    n = int(rng.integers(1, nb + 1))
    sel = rng.choice(nb, size=n, replace=False)
    bl, cm, pr = blobs[sel].copy(), comm[sel].copy(), proofs[sel].copy()
    mode = it % 4
    if mode == 1:
        pr[int(rng.integers(n))] = proofs[(sel[0] + 1) % nb]
    elif mode == 2:
        bl[int(rng.integers(n)), int(rng.integers(1, 131072))] ^= 1
    elif mode == 3:
        bl[int(rng.integers(n)), 0] = 0xFF                             # field element >= r: error
    both(lambda: ts.verify_blob_kzg_proof_batch(bl, cm, pr),
         lambda: K.verify_blob_kzg_proof_batch([x.tobytes() for x in bl], [x.tobytes() for x in cm], [x.tobytes() for x in pr], os_))
    stats["verify_blob_batch"] += 1
# cells: recovery from random subsets, cell-proof verification with random corruption
for it in range(6):
    i = int(rng.integers(nb))
    cells, cproofs = ts.compute_cells_and_kzg_proofs(blobs[i].tobytes())
    k = int(rng.integers(64, 129))
    keep = sorted(rng.choice(128, size=k, replace=False).tolist())
    got = ts.recover_cells_and_kzg_proofs(keep, [cells[j] for j in keep])
    assert got[0] == cells and got[1] == cproofs
    if it < 2:
        want = K.recover_cells_and_kzg_proofs(keep, [cells[j] for j in keep], os_)
        assert want[0] == cells and want[1] == cproofs
    stats["recover"] += 1
    for rep in range(4):
        m = int(rng.integers(1, 40))
        idx = rng.choice(128, size=m, replace=True).tolist()
        cs = [cells[j] for j in idx]
        ps = [cproofs[j] for j in idx]
        cm = [comm[i].tobytes()] * m
        if rep == 1:
            ps[0] = cproofs[(idx[0] + 1) % 128]
        elif rep == 2:
            cs[m - 1] = cells[(idx[m - 1] + 5) % 128]
        elif rep == 3:
            cm[m // 2] = comm[(i + 1) % nb].tobytes()
        both(lambda: ts.verify_cell_kzg_proof_batch(cm, idx, cs, ps), lambda: K.verify_cell_kzg_proof_batch(cm, idx, cs, ps, os_))
        stats["verify_cells"] += 1
stats["seconds"] = round(time.time() - t0, 1)
stats["seed"] = seed
print(json.dumps(stats))
