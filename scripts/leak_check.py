"""load / use / free cycles: device memory must return to its starting level (run under gpurun)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B

rng = np.random.default_rng(1)
blob = rng.integers(0, 256, size=(4096, 32), dtype=np.uint8)
blob[:, 0] = 0
blob = blob.tobytes()
torch.cuda.init()
torch.zeros(1, device="cuda")
levels = []
for it in range(6):
    free0, _ = torch.cuda.mem_get_info()
    ts = B.KZGSettings.load_trusted_setup_file()
    c = ts.blob_to_kzg_commitment(blob)
    p = ts.compute_blob_kzg_proof(blob, c)
    assert ts.verify_blob_kzg_proof(blob, c, p)
    cells, proofs = ts.compute_cells_and_kzg_proofs(blob)
    assert ts.verify_cell_kzg_proof_batch([c] * 4, [0, 1, 2, 3], cells[:4], proofs[:4])
    ts.recover_cells_and_kzg_proofs(list(range(64)), cells[:64])
    used, _ = torch.cuda.mem_get_info()
    ts.free()
    msm = B.PreparedMsm(np.zeros((4096, 12), np.uint64))
    msm.close()
    fs = B.FFTSettings(16)
    fs.close()
    free1, _ = torch.cuda.mem_get_info()
    levels.append((free0 - used) >> 20)
    print("cycle", it, "in use while loaded: %d MiB" % ((free0 - used) >> 20), "leaked: %d KiB" % ((free0 - free1) >> 10))
